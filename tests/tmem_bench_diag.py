"""Manual diagnostic: TMEM <-> register transfer cycle counts (ceeus_selftest_tmem)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ceeus_b200
lib = ceeus_b200._lib.load()
torch.zeros(1, device="cuda")
names = ["ld x1 1w", "ld x4 1w", "ld x16 1w", "ld x64 1w", "ld x4 4w", "ld x16 4w", "ld x64 4w", "ld x16 8w", "ld x64 8w",
         "ld+wait x16 1w", "ld+wait x16 4w", "st x1 1w", "st x16 4w", "st x64 4w", "st x64 8w"]
for rep in range(2):
    buf = (C.c_longlong * 32)()
    rc = lib.ceeus_selftest_tmem(buf)
    print("rc", rc, " | ".join(f"{n}: {buf[i]}" for i, n in enumerate(names)))

for nmma in (0, 1, 4, 8, 16, 32):
    buf = (C.c_longlong * 8)()
    rc = lib.ceeus_selftest_tmem_contention(nmma, buf)
    print(f"nmma={nmma:2d} rc={rc} ld_latency={buf[0]} st_latency={buf[1]} mma_batch_done={buf[2]} ld_issued_at={buf[3]}")

for mma in (0, 1):
    for bias in (0, 1, 2):  # bit 0: bias add, bit 1: folded LayerNorm affine
        for nw in (1, 2, 3):
            v = (C.c_longlong * 2)()
            rc = lib.ceeus_selftest_epilogue(nw, bias, mma, v)
            print(f"epilogue mma={mma} bias={bias} warps/scheduler={nw}: {v[0]} cycles per epilogue per warp -> {v[0] / nw:.0f} per scheduler; mmas {v[1]}")
