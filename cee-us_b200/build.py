"""Builds libceeus_b200.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libceeus_b200.so")
SOURCES = ["ceeus_api.cu", "rollout_simt.cu", "icem_kernels.cu", "tc_probe.cu", "rollout_tc.cu", "rollout_mlp_tc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-shared"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ceeus_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compiles every translation unit for sm_100a in parallel (one nvcc per source) and links the shared library."""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cflags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math") and f != "-shared"]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + cflags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        return obj, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    for obj, res in results:
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed building libceeus_b200.so")
        if verbose:
            print(res.stderr)
    link = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"] +
                          [obj for obj, _ in results] + ["-o", LIB], capture_output=True, text=True)
    if link.returncode != 0:
        sys.stderr.write(link.stdout + link.stderr)
        raise RuntimeError("linking libceeus_b200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
