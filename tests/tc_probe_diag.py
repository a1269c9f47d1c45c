import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from test_gpu_tc_probe import run_probe
cg, K, N = (int(x) for x in sys.argv[1:4])
variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rc, st, D, want = run_probe(cg, K, N, variant)
err = (D - want).abs().max().item() if rc == 0 else float('nan')
print(f"variant={variant} cg={cg} K={K} N={N} rc={rc} status={st} maxerr={err:.4g} ref_max={want.abs().max().item():.3g} nan={torch.isnan(D).sum().item()}", flush=True)
