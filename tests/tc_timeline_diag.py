"""Manual diagnostic: per-stage clock64 timeline of one warpgroup of the TC rollout kernel."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import conftest  # noqa
from cases import CASES
from helpers import build_inputs, product_env, product_model, rollout_actions
n_obj = int(sys.argv[1]) if len(sys.argv) > 1 else 2
P = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
c = dict(kind="rollout", env="construction", n_obj=n_obj, model="gnn", hidden=128, layers=2, P=P, H=30, wseed=1, identity_norm=True, running=False, oseed=1)
spec, weights, norms, obs, goal = build_inputs(c)
env = product_env(c, goal)
m = product_model(c, env, weights, norms, precision="tc")
eng = m._rollout_engine(P, 30)
acts = torch.from_numpy(rollout_actions(c, spec)).cuda()
start = torch.from_numpy(obs).view(1, -1).cuda()
eng.rollout(start, acts); torch.cuda.synchronize()
eng.lib.ceeus_tc_debug_timeline(eng.handle, 1, None, 0)
flag = ((1 if len(sys.argv) > 3 and sys.argv[3] == 'single' else 0) | (int(os.environ.get('TL_TID', '0')) << 1) |
        (int(os.environ.get('TL_RD', '0')) << 12) | (int(os.environ.get('TL_BLOCK', '0')) << 16))
eng.lib.ceeus_tc_debug_flag(eng.handle, flag)
eng.rollout(start, acts); torch.cuda.synchronize()
buf = (C.c_longlong * 500)()
eng.lib.ceeus_tc_debug_timeline(eng.handle, 1, buf, 250)
a = np.array(buf[:]).reshape(250, 2)
t0 = a[0, 1]
prev = t0
for tag, t in a[:120]:
    if tag == 0: break
    print(int(tag), int(t - t0), int(t - prev)); prev = t
