"""Manual diagnostic (not collected by pytest): TC rollout vs the fp32 CUDA path and the reference fixtures."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import conftest  # noqa
from cases import CASES
from helpers import GOLD, build_inputs, product_env, product_model, rel_err, rollout_actions

name = sys.argv[1] if len(sys.argv) > 1 else "rollout_gnn_c2"
P_override = int(sys.argv[2]) if len(sys.argv) > 2 else 0
H_override = int(sys.argv[3]) if len(sys.argv) > 3 else 0
c = dict(CASES[name])
if P_override: c["P"] = P_override
if H_override: c["H"] = H_override
spec, weights, norms, obs, goal = build_inputs(c)
env = product_env(c, goal)
acts = torch.from_numpy(rollout_actions(c, spec)).cuda()
start = torch.from_numpy(obs).view(1, -1).cuda()
out = {}
for prec in ("fp32", "tc"):
    m = product_model(c, env, weights, norms, precision=prec)
    eng = m._rollout_engine(c["P"], c["H"])
    t0 = time.time()
    traj = eng.rollout(start, acts).clone()
    torch.cuda.synchronize()
    print(prec, "done in", round(time.time() - t0, 3), "s; tc_status", eng.tc_status(), "finite", bool(torch.isfinite(traj).all()), flush=True)
    out[prec] = traj.cpu().numpy()
print("tc vs fp32 rel_err:", rel_err(out["tc"], out["fp32"]))
for h in range(1, c["H"] + 1):
    print(" h", h, "err", rel_err(out["tc"][h], out["fp32"][h]))
if not P_override and not H_override:
    g = np.load(os.path.join(GOLD, name + ".npz"))
    nxt = np.transpose(out["tc"][1:], (2, 1, 0, 3))
    print("tc vs reference fixture:", rel_err(nxt, g["next_observations"]))
a, b = out["tc"].astype(np.float64), out["fp32"].astype(np.float64)
print("rel-L2 all:", np.linalg.norm(a - b) / np.linalg.norm(b), " last step:", np.linalg.norm(a[-1] - b[-1]) / np.linalg.norm(b[-1]))
sd = spec.state_dim
print("rel-L2 state dims only:", np.linalg.norm(a[1:, ..., :sd] - b[1:, ..., :sd]) / np.linalg.norm(b[1:, ..., :sd]))
