"""Manual diagnostic: repeats one fp32 rollout in-process and reports elements that differ between repetitions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import conftest  # noqa
import test_gpu_parity as T
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
env_kind, n_obj, model, hidden, P, H = ("construction", 2, "gnn", 128, 300, 12)
c = dict(kind="rollout", env=env_kind, n_obj=n_obj, model=model, hidden=hidden, layers=2, P=P, H=H, wseed=1000 + n_obj + hidden, identity_norm=False, running=False, oseed=77 + P)
spec, weights, norms, obs, goal = T.build_inputs(c)
env = T.product_env(c, goal)
acts = T.rollout_actions(c, spec)
rs = np.random.RandomState(5)
starts = np.tile(obs, (P, 1)); starts[:, :spec.state_dim] += (0.05 * rs.standard_normal((P, spec.state_dim))).astype(np.float32)
ref = None; bad = 0
for rep in range(reps):
    if rep % 10 == 0:  # fresh engine (new allocations) every 10 repetitions
        pm = T.product_model(c, env, weights, norms, precision=prec)
    # churn the allocator / memory contents between repetitions
    junk = torch.full((int(np.random.randint(1, 64)) * 1024 * 256,), float("nan"), device="cuda"); del junk
    pm.predict_n_steps(start_observations=torch.from_numpy(starts).to("cuda"), start_states=None, policy=T._Policy(torch.from_numpy(acts).to("cuda")), horizon=H)
    got = pm._rollout_engine(P, H).traj_.cpu().numpy()
    if ref is None: ref = got.copy(); continue
    if not np.array_equal(got, ref):
        bad += 1
        d = np.argwhere(got != ref)
        print("rep", rep, "mismatches", len(d), "first", d[:6].tolist(), "h set", sorted(set(d[:, 0].tolist()))[:5], "e set", sorted(set(d[:, 1].tolist())), "p range", d[:, 2].min(), d[:, 2].max(), "cols", sorted(set(d[:, 3].tolist()))[:12], "max abs", float(np.abs(got - ref).max()), flush=True)
print("done", prec, "reps", reps, "bad", bad)
