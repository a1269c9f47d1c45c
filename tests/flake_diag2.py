"""Manual diagnostic: one fp32 rollout + oracle per process; prints digests so that run-to-run variation can be attributed."""
import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import conftest  # noqa
import test_gpu_parity as T
from helpers import rel_err
env_kind, n_obj, model, hidden, P, H = ("construction", 2, "gnn", 128, 300, 12)
c = dict(kind="rollout", env=env_kind, n_obj=n_obj, model=model, hidden=hidden, layers=2, P=P, H=H, wseed=1000 + n_obj + hidden, identity_norm=False, running=False, oseed=77 + P)
spec, weights, norms, obs, goal = T.build_inputs(c)
env = T.product_env(c, goal)
om = T.oracle_model(c, spec, weights, norms)
acts = T.rollout_actions(c, spec)
rs = np.random.RandomState(5)
starts = np.tile(obs, (P, 1)); starts[:, :spec.state_dim] += (0.05 * rs.standard_normal((P, spec.state_dim))).astype(np.float32)
md5 = lambda a: hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:10]
wd = md5(np.concatenate([np.asarray(v, np.float32).ravel() for v in weights.values()]))
want = T.orc.rollout(om, torch.from_numpy(starts), torch.from_numpy(acts), torch.from_numpy(goal)).numpy()
pm = T.product_model(c, env, weights, norms)
pm.predict_n_steps(start_observations=torch.from_numpy(starts).to("cuda"), start_states=None, policy=T._Policy(torch.from_numpy(acts).to("cuda")), horizon=H)
got = pm._rollout_engine(P, H).traj_.cpu().numpy()
print("weights", wd, "acts", md5(acts), "starts", md5(starts), "want", md5(want), "got", md5(got), "err %.2e" % rel_err(got, want), "threads", torch.get_num_threads(), flush=True)
