"""GPU: the lock-step batched rollout manager driving the batched planner (BASELINE config-4 semantics: N parallel environments, one
planner handle with num_envs = N, one CUDA-graph launch per time step for all of them) produces, environment by environment, the
rollouts that N independent single-environment controllers produce -- bit for bit, in both precision modes."""
import numpy as np
import pytest
import torch

from cases import CASES
from helpers import SeededNoise, build_inputs, product_controller, product_env, product_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _toy_env(c, spec, goal, start_obs):
    """A TensorEnv (the planner reads its dimensions / cost parameters) with a deterministic toy simulator attached: the agent
    moves by 0.05 * action, everything else stays."""
    env = product_env(c, goal)
    sd, U = spec.state_dim, spec.action_dim

    def reset_with_mode(mode):
        env._x = np.asarray(start_obs[:sd], np.float64).copy()
        return np.concatenate([env._x, np.asarray(env.goal, np.float64)])

    def step(ac):
        env._x[:U] += 0.05 * np.asarray(ac, np.float64)
        ob = np.concatenate([env._x, np.asarray(env.goal, np.float64)])
        return ob, -float(np.linalg.norm(env._x[:2])), False, {}

    env.reset_with_mode, env.step = reset_with_mode, step
    return env


class _NoisePolicy:
    """Hands the controller the step's noise tensor (parity hook of get_action) -- the manager itself only knows the Controller API."""

    has_state = True

    def __init__(self, ctrl, noise_per_step):
        self.ctrl, self.noise, self.t = ctrl, noise_per_step, 0

    def beginning_of_rollout(self, **kw):
        self.ctrl.beginning_of_rollout(**kw)

    def end_of_rollout(self, **kw):
        self.ctrl.end_of_rollout(**kw)

    def get_action(self, obs, state=None, mode="train"):
        self.t += 1
        return self.ctrl.get_action(obs, state, mode, noise=self.noise[self.t - 1])


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_batched_manager_equals_independent_single_env_rollouts(precision):
    import ceeus_b200

    c = dict(CASES["mpc_gnn_pg4_curious"])
    c["P"], c["H"] = 32, 6
    N, T = 3, 3
    spec, weights, norms, obs, goal = build_inputs(c)
    rs = np.random.RandomState(1)
    starts = [obs + np.concatenate([0.1 * rs.standard_normal(spec.state_dim), np.zeros(spec.goal_dim)]).astype(np.float32) for _ in range(N)]
    goals = [goal + np.float32(0.1 * i) for i in range(N)]
    P, U, H = c["P"], spec.action_dim, c["H"]
    noise = SeededNoise(7, 3.5)
    per = [[[noise((P, U, H)) for _ in range(3)] for _ in range(N)] for _ in range(T)]  # [step][env][iter]

    envs = [_toy_env(c, spec, goals[i], starts[i]) for i in range(N)]
    mgr = ceeus_b200.B200RolloutManager(envs, dict(task_horizon=T))
    view = mgr.controller_env()
    ctrl = product_controller(c, view, product_model(c, view, weights, norms, precision=precision), precision=precision, num_envs=N)
    flat = [torch.cat([torch.cat([per[t][e][it] for e in range(N)]).reshape(-1) for it in range(3)]).to(DEV) for t in range(T)]
    got = mgr.sample(_NoisePolicy(ctrl, flat))
    assert len(got) == N and all(len(r) == T for r in got)
    assert ctrl.engine.kernel_launches() > 0

    for i in range(N):
        env_i = _toy_env(c, spec, goals[i], starts[i])
        single = ceeus_b200.B200RolloutManager([env_i], dict(task_horizon=T))
        ctl = product_controller(c, single.controller_env(), product_model(c, single.controller_env(), weights, norms, precision=precision),
                                 precision=precision)
        flat_i = [torch.cat([per[t][i][it].reshape(-1) for it in range(3)]).to(DEV) for t in range(T)]
        want = single.sample(_NoisePolicy(ctl, flat_i))[0]
        for name in ("observations", "next_observations", "actions", "rewards", "dones"):
            np.testing.assert_array_equal(got[i][name], want[name])
        assert np.abs(got[i]["actions"]).max() <= 1.0 and np.abs(got[i]["actions"]).sum() > 0
