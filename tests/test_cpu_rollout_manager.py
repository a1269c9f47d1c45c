"""CPU: the lock-step batched rollout manager (cee-us_b200/rollout_manager.py) against the reference's own ``RolloutManager._sample``
(mbrl/rollout_utils.py:218-384) on the same environments and the same policy: N rollouts collected in lock step with ONE batched
``get_action`` per time step carry exactly the transitions of N sequential reference rollouts."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_harness  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference package not available")


class ToyEnv:
    """Deterministic stand-in for a goal-space environment: linear dynamics, done when close to the goal, an info dict."""

    def __init__(self, seed, with_info=False, done_radius=0.0):
        rs = np.random.RandomState(seed)
        self.A = np.eye(4) + 0.05 * rs.standard_normal((4, 4))
        self.B = 0.3 * rs.standard_normal((4, 2))
        self.start = rs.standard_normal(4)
        self.goal = rs.standard_normal(2).astype(np.float32)
        self.with_info, self.done_radius = with_info, done_radius
        self.x = None

    def _obs(self):
        return np.concatenate([self.x, self.goal])

    def reset_with_mode(self, mode):
        self.x = self.start.copy() + (0.5 if mode == "evaluate" else 0.0)
        return self._obs()

    def step(self, ac):
        self.x = self.A @ self.x + self.B @ np.asarray(ac, np.float64)
        d = float(np.linalg.norm(self.x[:2] - self.goal))
        info = {"dist": d, "x0": float(self.x[0])} if self.with_info else {}
        return self._obs(), -d, d < self.done_radius, info


class ToyGoalEnv(ToyEnv):
    def is_success(self, ob, ac, next_ob):
        return np.linalg.norm(next_ob[:, :2] - next_ob[:, 4:6], axis=-1) < 1.0


class ToyPolicy:
    """Deterministic, stateful (counts its calls), accepts one observation [O] or a batch [N, O]."""

    has_state = True

    def __init__(self):
        self.calls, self.began, self.ended = 0, [], []

    def beginning_of_rollout(self, *, observation, state=None, mode="train"):
        self.began.append((np.asarray(observation).shape, mode))

    def end_of_rollout(self, total_time=None, total_return=None, mode="train"):
        self.ended.append((total_time, total_return))

    def get_action(self, obs, state=None, mode="train"):
        self.calls += 1
        obs = np.asarray(obs)
        W = np.array([[0.3, -0.2, 0.1, 0.0, 0.5, -0.5], [0.1, 0.4, -0.3, 0.2, -0.5, 0.5]])
        if obs.ndim == 1:
            return np.tanh(W @ obs)
        return np.stack([np.tanh(W @ o) for o in obs])  # row by row: the same arithmetic as the single-observation call


def _reference_rollout(env, horizon, mode, only_final_reward):
    ref_harness._prepare()
    from mbrl.rollout_utils import RolloutManager

    if isinstance(env, ToyGoalEnv):
        # the reference keys the success column on its own interface class: give the toy environment that base at run time
        from mbrl.environments.abstract_environments import GoalSpaceEnvironmentInterface

        env.__class__ = type("ToyGoalEnvRef", (ToyGoalEnv, GoalSpaceEnvironmentInterface), {})
    return RolloutManager._sample(env=env, policy=ToyPolicy(), logger=None, render=False, mode=mode, start_ob=None, start_state=None,
                                  logging=False, use_env_states=False, task_horizon=horizon, use_tqdm=True,  # (the reference cannot run with use_tqdm=False: rollout_utils.py:263)
                                  only_final_reward=only_final_reward)


@pytest.mark.parametrize("env_cls,with_info,done_radius,only_final,mode", [
    (ToyEnv, False, 0.0, False, "train"), (ToyEnv, True, 0.0, True, "evaluate"), (ToyGoalEnv, False, 0.0, False, "train"),
    (ToyEnv, False, 1.2, False, "train"), (ToyGoalEnv, True, 0.0, False, "train")])
def test_lockstep_rollouts_equal_sequential_reference_rollouts(env_cls, with_info, done_radius, only_final, mode):
    import ceeus_b200.rollout_manager as rm

    N, horizon = 4, 12
    envs = [env_cls(seed=10 + i, with_info=with_info, done_radius=done_radius) for i in range(N)]
    mgr = rm.B200RolloutManager(envs, dict(task_horizon=horizon, use_env_states=False, only_final_reward=only_final))
    policy = ToyPolicy()
    got = mgr.sample(policy, mode=mode)
    assert len(got) == N and policy.began == [((N, 6), mode)] and len(policy.ended) == 1
    lens = []
    for i in range(N):
        want = _reference_rollout(env_cls(seed=10 + i, with_info=with_info, done_radius=done_radius), horizon, mode, only_final)
        assert got[i].field_names() == want.field_names()
        assert len(got[i]) == len(want)
        lens.append(len(want))
        for name in want.field_names():
            np.testing.assert_array_equal(got[i][name], want._data[name])
    # ONE batched planner call per time step for all N environments (the reference: one per environment and step)
    assert policy.calls == max(lens)
    if done_radius > 0:
        assert min(lens) < horizon  # some rollouts did end early, each at its own step


def test_more_rollouts_than_environments_run_in_waves_and_view_stacks_goals():
    import ceeus_b200.rollout_manager as rm

    envs = [ToyEnv(seed=3 + i) for i in range(3)]
    mgr = rm.B200RolloutManager(envs, dict(task_horizon=5))
    view = mgr.controller_env()
    assert view.goal.shape == (3, 2) and np.allclose(view.goal[1], envs[1].goal) and view.with_info is False
    policy = ToyPolicy()
    got = mgr.sample(policy, no_rollouts=5)
    assert len(got) == 5 and [b[0] for b in policy.began] == [(3, 6), (3, 6)]  # the ragged second wave still plans for 3 slots
    again = mgr.sample(ToyPolicy(), no_rollouts=2)
    for a, b in zip(got[:2], again):
        np.testing.assert_array_equal(a["observations"], b["observations"])
    with pytest.raises(ValueError):
        rm.B200RolloutManager(envs, dict(task_horizon=0))
