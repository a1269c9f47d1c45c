"""Lock-step batched counterpart of ``RolloutManager`` (mbrl/rollout_utils.py:26-420; SURVEY.md 8f rank 4 -- the caller that feeds
BASELINE config 4, "64 parallel environments x 1024 trajectories").

The reference runs ``num_parallel`` rollouts in worker PROCESSES, each with its own environment and its own copy of the controller
(``par_sample``, rollout_utils.py:183-216): N planners, N times the model evaluations.  Here the N environments are stepped in lock
step by this process and ONE ``policy.get_action(obs[N, O])`` call per time step plans for all of them -- the controller is built
with ``backend_params = {"num_envs": N}`` and the N planners share every kernel launch (one CUDA-graph launch per time step).  The
environments themselves stay whatever the caller has (MuJoCo, a learned model, a test double): only ``reset_with_mode / step / goal``
and, optionally, ``set_state_from_observation / set_GT_state / get_GT_state / is_success`` are used, exactly as ``_sample`` uses them.

Per rollout the semantics of ``RolloutManager._sample`` (rollout_utils.py:218-384) are kept: ``beginning_of_rollout`` before the first
step, one transition ``(ob, next_ob, ac, rew, done, state[, success][, info values])`` per step, ``only_final_reward``, the pre / post
env-step hooks (called per environment with ``env, ob, ac, t, index`` in their ``locals``), a rollout ends at its own ``done``,
``end_of_rollout`` after the last step.  What is not carried over: video recording, the real-robot timing branch and per-step logging.
"""
import numpy as np


class Rollout:
    """One rollout as a numpy structured array with the reference's field layout (mbrl/rolloutbuffer.py:11-60: one record per
    transition, every field stored as float64 with the shape of its first entry); ``rollout["rewards"]``, ``len(rollout)``, iteration
    and ``field_names()`` behave as there, so the reference's ``RolloutBuffer`` accepts it."""

    def __init__(self, field_names, transitions):
        transitions = list(transitions)
        if not transitions:
            raise ValueError("No data given!")
        self.dtype = [(name, "f8", np.array(item).shape) for name, item in zip(field_names, transitions[0])]
        self._data = np.array(transitions, dtype=self.dtype)

    def field_names(self):
        return [d[0] for d in self.dtype]

    def __len__(self):
        return len(self._data)

    def __iter__(self):
        return iter(self._data)

    def __getitem__(self, item):
        return self._data[item]

    def __repr__(self):
        return repr(self._data)


class BatchedEnvView:
    """What a batched controller sees as its ``env``: every attribute of the first environment (dimensions, cost parameters --
    ``describe_env`` reads them) and the GOALS OF ALL environments, stacked ``[N, G]``, re-read at every ``get_action``
    (``env.goal`` changes with every ``env.reset``)."""

    def __init__(self, envs):
        self._envs = list(envs)

    @property
    def goal(self):
        return np.stack([np.asarray(e.goal, np.float32).reshape(-1) for e in self._envs])

    def __getattr__(self, name):
        return getattr(self._envs[0], name)


def _param(params, key, default):
    if isinstance(params, dict):
        return params.get(key, default)
    return getattr(params, key, default)


class B200RolloutManager:
    """``roll_params``: the yaml ``rollout_params`` (``task_horizon``, ``use_env_states``, ``only_final_reward``,
    ``pre_env_step_hooks``, ``post_env_step_hooks``; ``num_parallel`` is ``len(envs)``)."""

    valid_modes = ["train", "evaluate"]

    def __init__(self, envs, roll_params):
        self.envs = list(envs)
        if not self.envs:
            raise ValueError("at least one environment")
        self.num_parallel = len(self.envs)
        self.task_horizon = int(_param(roll_params, "task_horizon", 0))
        if self.task_horizon < 1:
            raise ValueError("rollout_params.task_horizon must be >= 1")
        self.only_final_reward = bool(_param(roll_params, "only_final_reward", False))
        self.use_env_states = bool(_param(roll_params, "use_env_states", False))
        self.pre_env_step_hooks = list(_param(roll_params, "pre_env_step_hooks", []) or [])
        self.post_env_step_hooks = list(_param(roll_params, "post_env_step_hooks", []) or [])
        self.calls_counter = 0

    def controller_env(self):
        """The ``env`` to construct the batched controller / model with."""
        return BatchedEnvView(self.envs)

    def reset(self):
        self.calls_counter = 0

    def _env_state(self, env):
        # rollout_utils.py:386-391
        if self.use_env_states and hasattr(env, "get_GT_state"):
            return env.get_GT_state()
        return None

    def _start(self, env, mode, start_ob, start_state):
        # rollout_utils.py:237-246
        if start_ob is not None and hasattr(env, "set_state_from_observation"):
            if start_state is None:
                env.set_state_from_observation(start_ob)
            else:
                env.set_GT_state(start_state)
            return np.asarray(start_ob)
        return np.asarray(env.reset_with_mode(mode))

    def sample(self, policy, render=False, mode="train", name="", start_ob=None, start_state=None, no_rollouts=None,
               use_tqdm=False, desc="rollout_num"):
        """``no_rollouts`` rollouts (default: one per environment), ``num_parallel`` at a time; returns them as a list in order."""
        if mode not in self.valid_modes:
            raise ValueError(f"mode must be one of {self.valid_modes}")
        if render:
            raise NotImplementedError("video recording is not part of the batched manager")
        n_total = self.num_parallel if no_rollouts is None else int(no_rollouts)
        starts_ob = [None] * n_total if start_ob is None else list(start_ob)
        starts_state = [None] * n_total if start_state is None else list(start_state)
        rollouts = []
        for first in range(0, n_total, self.num_parallel):
            idx = list(range(first, min(first + self.num_parallel, n_total)))
            rollouts += self._sample_wave(policy, mode, [starts_ob[i] for i in idx], [starts_state[i] for i in idx])
        return rollouts

    def _sample_wave(self, policy, mode, starts_ob, starts_state):
        N, n_live = self.num_parallel, len(starts_ob)
        # a ragged last wave still plans for N environments: the spare slots replay environment 0's observations and their
        # actions are dropped
        obs = [self._start(self.envs[i], mode, starts_ob[i], starts_state[i]) for i in range(n_live)]
        batch = lambda rows: np.stack([rows[i] if i < n_live else rows[0] for i in range(N)])  # noqa: E731
        if getattr(policy, "has_state", True):
            states = [self._env_state(self.envs[i]) for i in range(n_live)]
            policy.beginning_of_rollout(observation=batch(obs), state=states if self.use_env_states else None, mode=mode)
        transitions = [[] for _ in range(n_live)]
        fields = [None] * n_live
        finished = [False] * n_live
        returns = [0.0] * n_live
        steps = 0
        for t in range(self.task_horizon):
            if all(finished):
                break
            states = [self._env_state(self.envs[i]) for i in range(n_live)]
            acs = np.asarray(policy.get_action(batch(obs), state=states if self.use_env_states else None, mode=mode))
            steps += 1
            for i in range(n_live):
                if finished[i]:
                    continue
                env, ob, ac, index = self.envs[i], obs[i], acs[i], i
                for hook in self.pre_env_step_hooks:
                    hook(locals(), globals())
                next_ob, rew, done, info_dict = env.step(ac)
                for hook in self.post_env_step_hooks:
                    hook(locals(), globals())
                if self.only_final_reward and t < self.task_horizon - 1:
                    rew = 0
                # value order and field order exactly as rollout_utils.py:303-311 / 357-366: the success flag is appended BEFORE the
                # info values, its NAME after the info keys (with a non-empty info_dict the reference's names are therefore shifted
                # by one -- reproduced, not repaired: consumers index these records the reference's way)
                transition = [ob, next_ob, ac, rew, done, states[i]]
                names = ["observations", "next_observations", "actions", "rewards", "dones", "env_states"]
                if hasattr(env, "is_success"):
                    transition.append(env.is_success(np.asarray(ob)[None, :], np.asarray(ac)[None, :], np.asarray(next_ob)[None, :])[0])
                if info_dict and len(info_dict) > 0:
                    transition += list(info_dict.values())
                    names += list(info_dict.keys())
                if hasattr(env, "is_success"):
                    names.append("successes")
                transitions[i].append(tuple(transition))
                fields[i] = names
                obs[i] = np.asarray(next_ob)
                returns[i] += rew
                finished[i] = bool(done)
        if getattr(policy, "has_state", True):
            policy.end_of_rollout(total_time=steps, total_return=float(np.sum(returns)), mode=mode)
        self.calls_counter += 1
        return [Rollout(field_names=tuple(fields[i]), transitions=transitions[i]) for i in range(n_live)]
