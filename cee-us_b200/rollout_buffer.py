"""Container with the ``SimpleRolloutBuffer`` protocol (mbrl/rolloutbuffer.py:120-154): a dict of tensors keyed
``observations / next_observations / actions`` shaped ``[P, E, H, *]``.  Here the three entries are *views* of the
buffers the CUDA rollout wrote (``traj[H+1,E,P,O]`` and ``actions[P,H,U]``): nothing is copied."""
from collections import abc

import torch


class RolloutView(abc.Sequence):
    def __init__(self, rollouts=None, **buffer):
        self.buffer = buffer if rollouts is None else rollouts

    @classmethod
    def from_device_buffers(cls, traj, actions):
        """traj [H+1,E,P,O], actions [P,H,U] -> views shaped like the reference's permuted buffers
        (gnn_ensemble.py:973-981)."""
        E = traj.shape[1]
        return cls(observations=traj[:-1].permute(2, 1, 0, 3), next_observations=traj[1:].permute(2, 1, 0, 3),
                   actions=actions.unsqueeze(1).expand(-1, E, -1, -1))

    def as_array(self, key):
        return self.buffer[key]

    def extend(self, other):
        self.buffer = {key: torch.cat([self[key], other[key]], 0) for key in self.buffer}

    def __getitem__(self, item):
        if isinstance(item, str):
            return self.buffer[item]
        return {key: value[item] for key, value in self.buffer.items()}

    def __len__(self):
        return list(self.buffer.values())[0].shape[0]
