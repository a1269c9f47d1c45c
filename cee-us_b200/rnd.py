"""Drop-in for ``TorchRNDMpcICem`` (yaml ``controller: mpc-rnd-icem-torch``, mbrl/controllers/mpc_torch_rnd.py:125-293): iCEM whose
intrinsic cost is the Random-Network-Distillation bonus of the predicted observations.

Sampling, rollout, top-k and refit run in libceeus_b200 exactly as for the other controllers; the RND networks (predictor /
target MLPs over the goal-free observation, 256 hidden / 128 output in curious-rnd-i-cem.yaml) are evaluated between the rollout
and the cost reduction, on the device views of the rollout, as two batched GEMM chains (torch on the CUDA device: ~0.1 GFLOP per
iteration, library code -- not a kernel worth owning).  The bonus enters the library through the same door as any task cost the
library does not know (``ceeus_plan_iter_costs(env_cost_dev)``), so ``sum`` / ``best`` and the ensemble reduction stay on the device.

Semantics kept from the reference:
  * ``RND.forward``: running-statistics normaliser of the observation, clamp to +-5, squared error between target and predictor,
    mean over the representation dims (:113-122);
  * ``compute_intr_reward``: the running mean / variance ``RMS`` of the prediction errors is UPDATED with every batch it scores
    (:26-47, :219-224) and the bonus is ``error / (sqrt(var) + 1e-8)``;
  * ``trajectory_cost_fn``: ``-bonus (+ extrinsic_reward_scale * env.cost_fn)`` (:253-275);
  * ``train_rnd`` (post-rollout hook): normaliser update from the latest rollouts, then Adam on the predictor (:186-217);
  * ``save`` / ``load`` checkpoint keys (:277-293).
The reference pairs this controller with its single-model forward models; here every member of the ensemble model scores its own
prediction ([P,E,H] bonuses, ensemble mean of the trajectory costs) -- with one member that is the reference's computation.
"""
import numpy as np
import torch
import torch.nn as nn

from .controllers import B200MpcICem
from .rollout_buffer import RolloutView


def _build_mlp(input_dim, output_dim, size, num_layers):
    """mbrl/models/utils.py:54-79 with activation relu, no LayerNorm."""
    dims = [input_dim] + [size] * num_layers
    layers = []
    for a, b in zip(dims[:-1], dims[1:]):
        layers += [nn.Linear(a, b), nn.ReLU()]
    return nn.Sequential(*layers, nn.Linear(dims[-1], output_dim))


class RMS:
    """running mean and variance (mpc_torch_rnd.py:26-47)"""

    def __init__(self, device, epsilon=1e-4, shape=(1,)):
        self.M = torch.zeros(shape, device=device)
        self.S = torch.ones(shape, device=device)
        self.n = epsilon

    def __call__(self, x):
        bs = x.size(0)
        delta = torch.mean(x, dim=0) - self.M
        new_M = self.M + delta * bs / (self.n + bs)
        new_S = (self.S * self.n + torch.var(x, dim=0) * bs + torch.square(delta) * self.n * bs / (self.n + bs)) / (self.n + bs)
        self.M, self.S = new_M, new_S
        self.n += bs
        return self.M, self.S


class RunningNormalizer:
    """torch_helpers.Normalizer (:804-883) for the RND observation input, statistics on the device."""

    def __init__(self, dim, device, eps=1e-6):
        self.eps, self.device = eps, device
        self.sum, self.sum_of_squares, self.count = np.zeros(dim), np.zeros(dim), 1.0
        self.mean, self.std = 0.0, 1.0
        self.mean_t, self.std_t = torch.zeros(1, device=device), torch.ones(1, device=device)

    def update(self, data):
        data = np.asarray(data, np.float64)
        self.sum += data.sum(0)
        self.sum_of_squares += np.square(data).sum(0)
        self.count += data.shape[0]
        self.mean = self.sum / self.count
        self.std = np.maximum(self.eps, np.sqrt(self.sum_of_squares / self.count - np.square(self.sum / self.count) + self.eps))
        self._sync()

    def _sync(self):
        self.mean_t = torch.as_tensor(np.asarray(self.mean), dtype=torch.float32, device=self.device)
        self.std_t = torch.as_tensor(np.asarray(self.std), dtype=torch.float32, device=self.device)

    def normalize(self, x):
        return (x - self.mean_t) / self.std_t

    def state_dict(self):
        return {"mean": self.mean, "std": self.std, "sum": self.sum, "sum_of_squares": self.sum_of_squares, "count": self.count}

    def load_state_dict(self, sd):
        self.mean, self.std, self.sum, self.sum_of_squares, self.count = (sd[k] for k in ("mean", "std", "sum", "sum_of_squares", "count"))
        self._sync()


class B200RndMpcICem(B200MpcICem):
    """yaml key ``mpc-rnd-icem-b200``."""

    _curiosity = False
    needs_training = True

    def __init__(self, *, model_params, train_params, extrinsic_reward=False, extrinsic_reward_scale=1.0, clip_val=5.0, **kwargs):
        self._w_extrinsic_reward = extrinsic_reward
        self._maybe_extrinsic_reward_scale = extrinsic_reward_scale
        self._rnd_model_params, self._rnd_train_params = dict(model_params), dict(train_params)
        self.clip_val = clip_val
        self.rnd = None
        super().__init__(**kwargs)
        env, dev = self.env, self._engine.device
        self.obs_dim = int(env.observation_space_size_preproc)
        mp = self._rnd_model_params
        if mp.get("rnd_network_type", "mlp") != "mlp":
            raise NotImplementedError
        gen = torch.Generator().manual_seed(int(self.backend_params.get("seed", 0)))
        self.predictor = _build_mlp(self.obs_dim, mp["rnd_rep_dim"], mp["rnd_hidden_dim"], mp["num_layers_predictor"])
        self.target = _build_mlp(self.obs_dim, mp["rnd_rep_dim"], mp["rnd_hidden_dim"], mp["num_layers_target"])
        for net in (self.predictor, self.target):  # weight_init: orthogonal weights, zero biases (:13-23)
            for m in net:
                if isinstance(m, nn.Linear):
                    nn.init.orthogonal_(m.weight.data, generator=gen)
                    m.bias.data.fill_(0.0)
            net.to(dev)
        for p in self.target.parameters():
            p.requires_grad = False
        self.obs_normalizer = RunningNormalizer(self.obs_dim, dev)
        self.intrinsic_reward_rms = RMS(device=dev)
        tp = self._rnd_train_params
        self.rnd_lr, self.batch_size = tp["learning_rate"], tp["batch_size"]
        n = tp.get("rnd_num_epochs_or_its", 20)
        self.epochs, self.iterations = (n, 0) if tp.get("rnd_epochs", True) else (0, n)
        self.rnd_opt = torch.optim.Adam(self.predictor.parameters(), lr=self.rnd_lr)

    def preallocate_memory(self):
        # the bonus is computed outside the library: the planner runs in "cost evaluated by the caller" mode
        self._extra_engine_kwargs = dict(external_cost=True)
        super().preallocate_memory()

    # -- RND ---------------------------------------------------------------------------------------------------------------
    def rnd_forward(self, obs):
        x = torch.clamp(self.obs_normalizer.normalize(obs), -self.clip_val, self.clip_val)
        return torch.square(self.target(x).detach() - self.predictor(x)).mean(dim=-1, keepdim=True)

    @torch.no_grad()
    def compute_intr_reward(self, obs):
        err = self.rnd_forward(obs[..., :self.obs_dim])
        _, var = self.intrinsic_reward_rms(err)
        return err / (torch.sqrt(var) + 1e-8)

    def _cost_from_rollout(self, view):
        """[P,E,H,*] views of the device rollout -> [P,E,H] costs (mpc_torch_rnd.py:240-275)."""
        nxt = view["next_observations"]
        P, E, H, O = nxt.shape
        bonus = self.compute_intr_reward(nxt.reshape(-1, O)).view(P, E, H)
        cost = -bonus
        if self._w_extrinsic_reward:
            cost = cost + self._maybe_extrinsic_reward_scale * self.env.cost_fn(view["observations"], view["actions"], nxt)
        return cost

    # -- training hook (post_rollout_hooks/train_rnd_module.py -> train_rnd) ---------------------------------------------------
    def update_rnd(self, obs):
        loss = self.rnd_forward(obs).mean()
        self.rnd_opt.zero_grad(set_to_none=True)
        loss.backward()
        self.rnd_opt.step()
        return {"rnd_loss": loss.item()}

    def train_rnd(self, rollout_buffer):
        latest = np.asarray(rollout_buffer.latest_rollouts["observations"], np.float32)[:, :self.obs_dim]
        self.obs_normalizer.update(latest)
        data = torch.as_tensor(np.asarray(rollout_buffer["observations"], np.float32)[:, :self.obs_dim], device=self._engine.device)
        n = data.shape[0]
        order = np.arange(n)
        if self.epochs:   # TrainingIterator.get_epoch_iterator (torch_helpers.py:520-535): the array is re-shuffled IN PLACE every epoch
            epoch_length = int(np.ceil(n / self.batch_size))

            def gen():
                for _ in range(self.epochs):
                    np.random.shuffle(order)
                    for j in range(epoch_length):
                        yield order[j * self.batch_size:(j + 1) * self.batch_size].copy()
            batches = gen()
        else:             # get_basic_iterator (:537-551): batch_size random rows per iteration
            epoch_length = 1
            batches = (np.random.randint(0, n, self.batch_size) for _ in range(self.iterations))
        accum, rnd_loss = 0.0, 0.0
        for i, idx in enumerate(batches):
            accum += self.update_rnd(data[torch.as_tensor(idx, device=data.device)])["rnd_loss"]
            if (i + 1) % epoch_length == 0 or i == 0:
                rnd_loss = accum / min(epoch_length, i + 1)
                accum = 0.0
        return {"epoch_rnd_loss": rnd_loss}

    def train(self):
        pass

    def save(self, path):
        torch.save({"rnd_target": self.target.state_dict(), "rnd_predictor": self.predictor.state_dict(),
                    "obs_normalizer": self.obs_normalizer.state_dict(), "optimizer": self.rnd_opt.state_dict()}, path)

    def load(self, path):
        sd = torch.load(path, map_location=self._engine.device, weights_only=False)
        self.target.load_state_dict(sd["rnd_target"])
        self.predictor.load_state_dict(sd["rnd_predictor"])
        self.obs_normalizer.load_state_dict(sd["obs_normalizer"])
        self.rnd_opt.load_state_dict(sd["optimizer"])
