"""Case table + seeded input builders shared by the golden generator (make_golden.py) and the tests.

TEST INFRASTRUCTURE.  Imports only the oracle module (no reference code), so it works on the GPU box."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if os.path.join(ROOT, "oracle") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))

import icem_oracle as orc  # noqa: E402


class SeededNoise:
    """Noise provider used on both sides of every parity test: standard normal draws from a numpy
    RandomState pushed through the float64 irDFT matrix of icem_oracle (-> realistic colored noise)."""

    def __init__(self, seed, beta):
        self.rs = np.random.RandomState(seed)
        self.beta = beta
        self._M = {}

    def __call__(self, size):
        n, U, H = size
        F = H // 2 + 1
        if H not in self._M:
            self._M[H] = orc.irdft_matrix(self.beta, H).astype(np.float64)
        z = self.rs.standard_normal((n, U, 2 * F))
        return torch.from_numpy((z @ self._M[H]).astype(np.float32))


# ---- case table ---------------------------------------------------------------------------------
# Every case is fully described by this dict; tests rebuild the inputs from it.
CASES = {
    # rollouts + costs
    "rollout_gnn_c2": dict(kind="rollout", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=24, H=6,
                           wseed=11, identity_norm=True, running=False, oseed=3),
    "rollout_gnn_c6_stack": dict(kind="rollout", env="construction", n_obj=6, model="gnn", hidden=128, layers=2, P=7,
                                 H=5, wseed=12, identity_norm=False, running=False, oseed=4),
    "rollout_gnn_pg4": dict(kind="rollout", env="playground", n_obj=4, model="gnn", hidden=64, layers=2, P=12, H=5,
                            wseed=13, identity_norm=False, running=True, oseed=5),
    "rollout_mlp_c4": dict(kind="rollout", env="construction", n_obj=4, model="mlp", hidden=256, layers=3, P=10, H=5,
                           wseed=14, identity_norm=False, running=False, oseed=6),
    "rollout_gnn_c3_pp_sparse": dict(kind="rollout", env="construction", n_obj=3, model="gnn", hidden=128, layers=2,
                                     P=9, H=4, wseed=15, identity_norm=True, running=False, oseed=7,
                                     case="PickAndPlace", sparse=False),
    "rollout_gnn_c2_flip": dict(kind="rollout", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=16, H=5,
                                wseed=16, identity_norm=True, running=False, oseed=13, case="Flip", sparse=False),
    "rollout_gnn_c2_flip_sparse": dict(kind="rollout", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=16,
                                       H=5, wseed=16, identity_norm=True, running=False, oseed=13, case="Flip", sparse=True),
    "rollout_gnn_c3_slide": dict(kind="rollout", env="construction", n_obj=3, model="gnn", hidden=128, layers=2, P=12, H=5,
                                 wseed=17, identity_norm=True, running=False, oseed=14, case="Slide", sparse=False,
                                 shaped=False),
    # the other GNN variants reachable through the same yaml: ignore_agent_object: false (edge_mlp_global),
    # and the homogeneous GNN of gnn_ensemble_homogeneous.yaml (agent_as_global_node: false) with 1 and 2 message-passing rounds
    "rollout_gnn_c3_edge_global": dict(kind="rollout", env="construction", n_obj=3, model="gnn", hidden=128, layers=2, P=14, H=5,
                                       wseed=31, identity_norm=False, running=False, oseed=21, variant="edge_global"),
    "rollout_gnn_pg4_edge_global": dict(kind="rollout", env="playground", n_obj=4, model="gnn", hidden=64, layers=2, P=12, H=4,
                                        wseed=32, identity_norm=False, running=True, oseed=22, variant="edge_global"),
    "rollout_gnn_c2_homog": dict(kind="rollout", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=16, H=5,
                                 wseed=33, identity_norm=False, running=False, oseed=23, variant="homog", emb=32, n_mp=1),
    "rollout_gnn_c4_homog_mp2": dict(kind="rollout", env="construction", n_obj=4, model="gnn", hidden=128, layers=2, P=10, H=4,
                                     wseed=34, identity_norm=False, running=False, oseed=24, variant="homog", emb=32, n_mp=2),
    "rollout_gnn_pg3_homog": dict(kind="rollout", env="playground", n_obj=3, model="gnn", hidden=64, layers=2, P=12, H=4,
                                  wseed=35, identity_norm=False, running=True, oseed=25, variant="homog", emb=16, n_mp=1),
    # task costs on synthetic trajectories close to the goal (a random-init model never gets there): every comparison of
    # fpp_construction_env.py:404-511 falls on both sides -- solved / unsolved blocks, the gripper mask, the 0.5 term of the
    # sparse branch, the buffer of PickAndPlace
    "cost_c3_tower_sparse": dict(kind="cost", env="construction", n_obj=3, model="gnn", hidden=128, layers=2, P=24, H=6,
                                 wseed=18, identity_norm=True, running=False, oseed=15, case="Singletower", sparse=True),
    "cost_c3_tower_dense": dict(kind="cost", env="construction", n_obj=3, model="gnn", hidden=128, layers=2, P=24, H=6,
                                wseed=18, identity_norm=True, running=False, oseed=15, case="Singletower", sparse=False),
    "cost_c4_pp_dense": dict(kind="cost", env="construction", n_obj=4, model="gnn", hidden=128, layers=2, P=24, H=6,
                             wseed=18, identity_norm=True, running=False, oseed=19, case="PickAndPlace", sparse=False),
    "cost_c2_pyramid_unshaped": dict(kind="cost", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=24, H=6,
                                     wseed=18, identity_norm=True, running=False, oseed=20, case="Pyramid", sparse=False,
                                     shaped=False),
    # RoboDesk task costs (robodesk_env.py:53-135) on synthetic observations around every threshold, dense and sparse
    **{f"cost_robodesk_{t}_{r}": dict(kind="cost", env="robodesk", task=t, reward=r, n_obj=8, model="gnn", hidden=128, layers=2, P=16,
                                      H=5, wseed=40, identity_norm=True, running=False, oseed=30 + i)
       for i, t in enumerate(("open_slide", "open_drawer", "push_red", "push_green", "push_blue", "ball_off_table",
                              "upright_block_off_table", "flat_block_off_table")) for r in ("dense", "sparse")},
    # full MPC steps (3 consecutive get_action calls after beginning_of_rollout)
    "mpc_gnn_c2_curious": dict(kind="mpc", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=128, H=20,
                               wseed=21, identity_norm=True, running=False, oseed=8, nseed=100, steps=3,
                               curiosity=True, cost="sum", sampler={}),  # BASELINE config 1
    "mpc_gnn_c6_stack_extr": dict(kind="mpc", env="construction", n_obj=6, model="gnn", hidden=128, layers=2, P=48,
                                  H=8, wseed=22, identity_norm=False, running=False, oseed=9, nseed=101, steps=3,
                                  curiosity=True, extrinsic=True, cost="best",
                                  sampler=dict(init_std=0.5, use_mean_actions=False)),  # config 3 style
    "mpc_gnn_c2_task": dict(kind="mpc", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=64, H=10,
                            wseed=23, identity_norm=True, running=False, oseed=10, nseed=102, steps=2,
                            curiosity=False, cost="best", sampler=dict(init_std=0.5, use_mean_actions=False)),
    "mpc_gnn_pg4_curious": dict(kind="mpc", env="playground", n_obj=4, model="gnn", hidden=64, layers=2, P=64, H=10,
                                wseed=24, identity_norm=False, running=True, oseed=11, nseed=103, steps=3,
                                curiosity=True, cost="sum",
                                sampler=dict(keep_previous_elites=False, shift_elites_over_time=False,
                                             fraction_elites_reused=0.1)),  # config 4 hyper-parameters
    # branches of the sampler / cost reduction no shipped yaml switches on (mpc_torch.py:189-202, 327-332, 392-406)
    "mpc_gnn_c3_task_coststd": dict(kind="mpc", env="construction", n_obj=3, model="gnn", hidden=128, layers=2, P=64, H=8,
                                    wseed=26, identity_norm=True, running=False, oseed=16, nseed=105, steps=3,
                                    curiosity=False, cost="sum",
                                    sampler=dict(init_std=0.5, use_ensemble_cost_std=True, execute_best_elite=False)),
    "mpc_gnn_c2_relinit_bounds": dict(kind="mpc", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=64, H=8,
                                      wseed=27, identity_norm=True, running=False, oseed=17, nseed=106, steps=2,
                                      curiosity=True, cost="sum", sampler={},
                                      abounds=([-1.0, -0.5, -0.8, -0.2], [0.6, 1.0, 0.4, 1.0])),
    "mpc_gnn_c2_absinit_bounds": dict(kind="mpc", env="construction", n_obj=2, model="gnn", hidden=128, layers=2, P=64, H=8,
                                      wseed=27, identity_norm=True, running=False, oseed=17, nseed=106, steps=2,
                                      curiosity=True, cost="sum", sampler=dict(relative_init=False),
                                      abounds=([-1.0, -0.5, -0.8, -0.2], [0.6, 1.0, 0.4, 1.0])),
    "mpc_mlp_c4_curious": dict(kind="mpc", env="construction", n_obj=4, model="mlp", hidden=256, layers=3, P=64, H=10,
                               wseed=25, identity_norm=False, running=False, oseed=12, nseed=104, steps=2,
                               curiosity=True, cost="sum", sampler={}),  # config 5 style
}


def build_inputs(c):
    """Everything derivable from the case dict (used by the generator *and* by the tests)."""
    if c["env"] == "construction":
        spec = orc.construction_spec(c["n_obj"], case=c.get("case", "Singletower"), sparse=c.get("sparse", False),
                                     shaped_reward=c.get("shaped", True))
    elif c["env"] == "robodesk":
        spec = orc.robodesk_spec(c["task"], c["reward"])
    else:
        spec = orc.playground_spec(c["n_obj"])
    E = 5
    if c["model"] == "gnn":
        if c.get("variant") == "homog":
            shapes = orc.homog_weight_shapes(spec, c["hidden"], c["layers"], E, c["emb"])
        else:
            shapes = orc.gnn_weight_shapes(spec, c["hidden"], c["layers"], E, edge_global=c.get("variant") == "edge_global")
        ndims = orc.gnn_normalizer_dims(spec)
    else:
        shapes = orc.mlp_weight_shapes(spec, c["hidden"], c["layers"], E)
        ndims = orc.mlp_normalizer_dims(spec)
    weights = orc.synth_weights(shapes, c["wseed"])
    norms = orc.synth_normalizers(ndims, c["wseed"], identity=c["identity_norm"])
    if "abounds" in c:  # asymmetric action bounds: relative_init then differs from absolute init (mpc_torch.py:189-202)
        spec.action_low = np.asarray(c["abounds"][0], np.float32)
        spec.action_high = np.asarray(c["abounds"][1], np.float32)
    obs, goal = orc.synth_start_obs(spec, c["oseed"])
    if c.get("case") in ("Flip", "Slide"):
        # per-coordinate tasks: goals within a few thresholds of the achieved values, so that both sides of every
        # comparison in the cost occur (Flip compares euler angles, dims 3..5 of a block; Slide positions)
        rs = np.random.RandomState(c["oseed"] + 500)
        ao = 3 if c["case"] == "Flip" else 0
        for i in range(c["n_obj"]):
            ach = obs[spec.agent_dim + spec.dyn_dim * i + ao:spec.agent_dim + spec.dyn_dim * i + ao + 3]
            goal[3 * i:3 * i + 3] = ach + rs.uniform(-0.25, 0.25, 3).astype(np.float32)
        obs[spec.state_dim:] = goal
    return spec, weights, norms, obs, goal


def rollout_actions(c, spec):
    rs = np.random.RandomState(c["oseed"] + 1000)
    return np.clip(0.8 * rs.standard_normal((c["P"], c["H"], spec.action_dim)), -1, 1).astype(np.float32)


def cost_trajectories(c, spec, obs, goal):
    """kind == "cost": synthetic trajectories [H+1,E,P,O] around the goal configuration.  traj[0] = the start observation
    with block 0 already on its goal (-> next block id k* = 1, fpp_construction_env.py:365-383)."""
    rs = np.random.RandomState(c["oseed"] + 900)
    E, P, H, O = 5, c["P"], c["H"], spec.obs_dim
    if spec.kind == "robodesk":
        # every feature the rewards read scattered around its threshold (slide x 0.25, drawer y 0.65, button z 0.7595 / 0.7585,
        # block z 0.6) and the gripper somewhere near
        traj = (0.3 * rs.standard_normal((H + 1, E, P, O)) + 0.5).astype(np.float32)
        traj[0] = obs
        traj[1:, ..., 37] = 0.25 + rs.uniform(-0.1, 0.1, (H, E, P))
        traj[1:, ..., 25] = 0.65 + rs.uniform(-0.1, 0.1, (H, E, P))
        for zi in (52, 65, 78):
            traj[1:, ..., zi] = 0.759 + rs.uniform(-0.002, 0.002, (H, E, P))
        for bx in (89, 102, 115):
            traj[1:, ..., bx + 2] = 0.6 + rs.uniform(-0.2, 0.25, (H, E, P))
        return traj.astype(np.float32)
    A, D, N, sd = spec.agent_dim, spec.dyn_dim, spec.n_obj, spec.state_dim
    start = obs.copy()
    start[A:A + 3] = goal[0:3] + np.float32(0.004)
    start[sd:] = goal
    traj = np.empty((H + 1, E, P, O), np.float32)
    traj[0] = start
    nxt = np.tile(start, (H, E, P, 1))
    for i in range(N):  # block positions within +-2.5 thresholds of their goals, some exactly solved
        noise = rs.uniform(-0.05, 0.05, (H, E, P, 3)) * rs.choice([0.0, 0.2, 1.0], (H, E, P, 1))
        nxt[..., A + D * i:A + D * i + 3] = goal[3 * i:3 * i + 3] + noise
    target = rs.randint(0, N + 1, (H, E, P))  # the gripper hovers near one of the blocks / its own goal
    centres = np.concatenate([nxt[..., A + D * i:A + D * i + 3][..., None, :] for i in range(N)] +
                             [np.broadcast_to(goal[3 * N:3 * N + 3], (H, E, P, 1, 3))], axis=-2)
    pick = np.take_along_axis(centres, target[..., None, None].repeat(3, -1), axis=-2)[..., 0, :]
    nxt[..., 0:3] = pick + rs.uniform(-0.03, 0.03, (H, E, P, 3)) * rs.choice([0.3, 1.0], (H, E, P, 1))
    traj[1:] = nxt.astype(np.float32)
    traj[1:, ..., sd:] = goal
    return traj
