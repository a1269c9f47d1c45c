/*
 * ceeus_b200.h -- C ABI of libceeus_b200.so: the B200 (sm_100a) implementation of the CEE-US iCEM planning
 * hot path (martius-lab/cee-us).  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * The reference has no native boundary for this path (it is pure Python + torch eager), so every entry point
 * below names the reference *Python* interface it stands behind (paths relative to the reference root):
 *
 *   ceeus_create / ceeus_destroy      TorchMpcICem.__init__ + preallocate_memory  mbrl/controllers/mpc_torch.py:20-110
 *                                     GNNForwardEnsembleModel.__init__            mbrl/models/gnn_ensemble.py:37-141
 *   ceeus_set_weights                 model.model.load_state_dict / state_dict()  mbrl/models/gnn_ensemble.py:793-808,
 *                                                                                 mbrl/models/mlp_ensemble.py:625-638
 *   ceeus_set_normalizers             Normalizer_v2 / Normalizer (mean, std)      mbrl/torch_helpers.py:767-883
 *   ceeus_set_goal                    env.goal read by env.obs_postproc           mbrl/environments/fpp_construction_env.py:133-138
 *   ceeus_rollout                     ForwardModel.predict_n_steps                mbrl/models/gnn_ensemble.py:946-983,
 *                                                                                 mbrl/models/mlp_ensemble.py:575-611
 *   ceeus_costs                       trajectory_cost_fn + ensemble mean/std      mbrl/controllers/mpc_torch.py:116-137,323-332
 *                                                                                 mbrl/controllers/mpc_torch_curiosity.py:45-80
 *                                     env.cost_fn                                 mbrl/environments/fpp_construction_env.py:404-511,
 *                                                                                 mbrl/environments/playground_env_wgoals.py:108-157
 *   ceeus_sample                      sample_action_sequences + colored noise     mbrl/controllers/mpc_torch.py:204-247,
 *                                                                                 mbrl/controllers/colored_noise.py:115-218
 *   ceeus_begin_rollout               beginning_of_rollout (reset mean/std/elites) mbrl/controllers/mpc_torch.py:159-202
 *   ceeus_plan_iter_local /           one iCEM iteration of get_action            mbrl/controllers/mpc_torch.py:300-359
 *   ceeus_plan_iter_rollout / _costs  the same iteration split around env.cost_fn mbrl/controllers/mpc_torch.py:116-137
 *   ceeus_plan_iter_update            update_distributions (top-k, refit)         mbrl/controllers/mpc_torch.py:445-478
 *   ceeus_plan_finish                 action pick + shift-initialisation          mbrl/controllers/mpc_torch.py:383-421
 *   ceeus_plan_step                   TorchMpcICem.get_action (single GPU)        mbrl/controllers/mpc_torch.py:282-439
 *
 * Conventions
 *   - every function returns 0 on success or a negative CEEUS_ERR_* code; ceeus_last_error() gives the text.
 *   - "dev" pointers are device pointers owned by the caller (PyTorch); the library borrows them for work
 *     enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - "host" pointers are plain host memory read/written synchronously.
 *   - all floating point data is IEEE float32, all index data int32.
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef CEEUS_B200_H
#define CEEUS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CEEUS_ABI_VERSION 3

enum {
  CEEUS_OK = 0,
  CEEUS_ERR_INVALID = -1, /* bad argument / unsupported configuration */
  CEEUS_ERR_CUDA = -2,    /* a CUDA runtime call failed */
  CEEUS_ERR_STATE = -3,   /* call order violated (e.g. plan before set_weights) */
};

enum { CEEUS_MODEL_GNN = 0, CEEUS_MODEL_MLP = 1 };
enum { CEEUS_ACT_NONE = 0, CEEUS_ACT_RELU = 1, CEEUS_ACT_SILU = 2, CEEUS_ACT_TANH = 3 };
/* arithmetic of the rollout GEMMs: fp32 FMA (parity 1e-5) or tcgen05 tensor cores with 10-bit-mantissa
 * operands and fp32 accumulation (parity 1e-3). */
enum { CEEUS_PREC_FP32 = 0, CEEUS_PREC_TC = 1 };
/* CEEUS_ENV_EXTERNAL: the task cost of this environment is not built in; the caller evaluates env.cost_fn on the
 * materialised trajectories and hands the [nE*P,E,H] result to ceeus_plan_iter_costs (SURVEY.md 8b "Env hand-off") */
enum { CEEUS_ENV_NONE = 0, CEEUS_ENV_CONSTRUCTION = 1, CEEUS_ENV_PLAYGROUND = 2, CEEUS_ENV_EXTERNAL = 3,
       CEEUS_ENV_ROBODESK = 4 /* mbrl/environments/robodesk_env.py:53-135; cost_case = CEEUS_RD_*, sparse = (reward == "sparse") */ };
/* RoboDesk tasks (robodesk_env.py:58-73) */
enum {
  CEEUS_RD_OPEN_SLIDE = 0, CEEUS_RD_OPEN_DRAWER = 1, CEEUS_RD_PUSH_RED = 2, CEEUS_RD_PUSH_GREEN = 3, CEEUS_RD_PUSH_BLUE = 4,
  CEEUS_RD_BALL_OFF_TABLE = 5, CEEUS_RD_UPRIGHT_BLOCK_OFF_TABLE = 6, CEEUS_RD_FLAT_BLOCK_OFF_TABLE = 7
};
/* construction cost branches, fpp_construction_env.py:475-508 */
enum {
  CEEUS_CASE_TOWER = 0 /* "*tower*" and "Pyramid" */,
  CEEUS_CASE_PICKANDPLACE = 1 /* every other non Flip/Slide case */,
  CEEUS_CASE_FLIP = 2 /* per-coordinate distances on the euler angles, :422-439, :498-503 */,
  CEEUS_CASE_SLIDE = 3 /* per-coordinate distances on the positions */
};
enum { CEEUS_COST_SUM = 0, CEEUS_COST_BEST = 1 }; /* cost_along_trajectory */

typedef struct CeeusConfig {
  int32_t abi_version; /* = CEEUS_ABI_VERSION */
  /* ---- dynamics model ---- */
  int32_t model_kind;    /* CEEUS_MODEL_* */
  int32_t ensemble_size; /* E */
  int32_t n_obj;         /* N (GNN: >= 2) */
  int32_t agent_dim;     /* A */
  int32_t dyn_dim;       /* D */
  int32_t stat_dim;      /* S */
  int32_t action_dim;    /* U */
  int32_t goal_dim;      /* G; observation = [agent | N x dyn | N x stat | goal] */
  int32_t hidden_dim;    /* Hd */
  int32_t num_layers;    /* hidden layers per MLP (GNN: 2, dense MLP: 3) */
  int32_t act_fn;        /* CEEUS_ACT_* */
  int32_t layer_norm;    /* 0/1 */
  int32_t aggr_mean;     /* 1 = mean aggregation, 0 = sum */
  int32_t precision;     /* CEEUS_PREC_* */
  /* ---- iCEM planner (yaml controller_params / action_sampler_params) ---- */
  int32_t num_envs;      /* independent planning problems batched together (1 for the reference semantics) */
  int32_t num_traj;      /* P: trajectories per env simulated on THIS rank */
  int32_t horizon;       /* H */
  int32_t opt_iterations;
  int32_t num_elites;    /* min(elites_size, P_total / 2), >= 2 */
  int32_t num_kept;      /* int(num_elites * fraction_elites_reused) */
  int32_t relative_init;
  int32_t use_mean_actions;
  int32_t keep_previous_elites;
  int32_t shift_elites_over_time;
  int32_t execute_best_elite;
  int32_t use_ensemble_cost_std;
  int32_t colored_noise;
  int32_t cost_along_trajectory; /* CEEUS_COST_* */
  int32_t curiosity;             /* 1: TorchCuriosityMpcICem (ensemble disagreement), 0: TorchMpcICem (task cost) */
  int32_t extrinsic_reward;      /* curiosity only: add extrinsic_scale * env cost */
  float alpha;
  float init_std;
  float noise_beta;
  float extrinsic_scale;
  /* ---- task cost ---- */
  int32_t env_kind;      /* CEEUS_ENV_* */
  int32_t cost_case;     /* CEEUS_CASE_* (construction) */
  int32_t sparse;
  int32_t shaped_reward;
  float threshold;
  float buffer_threshold;
  /* Flip / Slide only (fpp_construction_env.py:93-105, :297-355): per-coordinate weights, buffer zone, sparse thresholds;
   * Flip with shaped_reward measures the gripper's distance from gripper_init (:457-458) */
  float coord_weights[3];
  float buffer_xyz[3];
  float cost_thres[3];
  float gripper_init[3];
  /* ---- multi GPU: trajectories are sharded, rank r owns global indices [r*P, (r+1)*P) of every env ---- */
  int32_t world_size;
  int32_t rank;
  /* ---- where the planner's trajectory costs are reduced (built-in costs; predictions live in a compact scratch either way) ----
   * 0: by a separate one-warp-per-trajectory launch after the rollout (64 warps per SM hide the L2 / HBM latency);
   * 1: in the tail of the rollout kernel itself, by the warps of member p % E once all E members of trajectory p have arrived
   *    (no extra launch; measured slower on B200 at the BASELINE sizes) */
  int32_t cost_in_rollout_tail;
  /* ---- GNN variants (yaml model_params / forward_model_params; mbrl/models/gnn_ensemble.py:96-132) ----
   * gnn_variant 0: GraphNeuralNetworkEnsemble, the agent is the global node (agent_as_global_node: true).
   *   edge_global = 1: ignore_agent_object: false -> edge_mlp_global on [node, agent, action], aggregated per sample and fed to the
   *   global MLP as well (gnn_modules.py:98-106, 225-238).
   * gnn_variant 1: HomogeneousGraphNeuralNetworkEnsemble (agent_as_global_node: false, gnn_modules.py:281-545): agent and objects
   *   are embedded into embedding_dim features, num_message_passing rounds of edge / node MLPs over the N + 1 nodes, linear output
   *   heads.  Both variants run on the CEEUS_PREC_FP32 path. */
  int32_t gnn_variant;
  int32_t edge_global;
  int32_t num_message_passing;
  int32_t embedding_dim;
  /* ---- CEEUS_PREC_TC, GNN: how a trajectory is laid out on the 128-row MMA tile (csrc/rollout_tc.cu) ----
   * 0: chosen per handle from a stage-count / MMA-time model of a launch over num_envs * num_traj trajectories (env CEEUS_TC_ROWS =
   *    node | global overrides); 1: "node major" -- N rows per trajectory, node MLP and global MLP as separate stages (2(N-1) + 6
   *    stages per model step); 2: "global row" -- N + 1 rows per trajectory, node and global MLP K-stacked into one stage each
   *    (2(N-1) + 3 stages per step).  Results agree to the CEEUS_PREC_TC bar either way. */
  int32_t tc_row_layout;
  uint64_t seed;         /* Philox key of the on-device noise sampler */
} CeeusConfig;

/* Caller-owned device buffers the planner reads and writes (bound once, may be rebound). Shapes use
 * P = num_traj, Pk = P_total_pop = P*world_size + num_kept (population incl. kept elites), k = num_elites,
 * nE = num_envs, O = A + N(D+S) + G.  The reference attributes they mirror are given on the right. */
typedef struct CeeusBuffers {
  float* mean;            /* [nE,H,U]            TorchMpcICem.mean_ */
  float* std;             /* [nE,H,U]            TorchMpcICem.std_ */
  float* actions;         /* [nE*P,H,U]          action_sequences of the current iteration (this rank) */
  float* traj;            /* [H+1,E,nE*P,O]      traj[:-1] = observations, traj[1:] = next_observations; may be NULL unless the
                                                 task cost is evaluated by the caller (CEEUS_ENV_EXTERNAL): the planner itself
                                                 keeps only the predicted dims in a compact scratch of its own */
  float* costs_per_model; /* [nE,P,E]            costs_per_model_ (this rank's fresh samples) */
  float* costs;           /* [nE,P]              costs_ (this rank's fresh samples) */
  float* elite_actions;   /* [nE,k,H,U]          elite_samples["actions"][:,0] */
  float* elite_costs;     /* [nE,k]              costs_ of the elites, ascending */
  float* elite_cpm;       /* [nE,k,E]            costs_per_model_ of the elites */
  int32_t* elite_idx;     /* [nE,k]              index into the global population of the last iteration */
  float* cand;            /* [nE,k,rec]          this rank's top-k candidate records, rec = ceeus_cand_record_len() */
  float* action_out;      /* [nE,U]              executed action */
  float* obs;             /* [nE,O]              optional: observation(s) resident on the device (ceeus_plan_step_resident) */
} CeeusBuffers;

typedef struct CeeusHandle_t* CeeusHandle;

int ceeus_abi_version(void);
const char* ceeus_last_error(CeeusHandle h); /* h may be NULL: error of the last failed ceeus_create */
int ceeus_create(const CeeusConfig* cfg, CeeusHandle* out);
void ceeus_destroy(CeeusHandle h);

/* number of float32 tensors ceeus_set_weights expects and, for i in [0,n), their (rows, cols) per member:
 * GNN order: for mlp in (edge, node, global): for l in 0..L: weight[in,out], bias[1,out], then if l<L and
 * layer_norm: gamma[1,out], beta[1,out].  Dense MLP: for l in 0..L: weight, bias (+ gamma, beta). */
int ceeus_num_weight_tensors(CeeusHandle h);
int ceeus_weight_tensor_shape(CeeusHandle h, int i, int32_t* rows, int32_t* cols);
/* dev_ptrs[i] -> contiguous float32 [E, rows_i, cols_i] in device memory (the reference's x @ W layout).
 * The library repacks into its own resident layout; pointers are not retained. */
int ceeus_set_weights(CeeusHandle h, const float* const* dev_ptrs, int n, void* stream);
/* host array: for each normaliser group, mean[dim] then std[dim].
 * GNN groups: in_agent(A) in_dyn(D) in_stat(S) in_action(U) out_agent(A) out_dyn(D); MLP: in(sd+U) out(sd). */
int ceeus_set_normalizers(CeeusHandle h, const float* host, int n_floats);
int ceeus_set_action_bounds(CeeusHandle h, const float* low_host, const float* high_host);
int ceeus_set_goal(CeeusHandle h, const float* goal_host /* [nE,G] */);
int ceeus_bind_buffers(CeeusHandle h, const CeeusBuffers* bufs);
int ceeus_cand_record_len(CeeusHandle h); /* floats per candidate record: 2 + E + H*U */

/* ---- stand-alone stages (predict_n_steps / trajectory_cost_fn / sample_action_sequences) ---- */
/* start_obs_dev [n_start,O] (n_start = 1 broadcast per env block, or = num_traj_total one per trajectory),
 * actions_dev [n,H,U], traj_dev [H+1,E,n,O]; trajectories [i*n/nE,(i+1)*n/nE) use goal i. */
int ceeus_rollout(CeeusHandle h, const float* start_obs_dev, int n_start, const float* actions_dev, int n,
                  float* traj_dev, void* stream);
/* traj_dev [H+1,E,n,O] -> costs_per_model_dev [n,E], costs_dev [n] (ensemble mean (+ std)). */
int ceeus_costs(CeeusHandle h, const float* traj_dev, int n, float* costs_per_model_dev, float* costs_dev,
                void* stream);
/* actions_dev [n,H,U] = clip(y*std + mean, low, high), row r uses mean/std of env r / (n/nE).
 * noise_dev != NULL: y = noise_dev [n,U,H] (the torch_powerlaw_psd_gaussian boundary);
 * else y is drawn on device: Philox4x32-10(seed; stream, global row = first_row + r) -> Box-Muller -> irDFT.
 * If gauss_dev != NULL it receives/provides the standard normals [n,U,2F] (testing hook: write_gauss=1 stores the
 * drawn normals, write_gauss=0 uses them instead of Philox). */
int ceeus_sample(CeeusHandle h, const float* mean_dev, const float* std_dev, const float* noise_dev,
                 float* gauss_dev, int write_gauss, uint32_t stream_id, int64_t first_row, int n,
                 float* actions_dev, void* stream);

/* ---- planner ---- */
int ceeus_begin_rollout(CeeusHandle h, void* stream);
/* One iCEM iteration on this rank: sample (+ mean-action / shifted-elite insertion), rollout, costs, local top-k
 * into bufs.cand.  obs_dev [nE,O].  noise_dev: NULL (device RNG) or [nE*P,U,H]; shift_noise_dev: NULL or [nE*n_kept,U,H]
 * (only read in iteration 0 when elites are shifted over time). */
int ceeus_plan_iter_local(CeeusHandle h, int iter, const float* obs_dev, const float* noise_dev,
                          const float* shift_noise_dev, void* stream);
/* The same iteration in two halves, for environments whose cost_fn runs in the caller (CEEUS_ENV_EXTERNAL):
 * ceeus_plan_iter_rollout = sample (+ insertions) and predict_n_steps into bufs.traj (mpc_torch.py:301-317);
 * ceeus_plan_iter_costs = trajectory_cost_fn + ensemble reduce + local top-k (mpc_torch.py:323-337, 445-460) with
 * env_cost_dev [nE*P,E,H] = env.cost_fn(observations, actions, next_observations) evaluated by the caller on bufs.traj
 * (NULL = the built-in cost of env_kind).  ceeus_plan_iter_local == rollout + costs(NULL). */
int ceeus_plan_iter_rollout(CeeusHandle h, int iter, const float* obs_dev, const float* noise_dev,
                            const float* shift_noise_dev, void* stream);
int ceeus_plan_iter_costs(CeeusHandle h, const float* env_cost_dev, void* stream);
/* Merge candidate records of all ranks (cand_all_dev [world,nE,k,rec]; world=1: pass bufs.cand) with the kept elites,
 * select the global top-k (ties: lower global index first), refit mean/std. Identical result on every rank. */
int ceeus_plan_iter_update(CeeusHandle h, int iter, const float* cand_all_dev, void* stream);
/* executed action -> bufs.action_out, shift mean by one step, reset std. */
int ceeus_plan_finish(CeeusHandle h, void* stream);
/* Whole MPC step on one GPU: H2D obs, opt_iterations x (local, update), finish, D2H action (synchronises `stream`).
 * noise_host_or_null: concatenation in the reference's draw order: iter0 [nE*P,U,H], shift [nE*n_kept,U,H] (only if
 * elites are being shifted), iter1, iter2, ... ; NULL = device RNG. */
int ceeus_plan_step(CeeusHandle h, const float* obs_host, float* action_host, const float* noise_dev_or_null,
                    void* stream);
/* counters for bench.py: kernels launched by this handle since creation */
/* The same MPC step with the observation(s) already on the device (CeeusBuffers.obs) and the action left in
 * CeeusBuffers.action_out: no host copy, no host synchronisation; `stream` continues after the step (callers that keep their
 * environments on the GPU, and the device-timed leg of bench.py). */
int ceeus_plan_step_resident(CeeusHandle h, const float* noise_dev_or_null, void* stream);

/* ---- multi GPU (one process per GPU) ----
 * The trajectories of one planning problem are sharded over world_size ranks; per iCEM iteration the ranks exchange their k best
 * candidate records with one ncclAllGather that is part of the step's CUDA graph, so ceeus_plan_step stays ONE graph launch at any
 * number of GPUs.  Rank 0 obtains an id with ceeus_comm_unique_id and hands it to the other ranks by any means (the Python layer
 * broadcasts it through torch.distributed); every rank then calls ceeus_comm_init (collective).  NCCL is bound at run time
 * (dlopen of libnccl.so.2, preferably the copy already loaded into the process; CEEUS_NCCL_LIB overrides the path).
 * Reference precedent for sharding trajectories over workers: ParallelGroundTruthModel.predict_n_steps, mbrl/models/gt_par_model.py:91-128. */
/* Preferred exchange on one NVSwitch node (2..8 ranks): every rank exports its gather buffer as a CUDA IPC handle (64 bytes), the
 * handles of all ranks are gathered by the caller and imported; from then on the candidate records travel as direct NVLink stores
 * from a small kernel into the peers' buffers, followed by a flag -- no NCCL launch on the planning path.  A rank that stops
 * planning surfaces in its peers as protocol error 21 (ceeus_tc_status), not as a hang. */
int ceeus_p2p_export(CeeusHandle h, uint8_t* handle_out /* [64] */);
int ceeus_p2p_import(CeeusHandle h, const uint8_t* handles /* [world_size][64], rank-major */);
int ceeus_comm_unique_id(uint8_t* id_out /* [128] */);
int ceeus_comm_init(CeeusHandle h, const uint8_t* unique_id /* [128] */);

int64_t ceeus_kernel_launches(CeeusHandle h);
/* Measurement aid: with enable != 0 every rollout launch of the eager path (ceeus_rollout, ceeus_plan_iter_*) is bracketed by
 * a CUDA-event pair on its launch stream (up to 256 launches between two reads).  Each call first copies the durations (ms) of
 * the pairs recorded since the previous call into ms_out_host (may be NULL) and returns their number, then sets the mode. */
int ceeus_profile_rollouts(CeeusHandle h, int enable, float* ms_out_host, int max_out);
/* CEEUS_PREC_TC only: status word of the tensor-core kernels since the handle was created; synchronises the device.
 * 0 = clean.  Bits below CEEUS_TC_RANGE: protocol error code (a bounded mbarrier wait timed out; 21 = a peer rank never arrived).
 * CEEUS_TC_RANGE: an fp16 operand was saturated -- the operands are fp16 (5-bit exponent), conversions clamp to +-65504 instead of
 * producing inf, and the event is recorded here; ceeus_plan_step / _resident then fail with CEEUS_ERR_STATE.  With LayerNorm the
 * hidden activations cannot saturate; normalised observations / actions beyond +-65504 can. */
#define CEEUS_TC_RANGE 0x40000000
int ceeus_tc_status(CeeusHandle h);
/* CEEUS_PREC_TC: how a launch over n trajectories is tiled, per ensemble member e: rounds_out[e] = trajectory groups each
 * warpgroup ("slot") takes through the horizon one after the other, uneven_out[e] = 1 if the member's CTA pairs use the 3/4
 * split that keeps the MMA-issuing warp empty (GNN kernel only).  The parity tests assert with it that every tiling branch
 * of the kernels is compared with the oracle. */
int ceeus_tc_tiling(CeeusHandle h, int n, int32_t* rounds_out /* [E] */, int32_t* uneven_out /* [E] */);
/* CEEUS_PREC_TC, GNN: the tile row layout this handle runs (1 = node major, 2 = global row; see CeeusConfig.tc_row_layout); 0 for
 * the dense-MLP kernel. */
int ceeus_tc_row_layout(CeeusHandle h);
/* 1 once elites from a previous MPC step exist (controls whether shift noise is consumed) */
int ceeus_has_elites(CeeusHandle h);

#ifdef __cplusplus
}
#endif
#endif /* CEEUS_B200_H */
