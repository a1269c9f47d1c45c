// C ABI of libceeus_b200.so (see include/ceeus_b200.h): handle, weight hand-off, and the launch sequence of one
// iCEM planning step.  No CPU fallback: every entry point either enqueues CUDA work or returns an error.
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "cost_device.cuh"
#ifdef CEEUS_DEBUG_TOOLS
#include "debug_tools.h"
#endif
#include "kernels.cuh"

using namespace ceeus;

namespace {
std::string g_create_error;

// ---- NCCL, bound at run time ---------------------------------------------------------------------------------------------
// The library is dlopen'ed (the copy torch has already loaded into the process, if any) so that libceeus_b200.so has no link-time
// dependency on it; only the five entry points below are used.  Prototypes as in nccl.h (stable across NCCL 2.x).
struct NcclApi {
  using Comm = void*;
  struct UniqueId { char internal[128]; };
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int /* ncclDataType_t */, Comm, cudaStream_t) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  std::string why;
};
constexpr int kNcclFloat32 = 7;

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void* lib = nullptr;
  if (const char* path = std::getenv("CEEUS_NCCL_LIB")) lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy already in the process (torch)
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { api.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return api; }
  auto sym = [&](const char* n) { return dlsym(lib, n); };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString;
  if (!api.ok) api.why = "libnccl.so.2 lacks an expected symbol";
  return api;
}

struct TensorSpec {
  int rows, cols;  // per member
  int dst_off, dst_ld;
};
}  // namespace

struct CeeusHandle_t {
  CeeusConfig cfg{};
  ModelDesc md{};
  CostDesc cd{};
  int device = 0, num_sms = 0;
  float* d_weights = nullptr;
  float* d_norm = nullptr;
  float* d_low = nullptr;
  float* d_high = nullptr;
  float* d_goal = nullptr;
  float* d_obs = nullptr;
  float* d_irdft = nullptr;
  float* h_obs = nullptr;     // pinned staging
  float* h_action = nullptr;  // pinned staging
  int* h_tc_err = nullptr;    // pinned copy of the tensor-core kernels' protocol-error flag (read back with the action)
  int norm_floats = 0;
  TcLayout tc{};             // CEEUS_PREC_TC only (GNN ensemble)
  TcMlpLayout tcm{};         // CEEUS_PREC_TC only (dense-MLP ensemble)
  uint8_t* d_wimg = nullptr;
  float* d_wpar = nullptr;
  int* d_kmap = nullptr;         // K permutation / folding map of every stage matrix
  float* d_tc_scratch = nullptr; // logical fp32 stage matrices (pack-time only)
  bool tc_all_fold = false;      // every member's LayerNorm affine could be folded into the weights (read back after packing)
  int* d_tc_err = nullptr;
  // planner scratch of the fused cost tail: predicted dims of every (trajectory, step, member) + per-trajectory arrival counters
  float* d_pred = nullptr;
  int* d_done = nullptr;
  long long pred_stride = 0;
  int pred_Op = 0;
  bool fuse_costs = true;   // CEEUS_NO_FUSE=1: separate cost_kernel on fully materialised trajectories (A/B measurements)
  bool cost_tail = false;   // reduce the costs in the tail of the rollout kernel instead of in a separate launch (cost_in_rollout_tail)
  const float* last_plan_obs = nullptr;  // start observations of the last planner rollout (read by the compact cost kernel)
  bool l2_discard = true;   // CEEUS_NO_DISCARD=1: keep the consumed prediction blocks in L2 (write-back on eviction)
  long long* d_tc_dbg = nullptr;  // optional timeline of one warpgroup (ceeus_tc_debug_timeline)
  CeeusBuffers bufs{};
  // multi-GPU: NCCL communicator over the ranks that share the trajectories of one planning problem (ceeus_comm_init) and the
  // gathered candidate records of all ranks [world,nE,k,rec]
  void* comm = nullptr;
  float* d_cand_all = nullptr;
  // NVLink peer-memory exchange (ceeus_p2p_export / _import): this rank's gather double buffer + arrival flags in ONE allocation
  // that the peers map through CUDA IPC; p2p_buf[r] / p2p_flags[r] are the mapped buffers of all ranks (own included)
  void* p2p_local = nullptr;
  float* p2p_buf[8] = {nullptr};
  unsigned long long* p2p_flags[8] = {nullptr};
  bool p2p_ready = false;
  bool bound = false, weights_set = false, norms_set = false;
  bool has_elites = false;
  unsigned long long* d_step_ctr = nullptr;  // MPC steps planned so far (base of the Philox stream ids), lives on the device
  int64_t launches = 0;
  // ceeus_plan_step as a CUDA graph: [0] first step after begin_rollout (no elites to shift), [1] steady state
  cudaStream_t gstream = nullptr;
  cudaEvent_t gevent = nullptr;
  // [has elites][0: host observation in / host action out, 1: observation resident in bufs.obs, action left in bufs.action_out]
  cudaGraphExec_t graph_exec[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  int64_t graph_launches[2][2] = {{0, 0}, {0, 0}};
  bool capturing = false;
  // optional CUDA-event pairs around every rollout launch of the eager path (ceeus_profile_rollouts): in-step kernel time
  std::vector<cudaEvent_t> prof_ev;
  int prof_n = 0;
  bool prof_on = false;
  int F = 0;
  std::vector<TensorSpec> wspec;
  std::string err;
};

namespace {

int fail(CeeusHandle h, int code, const std::string& msg) {
  if (h) h->err = msg;
  else g_create_error = msg;
  return code;
}

#define CU_CHECK(h, expr)                                                                                \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return fail(h, CEEUS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
  } while (0)

#define LAUNCH(h, expr)   \
  do {                    \
    CU_CHECK(h, expr);    \
    (h)->launches += 1;   \
  } while (0)

void add_mlp(CeeusHandle h, MlpDesc& m, int din, int dout, int Hd, int L, bool ln, int& off) {
  m.n_layers = L + 1;
  for (int l = 0; l <= L; ++l) {
    LayerDesc& ld = m.layer[l];
    ld.K = l == 0 ? din : Hd;
    ld.N = l == L ? dout : Hd;
    ld.Kpad = round_up(ld.K, 4);
    ld.Npad = round_up(ld.N, 64);
    ld.w_off = off;
    off += ld.Kpad * ld.Npad;
    ld.b_off = off;
    off += ld.Npad;
    ld.g_off = ld.be_off = 0;
    h->wspec.push_back({ld.K, ld.N, ld.w_off, ld.Npad});
    h->wspec.push_back({1, ld.N, ld.b_off, ld.Npad});
    if (l < L && ln) {
      ld.g_off = off;
      off += ld.Npad;
      ld.be_off = off;
      off += ld.Npad;
      h->wspec.push_back({1, ld.N, ld.g_off, ld.Npad});
      h->wspec.push_back({1, ld.N, ld.be_off, ld.Npad});
    }
  }
}

// irDFT matrix of the power-law noise (colored_noise.py:174-216), built in double.
std::vector<float> build_irdft(double beta, int H) {
  const int F = H / 2 + 1;
  std::vector<double> s(F);
  for (int k = 0; k < F; ++k) s[k] = std::pow((double)(k == 0 ? 1 : k) / H, -beta / 2.0);
  double sum = 0;
  for (int k = 1; k < F; ++k) {
    double w = s[k];
    if (k == F - 1) w *= (1 + (H % 2)) / 2.0;
    sum += w * w;
  }
  const double sigma = 2.0 * std::sqrt(sum) / H;
  std::vector<float> M((size_t)2 * F * H, 0.f);
  const double pi = 3.14159265358979323846;
  for (int k = 0; k < F; ++k)
    for (int t = 0; t < H; ++t) {
      const bool nyq = (H % 2 == 0) && (k == F - 1);
      double re, im = 0.0;
      if (k == 0) re = s[0] * std::sqrt(2.0);
      else if (nyq) re = s[k] * std::sqrt(2.0) * std::cos(pi * t);
      else {
        re = 2 * s[k] * std::cos(2 * pi * k * t / H);
        im = -2 * s[k] * std::sin(2 * pi * k * t / H);
      }
      M[(size_t)k * H + t] = (float)(re / (H * sigma));
      M[(size_t)(F + k) * H + t] = (float)(im / (H * sigma));
    }
  return M;
}

int validate(const CeeusConfig& c, std::string& why) {
  auto bad = [&](const char* m) {
    why = m;
    return CEEUS_ERR_INVALID;
  };
  if (c.abi_version != CEEUS_ABI_VERSION) return bad("abi_version mismatch");
  if (c.model_kind != CEEUS_MODEL_GNN && c.model_kind != CEEUS_MODEL_MLP) return bad("model_kind");
  // (one member is enough where nothing is reduced over the ensemble: task costs of a single model, the RND controller)
  const int min_members = (c.curiosity || c.use_ensemble_cost_std) ? 2 : 1;
  if (c.ensemble_size < min_members || c.ensemble_size > kMaxEnsemble) return bad("ensemble_size must be in [2,8] (>= 1 without an ensemble statistic)");
  if (c.num_layers < 1 || c.num_layers + 1 > kMaxLayers) return bad("num_layers must be in [1,3]");
  if (c.hidden_dim != 64 && c.hidden_dim != 128 && c.hidden_dim != 256) return bad("hidden_dim must be 64, 128 or 256");
  if (c.model_kind == CEEUS_MODEL_GNN) {
    if (c.hidden_dim == 256 && c.precision != CEEUS_PREC_FP32) return bad("GNN hidden_dim 256 runs on the CEEUS_PREC_FP32 path only");
    if (c.n_obj < 2 && c.gnn_variant == 0) return bad("GNN needs n_obj >= 2 (gnn_modules.py:190-241: row undefined for one node)");
    if (c.gnn_variant != 0 && c.gnn_variant != 1) return bad("gnn_variant");
    if (c.gnn_variant == 1 && (c.embedding_dim < 1 || c.embedding_dim > 64 || c.n_obj < 1))
      return bad("homogeneous GNN: embedding_dim must be in [1,64]");
    if (c.gnn_variant == 1 && (c.n_obj + 1) * c.n_obj > kTileRows) return bad("homogeneous GNN: (N+1) N edges must be <= 64");
    // the fp32 path tiles whole trajectories' edges into 64 rows; the tensor-core path (node-major rows) takes N <= 32
    if (c.precision == CEEUS_PREC_FP32 && c.n_obj * (c.n_obj - 1) > kTileRows) return bad("n_obj too large for fp32 mode: N(N-1) must be <= 64");
    if (c.dyn_dim > 64 || c.agent_dim > 64) return bad("dyn_dim / agent_dim must be <= 64");
    if (c.agent_dim < 1) return bad("agent_dim must be >= 1 (agent_as_global_node)");
  } else {
    const int sd = c.agent_dim + c.n_obj * (c.dyn_dim + c.stat_dim);
    if (round_up(sd, 64) > c.hidden_dim || sd > 128) return bad("MLP state dim too large for this build");
  }
  if (c.num_envs < 1 || c.num_traj < 2 || c.horizon < 1 || c.horizon > 256) return bad("num_envs/num_traj/horizon");
  if (sample_smem_bytes(c.horizon, c.colored_noise) > 227 * 1024)
    return bad("horizon too long for the colored-noise sampler (the [2F x H] irDFT matrix must fit shared memory: H <= 224)");
  if (c.opt_iterations < 1) return bad("opt_iterations");
  if (c.num_elites < 1 || c.num_elites > 64 || c.num_elites > c.num_traj) return bad("num_elites must be in [1,64] and <= num_traj");
  if (c.num_kept < 0 || c.num_kept > c.num_elites) return bad("num_kept");
  if (c.world_size < 1 || c.rank < 0 || c.rank >= c.world_size) return bad("world_size/rank");
  if (c.world_size * c.num_elites + c.num_kept > 256) return bad("world_size * num_elites too large");
  if (c.precision != CEEUS_PREC_FP32 && c.precision != CEEUS_PREC_TC) return bad("precision");
  const bool need_env = !c.curiosity || c.extrinsic_reward;
  if (need_env && c.env_kind == CEEUS_ENV_NONE) return bad("task cost requested but env_kind is NONE");
  if (c.env_kind < CEEUS_ENV_NONE || c.env_kind > CEEUS_ENV_ROBODESK) return bad("env_kind");
  if (c.env_kind == CEEUS_ENV_ROBODESK && (c.agent_dim < 24 || c.dyn_dim != 13 || c.n_obj < 8 || c.cost_case < 0 || c.cost_case > 7))
    return bad("RoboDesk layout: agent_dim >= 24, 8 objects of 13 dynamic features (robodesk_env.py:20-23), cost_case = CEEUS_RD_*");
  if (c.env_kind == CEEUS_ENV_CONSTRUCTION && (c.goal_dim != 3 * c.n_obj + 3 || c.dyn_dim < 3 || c.agent_dim < 3))
    return bad("construction layout: goal_dim must be 3N+3");
  if (c.env_kind == CEEUS_ENV_PLAYGROUND && (c.goal_dim != 2 * c.n_obj || c.dyn_dim < 2)) return bad("playground layout: goal_dim must be 2N");
  if (c.cost_along_trajectory != CEEUS_COST_SUM && c.cost_along_trajectory != CEEUS_COST_BEST) return bad("cost_along_trajectory");
  return CEEUS_OK;
}

cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace

extern "C" {

int ceeus_abi_version(void) { return CEEUS_ABI_VERSION; }

const char* ceeus_last_error(CeeusHandle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ceeus_create(const CeeusConfig* cfg, CeeusHandle* out) {
  if (!cfg || !out) return fail(nullptr, CEEUS_ERR_INVALID, "null argument");
  std::string why;
  if (int rc = validate(*cfg, why)) return fail(nullptr, rc, why);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, CEEUS_ERR_CUDA, "no CUDA device: libceeus_b200 has no CPU fallback");
  CeeusHandle h = new CeeusHandle_t();
  h->cfg = *cfg;
  const CeeusConfig& c = h->cfg;
  cudaGetDevice(&h->device);
  cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device);

  ModelDesc& md = h->md;
  md.kind = c.model_kind;
  md.E = c.ensemble_size; md.N = c.n_obj; md.A = c.agent_dim; md.D = c.dyn_dim; md.S = c.stat_dim; md.U = c.action_dim;
  md.G = c.goal_dim; md.Hd = c.hidden_dim; md.L = c.num_layers;
  md.act = c.act_fn; md.layer_norm = c.layer_norm; md.aggr_mean = c.aggr_mean;
  md.sd = c.agent_dim + c.n_obj * (c.dyn_dim + c.stat_dim);
  md.O = md.sd + c.goal_dim;
  int off = 0;
  const bool ln = c.layer_norm != 0;
  if (c.model_kind == CEEUS_MODEL_GNN) {
    const int na = c.dyn_dim + c.stat_dim;
    md.variant = c.gnn_variant; md.edge_global = c.edge_global; md.n_mp = c.num_message_passing > 0 ? c.num_message_passing : 1;
    md.emb = c.embedding_dim;
    if (c.gnn_variant == 0) {
      // weight order = edge_mlp, node_mlp, global_mlp (, edge_mlp_global)
      add_mlp(h, md.mlp[0], 2 * na + c.agent_dim + c.action_dim, c.hidden_dim, c.hidden_dim, c.num_layers, ln, off);
      add_mlp(h, md.mlp[1], na + c.agent_dim + c.action_dim + c.hidden_dim, c.dyn_dim, c.hidden_dim, c.num_layers, ln, off);
      add_mlp(h, md.mlp[2], c.agent_dim + c.action_dim + c.hidden_dim * (c.edge_global ? 2 : 1), c.agent_dim, c.hidden_dim,
              c.num_layers, ln, off);
      if (c.edge_global) add_mlp(h, md.mlp[3], na + c.agent_dim + c.action_dim, c.hidden_dim, c.hidden_dim, c.num_layers, ln, off);
    } else {
      // weight order = edge_mlp, node_mlp, embed_agent, embed_obj, output_head_agent, output_head_objdyn (the heads are single
      // linear layers: num_layers 0, no LayerNorm, gnn_modules.py:351-381)
      const int F = c.embedding_dim;
      add_mlp(h, md.mlp[0], 2 * F + c.action_dim, c.hidden_dim, c.hidden_dim, c.num_layers, ln, off);
      add_mlp(h, md.mlp[1], F + c.action_dim + c.hidden_dim, F, c.hidden_dim, c.num_layers, ln, off);
      add_mlp(h, md.mlp[2], c.agent_dim, F, c.hidden_dim, 0, false, off);
      add_mlp(h, md.mlp[3], na, F, c.hidden_dim, 0, false, off);
      add_mlp(h, md.mlp[4], F, c.agent_dim, c.hidden_dim, 0, false, off);
      add_mlp(h, md.mlp[5], F, c.dyn_dim, c.hidden_dim, 0, false, off);
    }
    int o = 0;
    md.nrm.in_agent = o; o += 2 * md.A;
    md.nrm.in_dyn = o; o += 2 * md.D;
    md.nrm.in_stat = o; o += 2 * md.S;
    md.nrm.in_action = o; o += 2 * md.U;
    md.nrm.out_agent = o; o += 2 * md.A;
    md.nrm.out_dyn = o; o += 2 * md.D;
    h->norm_floats = o;
    const GnnTiling tl = make_gnn_tiling(md);
    if ((c.gnn_variant != 0 || c.edge_global) && c.precision != CEEUS_PREC_FP32) {
      delete h;
      return fail(nullptr, CEEUS_ERR_INVALID, "the homogeneous GNN and ignore_agent_object=false run on the CEEUS_PREC_FP32 path only");
    }
    if (c.precision == CEEUS_PREC_FP32 && (tl.G < 1 || tl.smem_bytes > 227 * 1024)) {
      delete h;
      return fail(nullptr, CEEUS_ERR_INVALID, "GNN tiling does not fit shared memory");
    }
  } else {
    add_mlp(h, md.mlp[0], md.sd + c.action_dim, md.sd, c.hidden_dim, c.num_layers, ln, off);
    md.nrm.in_all = 0;
    md.nrm.out_all = 2 * (md.sd + md.U);
    h->norm_floats = 2 * (md.sd + md.U) + 2 * md.sd;
  }
  md.member_stride = off;

  CostDesc& cd = h->cd;
  cd.curiosity = c.curiosity; cd.extrinsic = c.extrinsic_reward; cd.cost_along = c.cost_along_trajectory;
  cd.env_kind = c.env_kind; cd.cost_case = c.cost_case; cd.sparse = c.sparse; cd.shaped = c.shaped_reward;
  cd.use_std = c.use_ensemble_cost_std; cd.extrinsic_scale = c.extrinsic_scale; cd.threshold = c.threshold;
  cd.buffer_threshold = c.buffer_threshold;
  for (int i = 0; i < 3; ++i) {
    cd.coord_weights[i] = c.coord_weights[i]; cd.buffer_xyz[i] = c.buffer_xyz[i];
    cd.cost_thres[i] = c.cost_thres[i]; cd.gripper_init[i] = c.gripper_init[i];
  }

  h->fuse_costs = std::getenv("CEEUS_NO_FUSE") == nullptr;
  h->cost_tail = c.cost_in_rollout_tail != 0;
  if (const char* ev = std::getenv("CEEUS_COST_TAIL")) h->cost_tail = std::atoi(ev) != 0;
  h->l2_discard = std::getenv("CEEUS_NO_DISCARD") == nullptr;
  h->pred_Op = md.kind == CEEUS_MODEL_GNN ? md.A + md.N * md.D : md.sd;
  // trajectory-major blocks of whole 128-byte lines for the GNN kernels, step-major rows for the dense-MLP kernels (cost_device.cuh)
  h->pred_stride = md.kind == CEEUS_MODEL_GNN ? round_up(c.horizon * md.E * h->pred_Op, 32) : 0;
  h->F = c.horizon / 2 + 1;
  auto alloc = [&](float** p, size_t n) { return cudaMalloc(p, (n ? n : 1) * sizeof(float)); };
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = alloc(&h->d_weights, (size_t)md.E * md.member_stride);
  if (e == cudaSuccess) e = cudaMemset(h->d_weights, 0, (size_t)md.E * md.member_stride * sizeof(float));
  if (e == cudaSuccess) e = alloc(&h->d_norm, h->norm_floats);
  if (e == cudaSuccess) e = alloc(&h->d_low, md.U);
  if (e == cudaSuccess) e = alloc(&h->d_high, md.U);
  if (e == cudaSuccess) e = alloc(&h->d_goal, (size_t)c.num_envs * (md.G ? md.G : 1));
  if (e == cudaSuccess) e = cudaMemset(h->d_goal, 0, (size_t)c.num_envs * (md.G ? md.G : 1) * sizeof(float));
  if (e == cudaSuccess) e = alloc(&h->d_obs, (size_t)c.num_envs * md.O);
  if (e == cudaSuccess) e = alloc(&h->d_irdft, (size_t)2 * h->F * c.horizon);
  if (e == cudaSuccess) e = cudaMalloc(&h->d_step_ctr, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(h->d_step_ctr, 0, sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->gstream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->gevent, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMallocHost(&h->h_obs, (size_t)c.num_envs * md.O * sizeof(float));
  if (e == cudaSuccess) e = cudaMallocHost(&h->h_tc_err, sizeof(int));
  if (e == cudaSuccess) *h->h_tc_err = 0;
  if (e == cudaSuccess) e = cudaMallocHost(&h->h_action, (size_t)c.num_envs * md.U * sizeof(float));
  if (e == cudaSuccess && c.precision == CEEUS_PREC_TC && c.model_kind == CEEUS_MODEL_MLP) {
    if (!make_mlp_tc_layout(md, h->num_sms, &h->tcm)) {
      ceeus_destroy(h);
      return fail(nullptr, CEEUS_ERR_INVALID,
                  "CEEUS_PREC_TC supports the dense-MLP ensemble with hidden_dim 128/256, no LayerNorm, state dim <= 128");
    }
    e = cudaMalloc(&h->d_wimg, (size_t)md.E * 2 * h->tcm.image_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_tc_err, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(h->d_tc_err, 0, sizeof(int));
  } else if (e == cudaSuccess && c.precision == CEEUS_PREC_TC) {
    std::vector<int> kmap(4096, -1);
    int grow = 0;
    if (c.model_kind == CEEUS_MODEL_GNN) {
      grow = c.tc_row_layout == 1 ? 0 : c.tc_row_layout == 2 ? 1 : tc_pick_grow(md, h->num_sms, c.num_envs * c.num_traj);
      if (c.tc_row_layout == 0) {
        const char* ov = getenv("CEEUS_TC_ROWS");  // development / A-B override of the automatic choice
        if (ov && ov[0] == 'n') grow = 0;
        if (ov && ov[0] == 'g') grow = 1;
      }
    }
    if (c.model_kind != CEEUS_MODEL_GNN || !make_tc_layout(md, h->num_sms, &h->tc, kmap.data(), grow)) {
      ceeus_destroy(h);
      return fail(nullptr, CEEUS_ERR_INVALID,
                  "CEEUS_PREC_TC supports the GNN ensemble with hidden_dim 64/128, 2 hidden layers, n_obj <= 32, "
                  "agent_dim + action_dim <= 15, dyn_dim + stat_dim <= 16 and a trajectory group that fits shared memory");
    }
    e = cudaMalloc(&h->d_wimg, (size_t)md.E * 2 * h->tc.image_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_wpar, (size_t)md.E * h->tc.par_floats * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_kmap, (size_t)h->tc.kmap_len * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_kmap, kmap.data(), (size_t)h->tc.kmap_len * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_tc_scratch, (size_t)md.E * h->tc.scratch_floats * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_tc_err, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(h->d_tc_err, 0, sizeof(int));
  }
  if (e == cudaSuccess) {
    std::vector<float> M = build_irdft(c.noise_beta, c.horizon);
    e = cudaMemcpy(h->d_irdft, M.data(), M.size() * sizeof(float), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) {
    std::vector<float> lo(md.U, -1.f), hi(md.U, 1.f);
    e = cudaMemcpy(h->d_low, lo.data(), md.U * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_high, hi.data(), md.U * sizeof(float), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    std::string msg = std::string("allocation failed: ") + cudaGetErrorString(e);
    ceeus_destroy(h);
    return fail(nullptr, CEEUS_ERR_CUDA, msg);
  }
  *out = h;
  return CEEUS_OK;
}

static void drop_graphs(CeeusHandle h) {
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j)
      if (h->graph_exec[i][j]) {
        cudaGraphExecDestroy(h->graph_exec[i][j]);
        h->graph_exec[i][j] = nullptr;
      }
}

void ceeus_destroy(CeeusHandle h) {
  if (!h) return;
  drop_graphs(h);
  if (h->comm && nccl().ok) nccl().CommDestroy(h->comm);
  for (int r = 0; r < 8; ++r)
    if (h->p2p_ready && r != h->cfg.rank && h->p2p_buf[r]) cudaIpcCloseMemHandle(h->p2p_buf[r]);
  cudaFree(h->p2p_local);
  cudaFree(h->d_cand_all);
  if (h->gstream) cudaStreamDestroy(h->gstream);
  if (h->gevent) cudaEventDestroy(h->gevent);
  for (auto e : h->prof_ev) if (e) cudaEventDestroy(e);
  cudaFree(h->d_step_ctr);
  cudaFree(h->d_weights); cudaFree(h->d_norm); cudaFree(h->d_low); cudaFree(h->d_high);
  cudaFree(h->d_goal); cudaFree(h->d_obs); cudaFree(h->d_irdft);
  cudaFree(h->d_pred); cudaFree(h->d_done);
  cudaFree(h->d_wimg); cudaFree(h->d_wpar); cudaFree(h->d_kmap); cudaFree(h->d_tc_scratch); cudaFree(h->d_tc_err); cudaFree(h->d_tc_dbg);
  if (h->h_obs) cudaFreeHost(h->h_obs);
  if (h->h_action) cudaFreeHost(h->h_action);
  if (h->h_tc_err) cudaFreeHost(h->h_tc_err);
  delete h;
}

int ceeus_num_weight_tensors(CeeusHandle h) { return h ? (int)h->wspec.size() : CEEUS_ERR_INVALID; }

int ceeus_weight_tensor_shape(CeeusHandle h, int i, int32_t* rows, int32_t* cols) {
  if (!h || i < 0 || i >= (int)h->wspec.size() || !rows || !cols) return fail(h, CEEUS_ERR_INVALID, "weight_tensor_shape: bad index");
  *rows = h->wspec[i].rows;
  *cols = h->wspec[i].cols;
  return CEEUS_OK;
}

int ceeus_set_weights(CeeusHandle h, const float* const* dev_ptrs, int n, void* stream) {
  if (!h || !dev_ptrs) return fail(h, CEEUS_ERR_INVALID, "set_weights: null argument");
  if (n != (int)h->wspec.size()) return fail(h, CEEUS_ERR_INVALID, "set_weights: expected " + std::to_string(h->wspec.size()) + " tensors");
  // the captured plan-step graphs hold the rollout kernel VARIANT chosen from tc_all_fold (recomputed below): re-capture
  drop_graphs(h);
  for (int i = 0; i < n; ++i) {
    if (!dev_ptrs[i]) return fail(h, CEEUS_ERR_INVALID, "set_weights: null tensor pointer");
    const TensorSpec& t = h->wspec[i];
    LAUNCH(h, launch_repack(dev_ptrs[i], h->d_weights, h->md.E, t.rows, t.cols, h->md.member_stride, t.dst_off, t.dst_ld, S(stream)));
  }
  if (h->cfg.precision == CEEUS_PREC_TC && h->md.kind == CEEUS_MODEL_MLP)
    LAUNCH(h, launch_mlp_tc_pack(h->d_weights, h->md, h->tcm, h->d_wimg, S(stream)));
  else if (h->cfg.precision == CEEUS_PREC_TC)
  {
    LAUNCH(h, launch_tc_pack(h->d_weights, h->md, h->tc, h->d_kmap, h->d_tc_scratch, h->d_wimg, h->d_wpar, S(stream)));
    // which epilogue variant the rollout kernel may use: one small read-back per set_weights (rare: after training / load)
    CU_CHECK(h, cudaStreamSynchronize(S(stream)));
    h->tc_all_fold = true;
    for (int e = 0; e < h->md.E; ++e) {
      float f = 0.f;
      CU_CHECK(h, cudaMemcpy(&f, h->d_wpar + (size_t)e * h->tc.par_floats + h->tc.par_fold, sizeof(float), cudaMemcpyDeviceToHost));
      if (f == 0.f) h->tc_all_fold = false;
    }
  }
  h->weights_set = true;
  return CEEUS_OK;
}

int ceeus_set_normalizers(CeeusHandle h, const float* host, int n_floats) {
  if (!h || !host) return fail(h, CEEUS_ERR_INVALID, "set_normalizers: null argument");
  if (n_floats != h->norm_floats) return fail(h, CEEUS_ERR_INVALID, "set_normalizers: expected " + std::to_string(h->norm_floats) + " floats");
  CU_CHECK(h, cudaMemcpy(h->d_norm, host, (size_t)n_floats * sizeof(float), cudaMemcpyHostToDevice));
  h->norms_set = true;
  return CEEUS_OK;
}

int ceeus_set_action_bounds(CeeusHandle h, const float* low_host, const float* high_host) {
  if (!h || !low_host || !high_host) return fail(h, CEEUS_ERR_INVALID, "set_action_bounds: null argument");
  CU_CHECK(h, cudaMemcpy(h->d_low, low_host, h->md.U * sizeof(float), cudaMemcpyHostToDevice));
  CU_CHECK(h, cudaMemcpy(h->d_high, high_host, h->md.U * sizeof(float), cudaMemcpyHostToDevice));
  return CEEUS_OK;
}

int ceeus_set_goal(CeeusHandle h, const float* goal_host) {
  if (!h || !goal_host) return fail(h, CEEUS_ERR_INVALID, "set_goal: null argument");
  if (h->md.G == 0) return CEEUS_OK;
  CU_CHECK(h, cudaMemcpy(h->d_goal, goal_host, (size_t)h->cfg.num_envs * h->md.G * sizeof(float), cudaMemcpyHostToDevice));
  return CEEUS_OK;
}

int ceeus_bind_buffers(CeeusHandle h, const CeeusBuffers* b) {
  if (!h || !b) return fail(h, CEEUS_ERR_INVALID, "bind_buffers: null argument");
  if (!b->mean || !b->std || !b->actions || !b->costs_per_model || !b->costs || !b->elite_actions ||
      !b->elite_costs || !b->elite_cpm || !b->elite_idx || !b->cand || !b->action_out)
    return fail(h, CEEUS_ERR_INVALID, "bind_buffers: every buffer except traj must be non-null");
  if (!b->traj && (!h->fuse_costs || h->cd.env_kind == CEEUS_ENV_EXTERNAL))
    return fail(h, CEEUS_ERR_INVALID, "bind_buffers: traj is required when the task cost is evaluated by the caller "
                                      "(CEEUS_ENV_EXTERNAL) or the fused cost tail is disabled");
  h->bufs = *b;
  h->bound = true;
  drop_graphs(h);  // captured launches hold the old pointers
  return CEEUS_OK;
}

int ceeus_cand_record_len(CeeusHandle h) {
  return h ? 2 + h->cfg.ensemble_size + h->cfg.horizon * h->cfg.action_dim : CEEUS_ERR_INVALID;
}

static void fill_pred_layout(CeeusHandle h, costdev::FusedCost& fc) {
  const long long n_plan = (long long)h->cfg.num_envs * h->cfg.num_traj;
  fc.pred = h->d_pred; fc.Op = h->pred_Op; fc.pstride = h->pred_stride;
  fc.Op_magic = (unsigned)((1ull << 32) / (unsigned)h->pred_Op) + 1u;
  if (h->pred_stride) { fc.ps = h->pred_stride; fc.hs = (long long)h->md.E * h->pred_Op; fc.es = h->pred_Op; }
  else { fc.ps = h->pred_Op; fc.hs = (long long)h->md.E * n_plan * h->pred_Op; fc.es = n_plan * h->pred_Op; }
  fc.cpm = h->bufs.costs_per_model; fc.costs = h->bufs.costs;
  fc.dims = costdev::CostDims{h->cfg.horizon, h->md.E, h->md.O, h->md.N, h->md.A, h->md.D, h->md.S, h->md.G, h->md.sd, h->pred_Op};
  fc.cd = h->cd;
}

static int ensure_plan_scratch(CeeusHandle h) {
  if (h->d_pred) return CEEUS_OK;
  if (h->capturing) return fail(h, CEEUS_ERR_STATE, "planner scratch must be allocated before graph capture");
  const size_t n_plan = (size_t)h->cfg.num_envs * h->cfg.num_traj;
  const size_t per_traj = h->pred_stride ? (size_t)h->pred_stride : (size_t)h->cfg.horizon * h->md.E * h->pred_Op;
  CU_CHECK(h, cudaMalloc(&h->d_pred, n_plan * per_traj * sizeof(float)));
  CU_CHECK(h, cudaMalloc(&h->d_done, n_plan * sizeof(int)));
  CU_CHECK(h, cudaMemset(h->d_done, 0, n_plan * sizeof(int)));
  return CEEUS_OK;
}

// fused != 0: planner launch -- predictions go to the handle's compact scratch and every trajectory's cost is reduced in the tail of
// the rollout kernel by the warp that completes its last ensemble member (cost_device.cuh); traj_dev is not written.
static int do_rollout(CeeusHandle h, const float* start_obs_dev, int traj_per_start, const float* actions_dev, int n,
                      int traj_per_goal, float* traj_dev, cudaStream_t st, bool fused = false) {
  if (!h->weights_set || !h->norms_set) return fail(h, CEEUS_ERR_STATE, "rollout before set_weights / set_normalizers");
  RolloutArgs ra{};
  ra.weights = h->d_weights; ra.norm = h->d_norm; ra.start_obs = start_obs_dev; ra.goal = h->d_goal;
  ra.actions = actions_dev; ra.traj = traj_dev; ra.n = n; ra.H = h->cfg.horizon;
  ra.traj_per_start = traj_per_start; ra.traj_per_goal = traj_per_goal;
  costdev::FusedCost fc{};
  if (fused) {
    const int n_plan = h->cfg.num_envs * h->cfg.num_traj;
    if (n != n_plan) return fail(h, CEEUS_ERR_INVALID, "fused rollout: n must be num_envs * num_traj");
    if (int rc = ensure_plan_scratch(h)) return rc;
    fill_pred_layout(h, fc);
    fc.done = h->cost_tail ? h->d_done : nullptr;  // null: the costs are reduced by launch_costs_compact after the rollout
    h->last_plan_obs = start_obs_dev;
    static const int tail_dbg = std::getenv("CEEUS_TAIL_DEBUG") ? std::atoi(std::getenv("CEEUS_TAIL_DEBUG")) : 0;
    fc.dbg_mode = tail_dbg;
    fc.discard = h->l2_discard ? 1 : 0;
  } else if (!traj_dev) {
    return fail(h, CEEUS_ERR_INVALID, "rollout: no trajectory buffer");
  }
  const bool prof = h->prof_on && !h->capturing && 2 * h->prof_n + 2 <= (int)h->prof_ev.size();
  if (prof) CU_CHECK(h, cudaEventRecord(h->prof_ev[2 * h->prof_n], st));
  struct ProfStop {  // the closing event is recorded on every exit path
    CeeusHandle h; cudaStream_t st; bool on;
    ~ProfStop() { if (on) { cudaEventRecord(h->prof_ev[2 * h->prof_n + 1], st); h->prof_n += 1; } }
  } prof_stop{h, st, prof};
  if (h->cfg.precision == CEEUS_PREC_TC && h->md.kind == CEEUS_MODEL_MLP)
    LAUNCH(h, launch_rollout_mlp_tc(h->md, ra, h->tcm, h->d_wimg, h->d_tc_err, fc, st));
  else if (h->cfg.precision == CEEUS_PREC_TC)
    LAUNCH(h, launch_rollout_gnn_tc(h->md, ra, h->tc, h->d_wimg, h->d_wpar, h->d_tc_err, h->d_tc_dbg, h->tc_all_fold, fc, st));
  else if (h->md.kind == CEEUS_MODEL_GNN) LAUNCH(h, launch_rollout_gnn_simt(h->md, ra, h->num_sms, fc, st));
  else LAUNCH(h, launch_rollout_mlp_simt(h->md, ra, fc, st));
  return CEEUS_OK;
}

static bool plan_fuses_costs(CeeusHandle h) {
  const bool need_env = !h->cd.curiosity || h->cd.extrinsic;
  return h->fuse_costs && !(need_env && h->cd.env_kind == CEEUS_ENV_EXTERNAL);
}

int ceeus_rollout(CeeusHandle h, const float* start_obs_dev, int n_start, const float* actions_dev, int n, float* traj_dev,
                  void* stream) {
  if (!h || !start_obs_dev || !actions_dev || !traj_dev) return fail(h, CEEUS_ERR_INVALID, "rollout: null argument");
  if (n < 1 || n_start < 1 || n % n_start != 0 || n % h->cfg.num_envs != 0)
    return fail(h, CEEUS_ERR_INVALID, "rollout: n must be a multiple of n_start and of num_envs");
  return do_rollout(h, start_obs_dev, n / n_start, actions_dev, n, n / h->cfg.num_envs, traj_dev, S(stream));
}

static int do_costs(CeeusHandle h, const float* traj_dev, int n, float* cpm, float* costs, cudaStream_t st,
                    const float* env_cost_ext = nullptr, bool own_rollout = false) {
  const bool need_env = !h->cd.curiosity || h->cd.extrinsic;
  if (need_env && h->cd.env_kind == CEEUS_ENV_EXTERNAL && env_cost_ext == nullptr)
    return fail(h, CEEUS_ERR_STATE, "this environment's task cost is evaluated by the caller: use ceeus_plan_iter_rollout + "
                                    "ceeus_plan_iter_costs(env_cost_dev)");
  CostArgs ca{};
  ca.env_cost_ext = env_cost_ext;
  // GNN: static features and goal are copied into every member's prediction; dense MLP: it predicts the whole state, only the goal is copied
  ca.Ostd = !own_rollout ? h->md.O : h->md.kind == CEEUS_MODEL_GNN ? h->md.A + h->md.N * h->md.D : h->md.sd;
  ca.traj = traj_dev; ca.cpm = cpm; ca.costs = costs; ca.n = n; ca.H = h->cfg.horizon; ca.E = h->md.E; ca.O = h->md.O;
  ca.N = h->md.N; ca.A = h->md.A; ca.D = h->md.D; ca.S = h->md.S; ca.G = h->md.G; ca.sd = h->md.sd;
  LAUNCH(h, launch_costs(ca, h->cd, st));
  return CEEUS_OK;
}

int ceeus_costs(CeeusHandle h, const float* traj_dev, int n, float* costs_per_model_dev, float* costs_dev, void* stream) {
  if (!h || !traj_dev || !costs_per_model_dev || !costs_dev || n < 1) return fail(h, CEEUS_ERR_INVALID, "costs: bad argument");
  return do_costs(h, traj_dev, n, costs_per_model_dev, costs_dev, S(stream));
}

int ceeus_sample(CeeusHandle h, const float* mean_dev, const float* std_dev, const float* noise_dev, float* gauss_dev,
                 int write_gauss, uint32_t stream_id, int64_t first_row, int n, float* actions_dev, void* stream) {
  if (!h || !mean_dev || !std_dev || !actions_dev || n < 1 || n % h->cfg.num_envs != 0)
    return fail(h, CEEUS_ERR_INVALID, "sample: bad argument");
  SampleArgs a{};
  a.mean = mean_dev; a.std = std_dev; a.low = h->d_low; a.high = h->d_high; a.noise = noise_dev; a.gauss = gauss_dev;
  a.write_gauss = write_gauss; a.irdft = h->d_irdft; a.actions = actions_dev; a.n = n; a.H = h->cfg.horizon;
  a.U = h->md.U; a.F = h->F; a.rows_per_env = n / h->cfg.num_envs; a.colored = h->cfg.colored_noise;
  a.seed = h->cfg.seed; a.stream_id = stream_id; a.step_ctr = nullptr; a.streams_per_step = 0; a.first_row = first_row;
  a.global_rows_per_env = (long long)(n / h->cfg.num_envs) * h->cfg.world_size;
  LAUNCH(h, launch_sample(a, S(stream)));
  return CEEUS_OK;
}

static FinishArgs finish_args(CeeusHandle h) {
  FinishArgs f{};
  f.mean = h->bufs.mean; f.std = h->bufs.std; f.elite_actions = h->bufs.elite_actions; f.low = h->d_low; f.high = h->d_high;
  f.action_out = h->bufs.action_out; f.nE = h->cfg.num_envs; f.k = h->cfg.num_elites; f.H = h->cfg.horizon; f.U = h->md.U;
  f.execute_best = h->cfg.execute_best_elite; f.relative_init = h->cfg.relative_init; f.init_std = h->cfg.init_std;
  f.step_ctr = h->d_step_ctr;
  return f;
}

int ceeus_begin_rollout(CeeusHandle h, void* stream) {
  if (!h) return CEEUS_ERR_INVALID;
  if (!h->bound) return fail(h, CEEUS_ERR_STATE, "begin_rollout before bind_buffers");
  LAUNCH(h, launch_reset_dist(finish_args(h), S(stream)));
  h->has_elites = false;
  return CEEUS_OK;
}

int ceeus_has_elites(CeeusHandle h) { return h && h->has_elites ? 1 : 0; }

int ceeus_plan_iter_local(CeeusHandle h, int iter, const float* obs_dev, const float* noise_dev, const float* shift_noise_dev,
                          void* stream) {
  if (int rc = ceeus_plan_iter_rollout(h, iter, obs_dev, noise_dev, shift_noise_dev, stream)) return rc;
  return ceeus_plan_iter_costs(h, nullptr, stream);
}

int ceeus_plan_iter_rollout(CeeusHandle h, int iter, const float* obs_dev, const float* noise_dev, const float* shift_noise_dev,
                            void* stream) {
  if (!h || !obs_dev) return fail(h, CEEUS_ERR_INVALID, "plan_iter_local: null argument");
  if (!h->bound) return fail(h, CEEUS_ERR_STATE, "plan before bind_buffers");
  const CeeusConfig& c = h->cfg;
  if (iter < 0 || iter >= c.opt_iterations) return fail(h, CEEUS_ERR_INVALID, "plan_iter_local: iter out of range");
  cudaStream_t st = S(stream);
  const int n = c.num_envs * c.num_traj;
  const uint32_t streams_per_step = (uint32_t)c.opt_iterations + 1u;  // one Philox stream per iteration + one for shifted elites
  // 1. sample_action_sequences (mpc_torch.py:204-236)
  SampleArgs a{};
  a.mean = h->bufs.mean; a.std = h->bufs.std; a.low = h->d_low; a.high = h->d_high; a.noise = noise_dev; a.gauss = nullptr;
  a.write_gauss = 0; a.irdft = h->d_irdft; a.actions = h->bufs.actions; a.n = n; a.H = c.horizon; a.U = h->md.U; a.F = h->F;
  a.rows_per_env = c.num_traj; a.colored = c.colored_noise; a.seed = c.seed; a.stream_id = (uint32_t)iter;
  a.step_ctr = h->d_step_ctr; a.streams_per_step = streams_per_step;
  a.first_row = (long long)c.rank * c.num_traj; a.global_rows_per_env = (long long)c.num_traj * c.world_size;
  LAUNCH(h, launch_sample(a, st));
  // 2. mean-action row + elites shifted over time live at global rows 0..n_kept-1 => rank 0
  if (c.rank == 0) {
    FixupArgs f{};
    f.actions = h->bufs.actions; f.mean = h->bufs.mean; f.std = h->bufs.std; f.low = h->d_low; f.high = h->d_high;
    f.elite_actions = h->bufs.elite_actions; f.shift_noise = shift_noise_dev; f.irdft = h->d_irdft;
    f.nE = c.num_envs; f.P = c.num_traj; f.H = c.horizon; f.U = h->md.U; f.F = h->F; f.k = c.num_elites; f.n_kept = c.num_kept;
    f.colored = c.colored_noise;
    f.write_mean_row = c.use_mean_actions && iter == c.opt_iterations - 1;
    f.shift_elites = iter == 0 && c.shift_elites_over_time && h->has_elites && c.num_kept > 0;
    f.seed = c.seed; f.stream_id = (uint32_t)c.opt_iterations; f.step_ctr = h->d_step_ctr; f.streams_per_step = streams_per_step;
    if (f.write_mean_row || f.shift_elites) LAUNCH(h, launch_fixup(f, st));
  }
  // 3. predict_n_steps (+ trajectory costs in the kernel's tail unless the caller evaluates the task cost)
  return do_rollout(h, obs_dev, c.num_traj, h->bufs.actions, n, c.num_traj, h->bufs.traj, st, plan_fuses_costs(h));
}

int ceeus_plan_iter_costs(CeeusHandle h, const float* env_cost_dev, void* stream) {
  if (!h) return CEEUS_ERR_INVALID;
  if (!h->bound) return fail(h, CEEUS_ERR_STATE, "plan before bind_buffers");
  const CeeusConfig& c = h->cfg;
  cudaStream_t st = S(stream);
  const int n = c.num_envs * c.num_traj;
  // 4. trajectory costs.  Built-in costs are reduced from the compact prediction scratch: in the tail of the rollout kernel
  // (cost_in_rollout_tail) or by one warp per trajectory here; a task cost evaluated by the caller needs the full trajectories.
  if (!plan_fuses_costs(h)) {
    if (int rc = do_costs(h, h->bufs.traj, n, h->bufs.costs_per_model, h->bufs.costs, st, env_cost_dev, true)) return rc;
  } else if (!h->cost_tail) {
    costdev::FusedCost fc{};
    fill_pred_layout(h, fc);
    if (!h->last_plan_obs) return fail(h, CEEUS_ERR_STATE, "plan_iter_costs before plan_iter_rollout");
    LAUNCH(h, launch_costs_compact(fc, h->last_plan_obs, h->d_goal, c.num_traj, c.num_traj, n, st));
  }
  // 5. this rank's elite candidates
  TopkArgs t{};
  t.costs = h->bufs.costs; t.cpm = h->bufs.costs_per_model; t.actions = h->bufs.actions; t.cand = h->bufs.cand;
  t.nE = c.num_envs; t.P = c.num_traj; t.E = h->md.E; t.HU = c.horizon * h->md.U; t.k = c.num_elites;
  t.rec = ceeus_cand_record_len(h); t.first_index = (long long)c.rank * c.num_traj;
  LAUNCH(h, launch_topk(t, st));
  return CEEUS_OK;
}

static int plan_iter_update_impl(CeeusHandle h, int iter, const float* cand_all_dev, void* stream, bool p2p);

int ceeus_plan_iter_update(CeeusHandle h, int iter, const float* cand_all_dev, void* stream) {
  return plan_iter_update_impl(h, iter, cand_all_dev, stream, false);
}

static int plan_iter_update_impl(CeeusHandle h, int iter, const float* cand_all_dev, void* stream, bool p2p) {
  if (!h || !cand_all_dev) return fail(h, CEEUS_ERR_INVALID, "plan_iter_update: null argument");
  if (!h->bound) return fail(h, CEEUS_ERR_STATE, "plan before bind_buffers");
  const CeeusConfig& c = h->cfg;
  UpdateArgs u{};
  u.cand_all = cand_all_dev; u.elite_actions = h->bufs.elite_actions; u.elite_costs = h->bufs.elite_costs;
  u.elite_cpm = h->bufs.elite_cpm; u.elite_idx = h->bufs.elite_idx; u.mean = h->bufs.mean; u.std = h->bufs.std;
  u.world = c.world_size; u.nE = c.num_envs; u.k = c.num_elites; u.E = h->md.E; u.H = c.horizon; u.U = h->md.U;
  u.rec = ceeus_cand_record_len(h);
  u.n_kept_in = (iter > 0 && c.keep_previous_elites) ? c.num_kept : 0;  // mpc_torch.py:319-321
  u.pop_fresh = (long long)c.num_traj * c.world_size;
  u.alpha = c.alpha;
  if (p2p) {
    u.step_ctr = h->d_step_ctr; u.flags = h->p2p_flags[c.rank]; u.iters_per_step = c.opt_iterations; u.iter = iter;
    u.half_stride = (long long)c.world_size * c.num_envs * c.num_elites * u.rec;
    u.err = h->d_tc_err;
  }
  LAUNCH(h, launch_update(u, S(stream)));
  return CEEUS_OK;
}

int ceeus_plan_finish(CeeusHandle h, void* stream) {
  if (!h) return CEEUS_ERR_INVALID;
  if (!h->bound) return fail(h, CEEUS_ERR_STATE, "plan before bind_buffers");
  LAUNCH(h, launch_finish(finish_args(h), S(stream)));
  if (!h->capturing) h->has_elites = true;
  return CEEUS_OK;
}

// The launch sequence of one MPC step on `stream` (shared by the eager path and the graph capture).
// resident: the observation is already in bufs.obs and the action stays in bufs.action_out (no host copies).
static int enqueue_plan_step(CeeusHandle h, const float* noise_dev_or_null, void* stream, bool resident) {
  const CeeusConfig& c = h->cfg;
  cudaStream_t st = S(stream);
  const float* obs_dev = h->d_obs;
  if (resident) {
    obs_dev = h->bufs.obs;
  } else {
    const size_t ob = (size_t)c.num_envs * h->md.O * sizeof(float);
    CU_CHECK(h, cudaMemcpyAsync(h->d_obs, h->h_obs, ob, cudaMemcpyHostToDevice, st));
  }
  const size_t per_iter = (size_t)c.num_envs * c.num_traj * h->md.U * c.horizon;
  const size_t per_shift = (size_t)c.num_envs * c.num_kept * h->md.U * c.horizon;
  const float* np = noise_dev_or_null;
  for (int i = 0; i < c.opt_iterations; ++i) {
    const float* ni = np;
    const float* ns = nullptr;
    if (np) {
      np += per_iter;
      if (i == 0 && c.shift_elites_over_time && h->has_elites && c.num_kept > 0) {
        ns = np;
        np += per_shift;
      }
    }
    if (int rc = ceeus_plan_iter_local(h, i, obs_dev, ni, ns, stream)) return rc;
    const float* cand_all = h->bufs.cand;
    if (c.world_size > 1 && h->p2p_ready) {
      // the ranks' k best candidates -> every rank's gather buffer by direct NVLink stores + a flag (no NCCL launch); the merge
      // kernel waits for the flags of this exchange and reads the matching half of its double buffer
      ExchangeArgs x{};
      x.cand = h->bufs.cand; x.world = c.world_size; x.rank = c.rank;
      x.count = c.num_envs * c.num_elites * ceeus_cand_record_len(h);
      x.half_stride = (long long)c.world_size * x.count;
      x.step_ctr = h->d_step_ctr; x.iters_per_step = c.opt_iterations; x.iter = i;
      for (int r = 0; r < c.world_size; ++r) { x.peer_buf[r] = h->p2p_buf[r]; x.peer_flags[r] = h->p2p_flags[r]; }
      LAUNCH(h, launch_exchange(x, st));
      if (int rc = plan_iter_update_impl(h, i, h->p2p_buf[c.rank], stream, true)) return rc;
      continue;
    }
    if (c.world_size > 1) {
      // the ranks' k best candidates -> every rank (a few KB, latency bound); captured into the step's CUDA graph like any kernel
      const size_t cnt = (size_t)c.num_envs * c.num_elites * ceeus_cand_record_len(h);
      const int nr = nccl().AllGather(h->bufs.cand, h->d_cand_all, cnt, kNcclFloat32, h->comm, st);
      if (nr != 0) return fail(h, CEEUS_ERR_CUDA, std::string("ncclAllGather: ") + nccl().GetErrorString(nr));
      h->launches += 1;
      cand_all = h->d_cand_all;
    }
    if (int rc = ceeus_plan_iter_update(h, i, cand_all, stream)) return rc;
  }
  if (int rc = ceeus_plan_finish(h, stream)) return rc;
  if (!resident) {
    const size_t ab = (size_t)c.num_envs * h->md.U * sizeof(float);
    CU_CHECK(h, cudaMemcpyAsync(h->h_action, h->bufs.action_out, ab, cudaMemcpyDeviceToHost, st));
  }
  if (h->d_tc_err) CU_CHECK(h, cudaMemcpyAsync(h->h_tc_err, h->d_tc_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  return CEEUS_OK;
}

// One MPC step as ONE CUDA-graph launch on the handle's own stream (device RNG).  Nothing in the sequence depends on host state
// except "are there elites to shift" (two graphs per variant); the Philox stream base is a device counter.
static int launch_plan_graph(CeeusHandle h, cudaStream_t caller, bool resident, bool sync) {
  const int v = h->has_elites ? 1 : 0, r = resident ? 1 : 0;
  if (plan_fuses_costs(h))
    if (int rc = ensure_plan_scratch(h)) return rc;
  if (!h->graph_exec[v][r]) {
    cudaGraph_t g = nullptr;
    const int64_t l0 = h->launches;
    h->capturing = true;
    CU_CHECK(h, cudaStreamBeginCapture(h->gstream, cudaStreamCaptureModeRelaxed));
    const int rc = enqueue_plan_step(h, nullptr, h->gstream, resident);
    const cudaError_t ce = cudaStreamEndCapture(h->gstream, &g);
    h->capturing = false;
    h->graph_launches[v][r] = h->launches - l0;
    h->launches = l0;
    if (rc != CEEUS_OK) { if (g) cudaGraphDestroy(g); return rc; }
    CU_CHECK(h, ce);
    const cudaError_t ie = cudaGraphInstantiate(&h->graph_exec[v][r], g, 0);
    cudaGraphDestroy(g);
    CU_CHECK(h, ie);
  }
  CU_CHECK(h, cudaEventRecord(h->gevent, caller));  // work already queued on the caller's stream (weight repack, obs copy) comes first
  CU_CHECK(h, cudaStreamWaitEvent(h->gstream, h->gevent, 0));
  CU_CHECK(h, cudaGraphLaunch(h->graph_exec[v][r], h->gstream));
  h->launches += h->graph_launches[v][r];
  h->has_elites = true;
  if (sync) {
    CU_CHECK(h, cudaStreamSynchronize(h->gstream));
  } else {  // the caller's stream continues after the step
    CU_CHECK(h, cudaEventRecord(h->gevent, h->gstream));
    CU_CHECK(h, cudaStreamWaitEvent(caller, h->gevent, 0));
  }
  return CEEUS_OK;
}

static int plan_step_checks(CeeusHandle h) {
  if (!h->bound) return fail(h, CEEUS_ERR_STATE, "plan before bind_buffers");
  if (h->cfg.world_size != 1 && !h->comm && !h->p2p_ready)
    return fail(h, CEEUS_ERR_STATE, "world_size > 1: call ceeus_p2p_export / _import or ceeus_comm_init first (or drive plan_iter_local / plan_iter_update "
                                    "yourself with an all-gather of the candidate records in between)");
  const bool need_env = !h->cd.curiosity || h->cd.extrinsic;
  if (need_env && h->cd.env_kind == CEEUS_ENV_EXTERNAL)
    return fail(h, CEEUS_ERR_STATE, "this environment's task cost is evaluated by the caller: use ceeus_plan_iter_rollout / _costs");
  return CEEUS_OK;
}

// status word of the tensor-core kernels: protocol error code (bounded mbarrier wait timed out) and / or CEEUS_TC_RANGE
static int tc_error(CeeusHandle h, int st) {
  if (st & ~CEEUS_TC_RANGE)
    return fail(h, CEEUS_ERR_CUDA, "tensor-core rollout kernel reported protocol error " + std::to_string(st & ~CEEUS_TC_RANGE));
  return fail(h, CEEUS_ERR_STATE,
              "CEEUS_PREC_TC: an fp16 operand (normalised observation / action, or a hidden activation of a model without LayerNorm) "
              "exceeded +-65504 and was saturated; the plan is unreliable -- use precision fp32 for this model / these states");
}

int ceeus_plan_step(CeeusHandle h, const float* obs_host, float* action_host, const float* noise_dev_or_null, void* stream) {
  if (!h || !obs_host || !action_host) return fail(h, CEEUS_ERR_INVALID, "plan_step: null argument");
  if (int rc = plan_step_checks(h)) return rc;
  const CeeusConfig& c = h->cfg;
  cudaStream_t st = S(stream);
  std::memcpy(h->h_obs, obs_host, (size_t)c.num_envs * h->md.O * sizeof(float));
  static const bool no_graph = std::getenv("CEEUS_NO_GRAPH") != nullptr;
  if (noise_dev_or_null == nullptr && !no_graph) {
    if (int rc = launch_plan_graph(h, st, false, true)) return rc;
  } else {
    if (int rc = enqueue_plan_step(h, noise_dev_or_null, stream, false)) return rc;
    if (!h->capturing) h->has_elites = true;
    CU_CHECK(h, cudaStreamSynchronize(st));
  }
  if (h->d_tc_err && *h->h_tc_err != 0) return tc_error(h, *h->h_tc_err);
  std::memcpy(action_host, h->h_action, (size_t)c.num_envs * h->md.U * sizeof(float));
  return CEEUS_OK;
}

int ceeus_plan_step_resident(CeeusHandle h, const float* noise_dev_or_null, void* stream) {
  if (!h) return CEEUS_ERR_INVALID;
  if (int rc = plan_step_checks(h)) return rc;
  if (!h->bufs.obs) return fail(h, CEEUS_ERR_STATE, "plan_step_resident: no observation buffer bound (CeeusBuffers.obs)");
  static const bool no_graph = std::getenv("CEEUS_NO_GRAPH") != nullptr;
  // (with ceeus_profile_rollouts on, the step is enqueued kernel by kernel so that the event pairs around the rollouts exist)
  if (noise_dev_or_null == nullptr && !no_graph && !h->prof_on) return launch_plan_graph(h, S(stream), true, false);
  if (int rc = enqueue_plan_step(h, noise_dev_or_null, stream, true)) return rc;
  h->has_elites = true;
  return CEEUS_OK;
}

// ---- NVLink peer-memory exchange: buffer export / import through CUDA IPC ----
static size_t p2p_bytes(CeeusHandle h) {
  const size_t cnt = (size_t)h->cfg.world_size * h->cfg.num_envs * h->cfg.num_elites * ceeus_cand_record_len(h);
  return 2 * cnt * sizeof(float) + 8 * sizeof(unsigned long long);
}

int ceeus_p2p_export(CeeusHandle h, uint8_t* handle_out) {
  if (!h || !handle_out) return fail(h, CEEUS_ERR_INVALID, "p2p_export: null argument");
  if (h->cfg.world_size < 2 || h->cfg.world_size > 8) return fail(h, CEEUS_ERR_INVALID, "p2p exchange needs 2..8 ranks");
  if (!h->p2p_local) {
    CU_CHECK(h, cudaMalloc(&h->p2p_local, p2p_bytes(h)));
    CU_CHECK(h, cudaMemset(h->p2p_local, 0, p2p_bytes(h)));
  }
  cudaIpcMemHandle_t mh;
  CU_CHECK(h, cudaIpcGetMemHandle(&mh, h->p2p_local));
  static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(handle_out, &mh, sizeof(mh));
  return CEEUS_OK;
}

int ceeus_p2p_import(CeeusHandle h, const uint8_t* handles) {
  if (!h || !handles) return fail(h, CEEUS_ERR_INVALID, "p2p_import: null argument");
  if (!h->p2p_local) return fail(h, CEEUS_ERR_STATE, "p2p_import before p2p_export");
  const size_t flags_off = p2p_bytes(h) - 8 * sizeof(unsigned long long);
  for (int r = 0; r < h->cfg.world_size; ++r) {
    void* base = h->p2p_local;
    if (r != h->cfg.rank) {
      cudaIpcMemHandle_t mh;
      std::memcpy(&mh, handles + (size_t)r * 64, sizeof(mh));
      CU_CHECK(h, cudaIpcOpenMemHandle(&base, mh, cudaIpcMemLazyEnablePeerAccess));
    }
    h->p2p_buf[r] = reinterpret_cast<float*>(base);
    h->p2p_flags[r] = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(base) + flags_off);
  }
  if (!h->d_tc_err) {
    CU_CHECK(h, cudaMalloc(&h->d_tc_err, sizeof(int)));
    CU_CHECK(h, cudaMemset(h->d_tc_err, 0, sizeof(int)));
  }
  h->p2p_ready = true;
  drop_graphs(h);
  return CEEUS_OK;
}

int ceeus_comm_unique_id(uint8_t* id_out) {
  if (!id_out) return fail(nullptr, CEEUS_ERR_INVALID, "comm_unique_id: null argument");
  if (!nccl().ok) return fail(nullptr, CEEUS_ERR_CUDA, nccl().why);
  NcclApi::UniqueId id;
  const int nr = nccl().GetUniqueId(&id);
  if (nr != 0) return fail(nullptr, CEEUS_ERR_CUDA, std::string("ncclGetUniqueId: ") + nccl().GetErrorString(nr));
  std::memcpy(id_out, id.internal, sizeof(id.internal));
  return CEEUS_OK;
}

int ceeus_comm_init(CeeusHandle h, const uint8_t* unique_id) {
  if (!h || !unique_id) return fail(h, CEEUS_ERR_INVALID, "comm_init: null argument");
  if (h->cfg.world_size < 2) return fail(h, CEEUS_ERR_INVALID, "comm_init: the handle was created with world_size 1");
  if (!nccl().ok) return fail(h, CEEUS_ERR_CUDA, nccl().why);
  if (h->comm) return CEEUS_OK;
  NcclApi::UniqueId id;
  std::memcpy(id.internal, unique_id, sizeof(id.internal));
  const int nr = nccl().CommInitRank(&h->comm, h->cfg.world_size, id, h->cfg.rank);
  if (nr != 0) { h->comm = nullptr; return fail(h, CEEUS_ERR_CUDA, std::string("ncclCommInitRank: ") + nccl().GetErrorString(nr)); }
  const size_t cnt = (size_t)h->cfg.world_size * h->cfg.num_envs * h->cfg.num_elites * ceeus_cand_record_len(h);
  CU_CHECK(h, cudaMalloc(&h->d_cand_all, cnt * sizeof(float)));
  drop_graphs(h);
  return CEEUS_OK;
}

int64_t ceeus_kernel_launches(CeeusHandle h) { return h ? h->launches : 0; }

int ceeus_profile_rollouts(CeeusHandle h, int enable, float* ms_out_host, int max_out) {
  if (!h) return CEEUS_ERR_INVALID;
  int n = 0;
  if (ms_out_host && h->prof_n > 0) {
    CU_CHECK(h, cudaEventSynchronize(h->prof_ev[2 * h->prof_n - 1]));
    for (; n < h->prof_n && n < max_out; ++n)
      CU_CHECK(h, cudaEventElapsedTime(&ms_out_host[n], h->prof_ev[2 * n], h->prof_ev[2 * n + 1]));
  }
  h->prof_n = 0;
  if (enable && h->prof_ev.empty()) {
    h->prof_ev.resize(2 * 256);
    for (auto& e : h->prof_ev) CU_CHECK(h, cudaEventCreate(&e));
  }
  h->prof_on = enable != 0;
  return n;
}

#ifdef CEEUS_DEBUG_TOOLS
int ceeus_tc_debug_timeline(CeeusHandle h, int enable, long long* out_host, int n_pairs) {
  if (!h || h->cfg.precision != CEEUS_PREC_TC) return CEEUS_ERR_INVALID;
  drop_graphs(h);  // captured launches hold the kernel variant and the d_tc_dbg pointer
  if (enable && !h->d_tc_dbg) {
    if (cudaMalloc(&h->d_tc_dbg, 512 * sizeof(long long)) != cudaSuccess) return CEEUS_ERR_CUDA;
    cudaMemset(h->d_tc_dbg, 0, 512 * sizeof(long long));
  }
  if (out_host && h->d_tc_dbg) {
    if (n_pairs > 256) n_pairs = 256;
    if (cudaMemcpy(out_host, h->d_tc_dbg, (size_t)n_pairs * 2 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess)
      return CEEUS_ERR_CUDA;
  }
  if (!enable && h->d_tc_dbg) {
    cudaFree(h->d_tc_dbg);
    h->d_tc_dbg = nullptr;
  }
  return CEEUS_OK;
}

int ceeus_selftest_epilogue(int warps_per_scheduler, int with_bias, int with_mma, long long* cycles_host) {
  if (!cycles_host || warps_per_scheduler < 1 || warps_per_scheduler > 3) return CEEUS_ERR_INVALID;
  long long* d = nullptr;
  if (cudaMalloc(&d, 2 * sizeof(long long)) != cudaSuccess) return CEEUS_ERR_CUDA;
  cudaMemset(d, 0, 2 * sizeof(long long));
  cudaError_t e = launch_epi_bench(warps_per_scheduler, with_bias, with_mma, d);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(cycles_host, d, 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? CEEUS_OK : CEEUS_ERR_CUDA;
}

int ceeus_tc_debug_flag(CeeusHandle h, long long flag) {
  if (!h || !h->d_tc_dbg) return CEEUS_ERR_INVALID;
  return cudaMemcpy(h->d_tc_dbg + 499, &flag, sizeof(flag), cudaMemcpyHostToDevice) == cudaSuccess ? CEEUS_OK : CEEUS_ERR_CUDA;
}

#endif  // CEEUS_DEBUG_TOOLS

int ceeus_tc_tiling(CeeusHandle h, int n, int32_t* rounds_out, int32_t* uneven_out) {
  if (!h || n < 1 || !rounds_out || !uneven_out) return fail(h, CEEUS_ERR_INVALID, "tc_tiling: bad argument");
  if (h->cfg.precision != CEEUS_PREC_TC) return fail(h, CEEUS_ERR_INVALID, "tc_tiling: not a CEEUS_PREC_TC handle");
  const int E = h->md.E;
  for (int e = 0; e < E; ++e) { rounds_out[e] = 0; uneven_out[e] = 0; }
  const int pairs = h->md.kind == CEEUS_MODEL_MLP ? h->tcm.pairs : h->tc.pairs;
  for (int pair = 0; pair < pairs; ++pair) {
    if (h->md.kind == CEEUS_MODEL_MLP) {
      const TcTile t = tc_tile(pairs, E, pair, n, 1, 128, 32);  // one 128-trajectory tile per CTA (rollout_mlp_tc.cu)
      rounds_out[t.e] = t.rounds;
    } else {
      const TcTile t = tc_tile(pairs, E, pair, n, h->tc.NS, h->tc.Tmax, h->tc.tpw);
      rounds_out[t.e] = t.rounds;
      uneven_out[t.e] = t.uneven;
    }
  }
  return CEEUS_OK;
}

int ceeus_tc_row_layout(CeeusHandle h) {
  if (!h) return CEEUS_ERR_INVALID;
  if (h->cfg.precision != CEEUS_PREC_TC || h->md.kind != CEEUS_MODEL_GNN) return 0;
  return h->tc.grow ? 2 : 1;
}

int ceeus_tc_status(CeeusHandle h) {
  if (!h) return CEEUS_ERR_INVALID;
  if (!h->d_tc_err) return 0;
  int st = 0;
  if (cudaMemcpy(&st, h->d_tc_err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return CEEUS_ERR_CUDA;
  return st;
}

}  // extern "C"
