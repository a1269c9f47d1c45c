// Kernel launcher declarations shared between the translation units of libceeus_b200.
#pragma once
#include "common.cuh"

namespace ceeus {

struct RolloutArgs {
  const float* weights;    // packed [E][member_stride]
  const float* norm;       // normaliser buffer (NormDesc offsets)
  const float* start_obs;  // [n / traj_per_start, O]
  const float* goal;       // [n / traj_per_goal, G]
  const float* actions;    // [n, H, U]
  float* traj;             // [H+1, E, n, O]
  int n, H;
  int traj_per_start;      // trajectories sharing one start observation row
  int traj_per_goal;       // trajectories sharing one goal row (= per env)
};

struct GnnTiling {
  int G;     // trajectories per CTA group (G*N node rows <= 64)
  int TPT;   // whole trajectories per 64-row edge tile
  int XSe;   // row stride (floats) of the edge / global tile
  int XSn;   // row stride of the node tile
  size_t smem_bytes;
};

GnnTiling make_gnn_tiling(const ModelDesc& md);
cudaError_t launch_rollout_gnn_simt(const ModelDesc& md, const RolloutArgs& ra, int num_sms, cudaStream_t st);
cudaError_t launch_rollout_mlp_simt(const ModelDesc& md, const RolloutArgs& ra, cudaStream_t st);

// ---- tensor-core rollout (rollout_tc.cu) ----
struct TcMat {      // one B operand (fp16, K-major 8x16B core matrices, split by output column between the CTA pair)
  int K, N, nloc;   // padded K, total output columns, columns held per CTA
  int off;          // byte offset inside the per-(member, rank) weight image
  int src_mlp, src_layer;
  int seg_tc[2], seg_len[2], seg_src[2];  // K permutation: tile-K range [seg_tc, +len) <- source rows [seg_src, +len)
};
struct TcLayout {
  int HD, KE1, KX, T_r, ppm;  // ppm: CTA pairs per ensemble member
  TcMat mat[9];               // E1 E2 E3 N1 N2 N3 G1 G2 G3
  int image_bytes;            // per (member, rank)
  int par_off[9][3], par_len[9], par_floats;  // fp32 bias / gamma / beta blocks per member
  int sm_w, sm_slot[2], sm_par, sm_state[2], sm_rinfo, sm_bar, sraw_bytes, smem_bytes;
};
bool make_tc_layout(const ModelDesc& md, int num_sms, TcLayout* out);
cudaError_t launch_tc_pack(const float* w32, const ModelDesc& md, const TcLayout& tl, uint8_t* wimg, float* wpar, cudaStream_t st);
cudaError_t launch_rollout_gnn_tc(const ModelDesc& md, const RolloutArgs& ra, const TcLayout& tl, const uint8_t* wimg,
                                  const float* wpar, int* err_flag, long long* dbg, cudaStream_t st);

// ---- iCEM kernels (icem_kernels.cu) ----
struct SampleArgs {
  const float* mean;   // [nE,H,U]
  const float* std;    // [nE,H,U]
  const float* low;    // [U]
  const float* high;   // [U]
  const float* noise;  // [n,U,H] or null
  float* gauss;        // [n,U,2F] or null
  int write_gauss;
  const float* irdft;  // [2F,H] (colored) ; unused for white noise
  float* actions;      // [n,H,U]
  int n, H, U, F, rows_per_env, colored;
  unsigned long long seed;
  unsigned int stream_id;
  long long first_row;  // global index of local row 0 (RNG is keyed by the global row)
  long long global_rows_per_env;
};
cudaError_t launch_sample(const SampleArgs& a, cudaStream_t st);

struct FixupArgs {  // mean-action row and elites shifted over time (mpc_torch.py:238-267, 304-310)
  float* actions;            // [nE*P,H,U]
  const float* mean;         // [nE,H,U]
  const float* std;          // [nE,H,U]
  const float* low;
  const float* high;
  const float* elite_actions;  // [nE,k,H,U]
  const float* shift_noise;    // [nE*n_kept,U,H] or null
  const float* irdft;
  int nE, P, H, U, F, k, n_kept, colored;
  int write_mean_row, shift_elites;
  unsigned long long seed;
  unsigned int stream_id;
};
cudaError_t launch_fixup(const FixupArgs& a, cudaStream_t st);

struct CostArgs {
  const float* traj;  // [H+1,E,n,O]
  float* cpm;         // [n,E]
  float* costs;       // [n]
  int n, H, E, O, N, A, D, S, G, sd;
};
cudaError_t launch_costs(const CostArgs& a, const CostDesc& cd, cudaStream_t st);

struct TopkArgs {
  const float* costs;    // [nE,P]
  const float* cpm;      // [nE,P,E]
  const float* actions;  // [nE*P,H,U]
  float* cand;           // [nE,k,rec]
  int nE, P, E, HU, k, rec;
  long long first_index;  // global population index of local trajectory 0 (rank * P)
};
cudaError_t launch_topk(const TopkArgs& a, cudaStream_t st);

struct UpdateArgs {
  const float* cand_all;  // [world,nE,k,rec]
  float* elite_actions;   // [nE,k,H,U]
  float* elite_costs;     // [nE,k]
  float* elite_cpm;       // [nE,k,E]
  int* elite_idx;         // [nE,k]
  float* mean;            // [nE,H,U]
  float* std;             // [nE,H,U]
  int world, nE, k, E, H, U, rec, n_kept_in;  // n_kept_in: kept elites joining the population (0 if none)
  long long pop_fresh;                         // P * world: global index of the first kept elite
  float alpha;
};
cudaError_t launch_update(const UpdateArgs& a, cudaStream_t st);

struct FinishArgs {
  float* mean;
  float* std;
  const float* elite_actions;
  const float* low;
  const float* high;
  float* action_out;  // [nE,U]
  int nE, k, H, U, execute_best, relative_init;
  float init_std;
};
cudaError_t launch_finish(const FinishArgs& a, cudaStream_t st);
cudaError_t launch_reset_dist(const FinishArgs& a, cudaStream_t st);  // beginning_of_rollout: mean + std reset

cudaError_t launch_repack(const float* src, float* dst, int E, int rows, int cols, long long member_stride, int dst_off,
                          int dst_ld, cudaStream_t st);

}  // namespace ceeus
