// Kernel launcher declarations shared between the translation units of libceeus_b200.
#pragma once
#include "common.cuh"

namespace ceeus {

namespace costdev { struct FusedCost; }  // cost_device.cuh: compact prediction scratch + arrival counters of a planner launch

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
cudaError_t launch_rollout_gnn_simt(const ModelDesc& md, const RolloutArgs& ra, int num_sms, const costdev::FusedCost& fc, cudaStream_t st);
cudaError_t launch_rollout_mlp_simt(const ModelDesc& md, const RolloutArgs& ra, const costdev::FusedCost& fc, cudaStream_t st);

// ---- tensor-core rollout (rollout_tc.cu) ----
// Stage matrices of one model step, in issue order per edge pass / node tile / global tile.
enum { TC_E1 = 0, TC_E2, TC_N1, TC_N2, TC_N3, TC_G1, TC_G2, TC_G3, TC_NMAT };
struct TcMat {      // one B operand (fp16, K-major 8x16B core matrices, split by output column between the CTA pair)
  int K, N, nloc;   // padded K, output columns, columns held per CTA
  int off;          // byte offset inside the per-(member, rank) weight image
  int kx;           // leading K range fed from the "x part" of the A operand (R1); the rest comes from R2 (aggregate)
  int ln;           // LayerNorm follows (weights are centred over the output columns at pack time)
  int par_bias, par_gamma, par_beta;  // float offsets in the per-member parameter block (-1: none / folded into K)
  int kmap_off;     // offset of this matrix' K map in the kmap buffer
};
struct TcLayout {
  int HD;
  int AUp, NAp, SNs;       // halves: header [agent|action|1|0..], node block [dyn|stat|0..], per-trajectory snrm row
  int KE1, KXn, KXg;       // padded K of the edge input, node x-part, global x-part
  int tpw, Tmax;           // whole trajectories per warp, per group (4 warps)
  int tail;                // N*S + G floats behind the dynamic part of an observation
  TcMat mat[TC_NMAT];
  int image_bytes;         // per (member, rank)
  int par_floats;          // per member
  int par_fold;            // offset of the "LayerNorm affine folded into the weights" flag (1.0 / 0.0)
  int kmap_len;
  int NS;                  // warpgroups (slots) per CTA: 2 at Hd = 128, 4 at Hd = 64
  int sm_w, sm_par, sm_nrm, sm_out, nrm_floats, sm_slot[4], sm_bar, smem_bytes;
  int so_obs, obs_bytes, so_glob, so_snrm, so_tail, so_actn;  // offsets inside a slot region; bytes of one warp's observation staging
  int pairs;               // CTA pairs in the grid
  int scratch_floats;      // per member: logical fp32 matrices built by the pack kernels
  // "global row" tiling (grow = 1): a trajectory owns N node rows + 1 global row of the tile, the node and the global MLP run as
  // ONE K-stacked stage each (NG1, NG2, NG3: 2(N-1) + 3 stages per model step instead of 2(N-1) + 6).  The raw-input K blocks
  // ([hdr | x_i] of the node rows, [hdr] of the global rows) and the constant-one bias blocks are shared-memory A operands.
  int grow;
  int so_xa;               // inside a slot region: A tile [6 chunks][128 rows][16 B]: hdr_N(2) | x_i(2) | hdr_G(2)
  int sm_one;              // per CTA: constant A tile [4 chunks][128 rows][16 B]: one-hot(node rows)(2) | one-hot(global rows)(2)
  int sm_neg;              // per CTA: HD floats of -60000 (LayerNorm beta that zeroes a row through the ReLU)
  int even_rounds;         // 1: split a worker's trajectories evenly over its rounds (A/B switch; default: whole warps first)
};
// How the n trajectories of one launch are dealt to the CTA pairs / slots of one ensemble member (shared by the kernel and by
// ceeus_tc_tiling, so the parity tests can assert which tiling branch a case exercises).
struct TcTile {
  int e, lp, np;        // member, index of the pair inside the member, pairs of the member
  int per_worker, rounds, T;  // trajectories per worker (= slot), groups per worker, trajectories per group (<= Tmax)
  int q_un;             // uneven split: trajectories per warp and round
  int uneven;           // 1: in every round the leader-CTA slots get 3 warps' worth, the other CTA's slots 4 (the issuing warp stays
                        // empty); 2: the first k_pure rounds do that with whole warps, the remaining rounds split the rest evenly
  int k_pure, per_last; // uneven == 2: leading "pure issuer" rounds; trajectories per worker in each of the remaining rounds
};
__host__ __device__ inline TcTile tc_tile(int pairs, int E, int pair, int n, int NS, int Tmax, int tpw) {
  TcTile t;
  const int base = pairs / E, rem = pairs % E;
  if (pair < rem * (base + 1)) { t.e = pair / (base + 1); t.lp = pair % (base + 1); t.np = base + 1; }
  else { const int q = pair - rem * (base + 1); t.e = rem + q / base; t.lp = q % base; t.np = base; }
  t.per_worker = (n + 2 * NS * t.np - 1) / (2 * NS * t.np);
  t.rounds = (t.per_worker + Tmax - 1) / Tmax;
  if (t.rounds < 1) t.rounds = 1;
  t.T = (t.per_worker + t.rounds - 1) / t.rounds;
  t.q_un = (n + t.np * NS * 7 * t.rounds - 1) / (t.np * NS * 7 * t.rounds);
  t.uneven = 0;
  t.k_pure = 0;
  t.per_last = 0;
  if (t.T > 3 * tpw) {  // the even split would put trajectories on the MMA-issuing warp in every round
    if (t.q_un <= tpw) {
      t.uneven = 1;
    } else {
      // not every round can give up an eighth of its rows: as many leading rounds as the spare capacity pays for do
      const int cap_even = t.np * 2 * NS * Tmax, cap_pure = t.np * NS * 7 * tpw;
      int k = (t.rounds * cap_even - n) / (cap_even - cap_pure);
      if (k > t.rounds - 1) k = t.rounds - 1;
      if (k >= 1) {
        const int rest = n - k * cap_pure, workers = (t.rounds - k) * t.np * 2 * NS;
        t.uneven = 2;
        t.k_pure = k;
        t.per_last = (rest + workers - 1) / workers;
      }
    }
  }
  return t;
}
// first trajectory and number of live trajectories of the group a worker (cluster rank, slot s) takes in round rd
__host__ __device__ inline void tc_group(const TcTile& t, int n, int NS, int tpw, int rank, int s, int rd, int even_rounds, int* p0,
                                         int* Gc) {
  int p, Tw;
  const int worker = (t.lp * 2 + rank) * NS + s;
  const int u_off = t.lp * NS * 7 + (rank == 0 ? 3 * s : 3 * NS + 4 * s), units = rank == 0 ? 3 : 4;
  if (t.uneven == 1) {
    Tw = units * t.q_un;
    const int lo = u_off * t.q_un * t.rounds;
    p = lo + rd * Tw;
    const int hi = lo + Tw * t.rounds;
    if (hi < n) n = hi;
  } else if (t.uneven == 2 && rd < t.k_pure) {
    Tw = units * tpw;
    p = rd * (t.np * NS * 7 * tpw) + u_off * tpw;
  } else if (t.uneven == 2) {
    Tw = t.per_last;
    p = t.k_pure * (t.np * NS * 7 * tpw) + ((rd - t.k_pure) * t.np * 2 * NS + worker) * t.per_last;
  } else {
    // rounds are filled warp by warp -- three whole warps per round where that is enough (the issuing warp stays empty), the
    // remainder in the last round -- instead of evenly: fewer busy warps in total, and a light last round
    Tw = (t.T <= 3 * tpw && !even_rounds) ? 3 * tpw : t.T;
    const int lo = worker * t.per_worker;
    p = lo + rd * Tw;
    const int hi = lo + t.per_worker;
    if (hi < n) n = hi;
  }
  *p0 = p < n ? p : n;
  int g = n - p;
  if (g > Tw) g = Tw;
  *Gc = g > 0 ? g : 0;
}
bool make_tc_layout(const ModelDesc& md, int num_sms, TcLayout* out, int* kmap_host /* [kmap_len] or null */, int grow = 0);
// Row layout of the tensor-core GNN rollout for launches over n trajectories: 1 if the "global row" tiling is expected to be
// faster (fewer stages per step at the price of N + 1 instead of N tile rows per trajectory), by a stage-count / MMA-time model.
int tc_pick_grow(const ModelDesc& md, int num_sms, int n);
cudaError_t launch_tc_pack(const float* w32, const ModelDesc& md, const TcLayout& tl, const int* kmap_dev, float* scratch,
                           uint8_t* wimg, float* wpar, cudaStream_t st);
cudaError_t launch_rollout_gnn_tc(const ModelDesc& md, const RolloutArgs& ra, const TcLayout& tl, const uint8_t* wimg,
                                  const float* wpar, int* err_flag, long long* dbg, bool all_fold, const costdev::FusedCost& fc,
                                  cudaStream_t st);

// ---- tensor-core rollout of the dense-MLP ensemble (rollout_mlp_tc.cu) ----
struct TcMlpMat { int K, N, nloc, off; };
struct TcMlpLayout {
  int HD, L, K1, NOUT;      // hidden width, hidden layers, padded input K, padded output width
  TcMlpMat mat[kMaxLayers];
  int image_bytes, nrm_floats;
  int sm_w, sm_nrm, sm_state, sm_snrm, sm_tab, sm_act, sm_bar, smem_bytes;
  int pairs;
};
bool make_mlp_tc_layout(const ModelDesc& md, int num_sms, TcMlpLayout* out);
cudaError_t launch_mlp_tc_pack(const float* w32, const ModelDesc& md, const TcMlpLayout& tl, uint8_t* wimg, cudaStream_t st);
cudaError_t launch_rollout_mlp_tc(const ModelDesc& md, const RolloutArgs& ra, const TcMlpLayout& tl, const uint8_t* wimg, int* err_flag,
                                  const costdev::FusedCost& fc, cudaStream_t st);

cudaError_t launch_epi_bench(int nwg, int with_bias, int with_mma, long long* out_dev);  // profiling aid (ceeus_selftest_epilogue)

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
  unsigned int stream_id;     // Philox stream: stream_id + (*step_ctr) * streams_per_step  (step_ctr may be null)
  const unsigned long long* step_ctr;
  unsigned int streams_per_step;
  long long first_row;  // global index of local row 0 (RNG is keyed by the global row)
  long long global_rows_per_env;
};
cudaError_t launch_sample(const SampleArgs& a, cudaStream_t st);
size_t sample_smem_bytes(int H, int colored);

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
  unsigned int stream_id;     // + (*step_ctr) * streams_per_step, as in SampleArgs
  const unsigned long long* step_ctr;
  unsigned int streams_per_step;
};
cudaError_t launch_fixup(const FixupArgs& a, cudaStream_t st);

struct CostArgs {
  const float* traj;  // [H+1,E,n,O]
  float* cpm;         // [n,E]
  float* costs;       // [n]
  const float* env_cost_ext;  // [n,E,H] task cost evaluated by the caller, or null (built-in env cost)
  int n, H, E, O, N, A, D, S, G, sd;
  int Ostd;  // observation dims that enter the ensemble std: O for caller-supplied trajectories; for the planner's own rollouts
             // only the predicted dims (GNN: A + N*D, dense MLP: sd) -- the rest is copied identically into every member, its
             // std is exactly 0 either way
};
cudaError_t launch_costs(const CostArgs& a, const CostDesc& cd, cudaStream_t st);
// the same reduction on the planner's compact prediction scratch (cost_device.cuh: FusedCost), one warp per trajectory
cudaError_t launch_costs_compact(const costdev::FusedCost& f, const float* start_obs, const float* goal, int traj_per_start,
                                 int traj_per_goal, int n, cudaStream_t st);

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
  // peer-memory exchange (launch_exchange): cand_all is the rank's own double buffer [2][world,nE,k,rec]; the kernel picks the half of
  // exchange number x = *step_ctr * iters_per_step + iter and first waits until every rank's flag has reached x + 1
  const unsigned long long* step_ctr;          // null: cand_all is a plain gathered array (NCCL / caller-gathered)
  const unsigned long long* flags;             // [world] arrival flags written by the peers
  int iters_per_step, iter;
  long long half_stride;                       // floats per half of the double buffer
  int* err;                                    // set to 21 if a peer never arrives
};
cudaError_t launch_update(const UpdateArgs& a, cudaStream_t st);

// Candidate exchange over NVLink peer memory: every rank stores its k candidate records straight into every peer's gather buffer
// (slot `rank` of the half picked by the exchange number) and then raises its flag there -- a few KB per peer, no NCCL launch.
struct ExchangeArgs {
  const float* cand;            // [nE,k,rec] this rank's records
  float* peer_buf[8];           // gather double buffers of all ranks (own included), [2][world,nE,k,rec]
  unsigned long long* peer_flags[8];  // [world] flags of all ranks
  int world, rank, count;       // count = nE*k*rec floats
  long long half_stride;
  const unsigned long long* step_ctr;
  int iters_per_step, iter;
};
cudaError_t launch_exchange(const ExchangeArgs& a, cudaStream_t st);

struct FinishArgs {
  float* mean;
  float* std;
  const float* elite_actions;
  const float* low;
  const float* high;
  float* action_out;  // [nE,U]
  int nE, k, H, U, execute_best, relative_init;
  float init_std;
  unsigned long long* step_ctr;  // MPC step counter on the device (RNG stream base); incremented by finish_kernel
};
cudaError_t launch_finish(const FinishArgs& a, cudaStream_t st);
cudaError_t launch_reset_dist(const FinishArgs& a, cudaStream_t st);  // beginning_of_rollout: mean + std reset

cudaError_t launch_repack(const float* src, float* dst, int E, int rows, int cols, long long member_stride, int dst_off,
                          int dst_ld, cudaStream_t st);

}  // namespace ceeus
