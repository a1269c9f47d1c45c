// CEEUS_PREC_TC: tensor-core (tcgen05) GNN rollout -- the whole horizon of GNNForwardEnsembleModel.predict_n_steps
// (mbrl/models/gnn_ensemble.py:876-983, gnn_modules.py:173-278) in one persistent launch.
//
// Execution model (see DESIGN.md "rollout_gnn_tc_kernel"):
//   * grid = CTA PAIRS (cluster of 2, tcgen05 cta_group::2).  A pair owns one ensemble member and a contiguous chunk of
//     trajectories.  Every linear layer of the member's three MLPs is resident in shared memory for the whole kernel as
//     fp16, split by output column between the two CTAs (each CTA holds N/2 columns of every B operand) -> 114 KB per
//     CTA instead of 224 KB, which is what makes residency possible at all.
//   * per CTA two warpgroups (WG0 = warps 0-3, WG1 = warps 4-7), each with its own A-operand tile in shared memory
//     ("slot") and its own 2 x HD TMEM accumulator columns.  WG s of CTA 0 and WG s of CTA 1 form one M = 256 MMA tile.
//     Warp 8 lane 0 of the leader CTA issues every tcgen05.mma; completion is multicast-committed to both CTAs.
//   * a WG runs groups of T_r trajectories through all H steps.  Per step: edge tile (T_r * N(N-1) rows) through the
//     3-layer edge MLP, on-chip aggregation (mean over the N-1 edges of each source node, mean over all edges), then ONE
//     combined tile whose rows are the T_r * N node rows and the T_r global rows; that tile is multiplied by both the
//     node-MLP and the global-MLP weights (two accumulators) and every row keeps the result of its own MLP.
//   * thread == tile row == TMEM lane in every epilogue: bias + LayerNorm + ReLU run on the 128 fp32 accumulator
//     values of the row held in registers; the fp16 result is written straight back as the next layer's A operand.
//   * operands are fp16 (10-bit mantissa, like TF32) with fp32 accumulation; states, residual update, normalisers and
//     LayerNorm statistics stay fp32.  Parity bar: 1e-3 (BASELINE.json north_star, tf32/bf16 mode).
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace ceeus {

namespace {
using namespace tc;

constexpr int kTcThreads = 320;  // 2 epilogue warpgroups + 2 MMA-issuer warps (warp 8 also does the set-up)
constexpr uint32_t FULL = 0xffffffffu;

__device__ __forceinline__ void wg_sync(int s) { asm volatile("bar.sync %0, 128;" ::"r"(1 + s) : "memory"); }

__device__ __forceinline__ float act_fn(float x, int act) {
  switch (act) {
    case CEEUS_ACT_RELU: return fmaxf(x, 0.f);
    case CEEUS_ACT_SILU: return x / (1.f + __expf(-x));
    case CEEUS_ACT_TANH: return tanhf(x);
    default: return x;
  }
}

struct Ctx {
  uint8_t* smem;
  uint64_t* a_ready;   // [2] (used in the leader CTA)
  uint64_t* d_ready;   // [2]
  volatile int* abort; // per CTA
  int* err;            // global
  uint32_t tmem_base;
  uint32_t a_ready_leader[2];  // cluster addresses of the leader's a_ready barriers
};

__device__ __forceinline__ bool wait_bar(const Ctx& cx, uint64_t* bar, uint32_t parity, int code) {
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    if (mbar_try_wait(bar, parity)) return true;
    if ((it & 1023u) == 1023u && *cx.abort) return false;
  }
  *cx.abort = 1;
  atomicExch(cx.err, code);
  return false;
}

// A-operand tile of this WG is complete: make the generic-proxy writes visible to the tensor core, order prior TMEM
// loads, one arrival per warp on the leader's barrier.
__device__ __forceinline__ void signal_a(const Ctx& cx, int s, int lane) {
  fence_proxy_async();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(cx.a_ready_leader[s]);
}

template <int HD>
__device__ __forceinline__ void load_acc(uint32_t taddr, float (&x)[HD]) {
#pragma unroll
  for (int c0 = 0; c0 < HD; c0 += 32) tmem_ld32(taddr + c0, *reinterpret_cast<uint32_t(*)[32]>(&x[c0]));
}

// bias + LayerNorm + activation on the row held in x[], fp16 result -> A tile chunks [chunk0, chunk0 + HD/8) of `row`
template <int HD>
__device__ __forceinline__ void finish_row(float (&x)[HD], const float* __restrict__ pb, const float* __restrict__ pg,
                                           const float* __restrict__ pbe, bool ln, int act, uint8_t* tile, int chunk0,
                                           int row, bool store) {
#pragma unroll
  for (int j = 0; j < HD; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(pb + j);
    x[j] += b.x; x[j + 1] += b.y; x[j + 2] += b.z; x[j + 3] += b.w;
  }
  if (ln) {
    float s4[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < HD; j += 8) {
#pragma unroll
      for (int t = 0; t < 8; ++t) s4[t] += x[j + t];
    }
    const float mean = (((s4[0] + s4[1]) + (s4[2] + s4[3])) + ((s4[4] + s4[5]) + (s4[6] + s4[7]))) * (1.f / HD);
    float v4[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < HD; j += 8) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        x[j + t] -= mean;
        v4[t] = fmaf(x[j + t], x[j + t], v4[t]);
      }
    }
    const float v = ((v4[0] + v4[1]) + (v4[2] + v4[3])) + ((v4[4] + v4[5]) + (v4[6] + v4[7]));
    const float rstd = rsqrtf(v * (1.f / HD) + 1e-5f);
#pragma unroll
    for (int j = 0; j < HD; j += 4) {
      const float4 g = *reinterpret_cast<const float4*>(pg + j);
      const float4 be = *reinterpret_cast<const float4*>(pbe + j);
      x[j] = fmaf(x[j] * rstd, g.x, be.x);
      x[j + 1] = fmaf(x[j + 1] * rstd, g.y, be.y);
      x[j + 2] = fmaf(x[j + 2] * rstd, g.z, be.z);
      x[j + 3] = fmaf(x[j + 3] * rstd, g.w, be.w);
    }
  }
  if (store) {
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      uint4 o;
      o.x = pack_half2(act_fn(x[8 * c + 0], act), act_fn(x[8 * c + 1], act));
      o.y = pack_half2(act_fn(x[8 * c + 2], act), act_fn(x[8 * c + 3], act));
      o.z = pack_half2(act_fn(x[8 * c + 4], act), act_fn(x[8 * c + 5], act));
      o.w = pack_half2(act_fn(x[8 * c + 6], act), act_fn(x[8 * c + 7], act));
      *reinterpret_cast<uint4*>(tile + (size_t)(chunk0 + c) * 2048 + row * 16) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
    rollout_gnn_tc_kernel(const ModelDesc md, const RolloutArgs ra, const TcLayout tl, const uint8_t* __restrict__ wimg,
                          const float* __restrict__ wpar, int* __restrict__ err, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int e = pair / tl.ppm;
  const int chunk_id = pair % tl.ppm;
  const int N = md.N, A = md.A, D = md.D, S = md.S, U = md.U, sd = md.sd, O = md.O;
  const int NA = D + S, NE = N * (N - 1), N1 = N + 1;
  const int T_r = tl.T_r;
  const int KXc = tl.KX / 8;  // chunks of the [x|g|u] part of the combined tile

  uint8_t* sW = smem + tl.sm_w;
  float* sPar = reinterpret_cast<float*>(smem + tl.sm_par);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + tl.sm_bar);
  uint64_t* wbar = bars;        // weights landed
  uint64_t* a_ready = bars + 1; // [2]
  uint64_t* d_ready = bars + 3; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  // ---- one-time setup ----
  if (tid == 0) *abort_flag = 0;
  if (warp == 8) {
    if (lane == 0) {
      mbar_init(wbar, 1);
      mbar_init(&a_ready[0], 8);
      mbar_init(&a_ready[1], 8);
      mbar_init(&d_ready[0], 1);
      mbar_init(&d_ready[1], 1);
      fence_barrier_init();
      mbar_expect_tx(wbar, (uint32_t)tl.image_bytes);
      bulk_g2s(sW, wimg + ((size_t)e * 2 + rank) * tl.image_bytes, (uint32_t)tl.image_bytes, wbar);
    }
    __syncwarp();
    tmem_alloc<2>(tmem_slot, 512);
    tmem_relinquish<2>();
  }
  for (int i = tid; i < tl.par_floats; i += kTcThreads) sPar[i] = __ldg(wpar + (size_t)e * tl.par_floats + i);
  tc_fence_before();
  __syncthreads();
  Ctx cx;
  cx.smem = smem; cx.a_ready = a_ready; cx.d_ready = d_ready; cx.abort = abort_flag; cx.err = err;
  if (warp == 8 && lane == 0) wait_bar(cx, wbar, 0, 1);  // this CTA's weight image has landed
  __syncthreads();
  cluster_sync();  // both CTAs: barriers initialised, TMEM allocated, weights resident
  tc_fence_after();
  cx.tmem_base = *tmem_slot;
  cx.a_ready_leader[0] = mapa(smem_u32(&a_ready[0]), 0);
  cx.a_ready_leader[1] = mapa(smem_u32(&a_ready[1]), 0);

  // trajectories of this pair
  const int chunk = (ra.n + tl.ppm - 1) / tl.ppm;
  const int c_lo = min(ra.n, chunk_id * chunk), c_hi = min(ra.n, c_lo + chunk);
  const int n_groups = (c_hi - c_lo + T_r - 1) / T_r;
  const int n_iter = (n_groups + 3) / 4;  // 4 workers per pair, lock-step
  const int total_stages = n_iter * ra.H * 6;

  if (warp >= 8) {
    // =========================== MMA issuers (leader CTA): warp 8 -> slot 0, warp 9 -> slot 1 ===========================
    // One thread per slot, each blocking on its own barrier: mbarrier.try_wait suspends the thread for a
    // system-defined time, so a single thread polling both slots alternately stalled every stage of one slot
    // behind the other's time-out (first version: ~10 us per stage).
    if (rank == 0 && lane == 0) {
      const int s = warp - 8;
      const uint32_t idesc_h = make_idesc_f16(256, HD);
      const uint32_t idesc_o = make_idesc_f16(256, 16);
      const uint32_t wbase = smem_u32(sW);
      const uint32_t a_addr = smem_u32(smem + tl.sm_slot[s]);
      const uint32_t d0 = cx.tmem_base + (uint32_t)(s * 2 * HD);
      const uint32_t d1 = d0 + HD;
      auto issue = [&](uint32_t d, const TcMat& m, uint32_t idesc) {
        const uint32_t lbo_b = (uint32_t)m.nloc * 16;
#pragma unroll 1
        for (int ks = 0; ks < m.K / 16; ++ks) {
          const uint64_t da = make_smem_desc(a_addr + (uint32_t)(2 * ks) * 2048, 2048, 128);
          const uint64_t db = make_smem_desc(wbase + (uint32_t)m.off + (uint32_t)(2 * ks) * lbo_b, lbo_b, 128);
          umma_f16_ss<2>(d, da, db, idesc, ks > 0 ? 1u : 0u);
        }
      };
      uint32_t ph = 0;
#pragma unroll 1
      for (int stage = 0; stage < total_stages; ++stage) {
        if (!wait_bar(cx, &a_ready[s], ph, 2)) break;
        ph ^= 1u;
        tc_fence_after();
        switch (stage % 6) {
          case 0: issue(d0, tl.mat[0], idesc_h); break;
          case 1: issue(d0, tl.mat[1], idesc_h); break;
          case 2: issue(d0, tl.mat[2], idesc_h); break;
          case 3: issue(d0, tl.mat[3], idesc_h); issue(d1, tl.mat[6], idesc_h); break;
          case 4: issue(d0, tl.mat[4], idesc_h); issue(d1, tl.mat[7], idesc_h); break;
          default: issue(d0, tl.mat[5], idesc_o); issue(d1, tl.mat[8], idesc_o); break;
        }
        umma_commit_cg2_mc(&d_ready[s], 0x3);
      }
    }
  } else if (warp < 8) {
    // =========================== warpgroup: tile builder + epilogue ===========================
    const int s = warp >> 2;       // slot
    const int q = warp & 3;        // TMEM lane quadrant
    const int row = tid & 127;     // tile row == TMEM lane
    const int worker = (int)rank * 2 + s;
    uint8_t* tile = smem + tl.sm_slot[s];
    float* sraw = reinterpret_cast<float*>(smem + tl.sm_state[s]);
    __half* snrm = reinterpret_cast<__half*>(smem + tl.sm_state[s] + tl.sraw_bytes);
    const int SN = A + U + N * NA;  // normalised [agent | action | nodes]
    const float* __restrict__ nrm = ra.norm;
    const bool ln = md.layer_norm != 0;
    const int act = md.act;
    const float inv_node = md.aggr_mean ? 1.f / fmaxf((float)(N - 1), 1.f) : 1.f;
    const float inv_glob = md.aggr_mean ? 1.f / fmaxf((float)NE, 1.f) : 1.f;
    const uint32_t tm_row = cx.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 2 * HD);
    uint32_t dph = 0;
    auto par = [&](int m, int which) { return sPar + tl.par_off[m][which]; };
    // combined row r = b*(N+1)+i  ->  (first message row, number of messages): node i<N: its N-1 outgoing edges,
    // i==N (global row): all N(N-1) edges of trajectory b
    int dbg_n = 0;
    const bool dbg_on = dbg != nullptr && blockIdx.x == 0 && tid == 0;
    auto stamp = [&](int tag) {
      if (dbg_on && dbg_n < 250) { dbg[2 * dbg_n] = tag; dbg[2 * dbg_n + 1] = clock64(); ++dbg_n; }
    };
    int2* rinfo = reinterpret_cast<int2*>(smem + tl.sm_rinfo) + s * 128;
    {
      const int b = row / N1, i = row % N1;
      rinfo[row] = make_int2(b * NE + (i < N ? i * (N - 1) : 0), i < N ? N - 1 : NE);
    }

    for (int it = 0; it < n_iter; ++it) {
      const int g = it * 4 + worker;
      const int p0 = c_lo + g * T_r;
      const int Gc = g < n_groups ? min(T_r, c_hi - p0) : 0;  // live trajectories of this group (0: pure lock-step filler)
      const int rows_e = Gc * NE, rows_c = Gc * N1;
      wg_sync(s);
      // ---- start observation -> raw state, traj[0] ----
      for (int idx = row; idx < Gc * sd; idx += 128) {
        const int b = idx / sd, c = idx % sd;
        sraw[idx] = __ldg(ra.start_obs + (size_t)((p0 + b) / ra.traj_per_start) * O + c);
      }
      for (int idx = row; idx < Gc * O; idx += 128) {
        const int b = idx / O, c = idx % O;
        ra.traj[((size_t)e * ra.n + p0 + b) * O + c] = __ldg(ra.start_obs + (size_t)((p0 + b) / ra.traj_per_start) * O + c);
      }
      wg_sync(s);

      for (int h = 0; h < ra.H; ++h) {
        stamp(100);
        // ---- normalise state + action (gnn_ensemble.py:190-243) -> fp16 ----
        for (int idx = row; idx < Gc * SN; idx += 128) {
          const int b = idx / SN, c = idx % SN;
          float v;
          if (c < A) {
            v = (sraw[b * sd + c] - __ldg(nrm + md.nrm.in_agent + c)) / __ldg(nrm + md.nrm.in_agent + A + c);
          } else if (c < A + U) {
            const int u = c - A;
            const float a = __ldg(ra.actions + ((size_t)(p0 + b) * ra.H + h) * U + u);
            v = (a - __ldg(nrm + md.nrm.in_action + u)) / __ldg(nrm + md.nrm.in_action + U + u);
          } else {
            const int qq = c - A - U, i = qq / NA, f = qq % NA;
            if (f < D)
              v = (sraw[b * sd + A + i * D + f] - __ldg(nrm + md.nrm.in_dyn + f)) / __ldg(nrm + md.nrm.in_dyn + D + f);
            else
              v = (sraw[b * sd + A + N * D + i * S + (f - D)] - __ldg(nrm + md.nrm.in_stat + (f - D))) /
                  __ldg(nrm + md.nrm.in_stat + S + (f - D));
          }
          snrm[idx] = __float2half_rn(v);
        }
        wg_sync(s);
        stamp(101);
        // ---- edge tile, layer-1 input [x_i | x_j | agent | action | 0] (gnn_modules.py:194-211) ----
        if (row < rows_e) {
          const int b = row / NE, qq = row % NE;
          const int i = qq / (N - 1), jj = qq % (N - 1);
          const int j = jj + (jj >= i ? 1 : 0);
          const __half* sb = snrm + b * SN;
          for (int c = 0; c < tl.KE1 / 8; ++c) {
            __align__(16) __half tmp[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const int k = c * 8 + t;
              __half v = __float2half_rn(0.f);
              if (k < NA) v = sb[A + U + i * NA + k];
              else if (k < 2 * NA) v = sb[A + U + j * NA + (k - NA)];
              else if (k < 2 * NA + A) v = sb[k - 2 * NA];
              else if (k < 2 * NA + A + U) v = sb[A + (k - 2 * NA - A)];
              tmp[t] = v;
            }
            *reinterpret_cast<uint4*>(tile + (size_t)c * 2048 + row * 16) = *reinterpret_cast<const uint4*>(tmp);
          }
        }
        stamp(102);
        signal_a(cx, s, lane);
        stamp(103);

        // ---- six MMA stages of this step: E1 E2 E3 (edge MLP) C1 C2 C3 (node + global MLP on the combined tile) ----
        // One copy of the row epilogue serves all hidden stages: code size matters (the first version unrolled it five
        // times into 350 KB of SASS and stalled on instruction fetch).
        const int cb = row / N1, ci = row % N1;
        const bool is_glob = ci == N;
        const bool live_c = row < rows_c, live_e = row < rows_e;
        const bool has_n = __any_sync(FULL, live_c && !is_glob);
        const bool has_g = __any_sync(FULL, live_c && is_glob);
        const bool warp_live_e = q * 32 < rows_e;
#pragma unroll 1
        for (int st = 0; st < 6; ++st) {
          stamp(10 + st);
          wait_bar(cx, &d_ready[s], dph, 10 + st);
          dph ^= 1u;
          tc_fence_after();
          stamp(20 + st);
          const bool cstage = st >= 3;
          const bool warp_on = cstage ? (has_n || has_g) : warp_live_e;
          if (st < 5) {
            if (warp_on) {
              float x[HD];
              const bool from_g = cstage && !has_n;  // warp holds only global rows
              load_acc<HD>(tm_row + (from_g ? HD : 0), x);
              tmem_wait_ld();
              if (cstage && has_n && has_g) {  // mixed warp: global rows take the second accumulator
#pragma unroll
                for (int c0 = 0; c0 < HD; c0 += 32) {
                  uint32_t t32[32];
                  tmem_ld32(tm_row + HD + c0, t32);
                  tmem_wait_ld();
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (is_glob) x[c0 + j] = __uint_as_float(t32[j]);
                }
              }
              const int m = cstage ? (is_glob ? 6 : 3) + (st - 3) : st;
              const bool hidden = st != 2;
              finish_row<HD>(x, par(m, 0), par(m, hidden ? 1 : 0), par(m, hidden ? 2 : 0), ln && hidden,
                             hidden ? act : CEEUS_ACT_NONE, tile, st == 2 ? KXc : 0, row, cstage ? live_c : live_e);
            }
            stamp(30 + st);
            if (st == 2) {
              // ---- aggregation (models/utils.py:42-51): read every message chunk this thread needs, then write ----
              tc_fence_before();
              wg_sync(s);
              constexpr int CH = HD / 8;
              uint4 outv[CH];  // rows_c <= 128  =>  at most CH items per thread
              const int n_items = rows_c * CH;
              {
                int c = 0, r = row;  // item = c * rows_c + r, advanced by 128 per slot without divisions
                if (rows_c > 0) { c = row / rows_c; r = row % rows_c; }
                const int dq = rows_c > 0 ? 128 / rows_c : 0, dr = rows_c > 0 ? 128 % rows_c : 0;
#pragma unroll
                for (int t = 0; t < CH; ++t) {
                  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                  if (row + t * 128 < n_items) {
                    const int2 info = rinfo[r];  // (first message row, message count) of combined row r
                    const uint8_t* src = tile + (size_t)(KXc + c) * 2048 + (size_t)info.x * 16;
                    for (int e2 = 0; e2 < info.y; ++e2) {
                      const uint4 mm = *reinterpret_cast<const uint4*>(src + e2 * 16);
                      const __half2* hp = reinterpret_cast<const __half2*>(&mm);
#pragma unroll
                      for (int t2 = 0; t2 < 4; ++t2) {
                        const float2 f = __half22float2(hp[t2]);
                        acc[2 * t2] += f.x;
                        acc[2 * t2 + 1] += f.y;
                      }
                    }
                    const float sc = info.y == NE ? inv_glob : inv_node;  // N-1 != N(N-1) for every N >= 2
#pragma unroll
                    for (int t2 = 0; t2 < 8; ++t2) acc[t2] *= sc;
                  }
                  outv[t].x = pack_half2(acc[0], acc[1]);
                  outv[t].y = pack_half2(acc[2], acc[3]);
                  outv[t].z = pack_half2(acc[4], acc[5]);
                  outv[t].w = pack_half2(acc[6], acc[7]);
                  c += dq; r += dr;
                  if (r >= rows_c && rows_c > 0) { r -= rows_c; c += 1; }
                }
              }
              wg_sync(s);
              {
                int c = 0, r = row;
                if (rows_c > 0) { c = row / rows_c; r = row % rows_c; }
                const int dq = rows_c > 0 ? 128 / rows_c : 0, dr = rows_c > 0 ? 128 % rows_c : 0;
#pragma unroll
                for (int t = 0; t < CH; ++t) {
                  if (row + t * 128 < n_items) *reinterpret_cast<uint4*>(tile + (size_t)(KXc + c) * 2048 + r * 16) = outv[t];
                  c += dq; r += dr;
                  if (r >= rows_c && rows_c > 0) { r -= rows_c; c += 1; }
                }
              }
              // ---- combined tile, [x_i | agent | action | 0] part (global rows: x_i = 0)  gnn_modules.py:213-223,250-254 ----
              if (live_c) {
                const __half* sb = snrm + cb * SN;
                for (int c = 0; c < KXc; ++c) {
                  __align__(16) __half tmp[8];
#pragma unroll
                  for (int t = 0; t < 8; ++t) {
                    const int k = c * 8 + t;
                    __half v = __float2half_rn(0.f);
                    if (k < NA) { if (!is_glob) v = sb[A + U + ci * NA + k]; }
                    else if (k < NA + A) v = sb[k - NA];
                    else if (k < NA + A + U) v = sb[A + (k - NA - A)];
                    tmp[t] = v;
                  }
                  *reinterpret_cast<uint4*>(tile + (size_t)c * 2048 + row * 16) = *reinterpret_cast<const uint4*>(tmp);
                }
              }
            }
            stamp(40 + st);
            signal_a(cx, s, lane);
            stamp(50 + st);
          } else if (warp_on) {
            // ---- output layers: deltas -> de-normalise, residual update (gnn_ensemble.py:302-334, 935-939) ----
            uint32_t v[16];
            tmem_ld16(tm_row + (has_n ? 0 : HD), v);
            tmem_wait_ld();
            if (has_n && has_g) {
              uint32_t w[16];
              tmem_ld16(tm_row + HD, w);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (is_glob) v[j] = w[j];
            }
            if (live_c) {
              const float* pb = par(is_glob ? 8 : 5, 0);
              const int nout = is_glob ? A : D;
              const int so = is_glob ? md.nrm.out_agent : md.nrm.out_dyn;
              float* dst = sraw + cb * sd + (is_glob ? 0 : A + ci * D);
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < nout) {
                  const float dl = __uint_as_float(v[j]) + pb[j];
                  dst[j] += fmaf(__ldg(nrm + so + nout + j), dl, __ldg(nrm + so + j));
                }
            }
          }
        }
        tc_fence_before();
        wg_sync(s);
        stamp(60);
        // ---- emit traj[h+1] = [state | goal] (env.obs_postproc) ----
        {
          float* outp = ra.traj + (((size_t)(h + 1) * md.E + e) * ra.n + p0) * O;
          for (int idx = row; idx < Gc * O; idx += 128) {
            const int b = idx / O, c = idx % O;
            outp[idx] = c < sd ? sraw[b * sd + c] : __ldg(ra.goal + (size_t)((p0 + b) / ra.traj_per_goal) * md.G + (c - sd));
          }
        }
      }
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 8) tmem_dealloc<2>(cx.tmem_base, 512);
}

// fp32 packed weights (SIMT layout) -> fp16 B-operand images + fp32 parameter blocks
__global__ void tc_pack_kernel(const float* __restrict__ w32, const ModelDesc md, const TcLayout tl, uint8_t* __restrict__ wimg,
                               float* __restrict__ wpar) {
  const int e = blockIdx.y;
  const float* wm = w32 + (size_t)e * md.member_stride;
  const int halves = tl.image_bytes / 2;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * halves; idx += gridDim.x * blockDim.x) {
    const int rank = idx / halves;
    const int off_b = (idx % halves) * 2;
    int m = 0;
    for (int t = 1; t < 9; ++t)
      if (off_b >= tl.mat[t].off) m = t;
    const TcMat& M = tl.mat[m];
    const int el = (off_b - M.off) / 2;
    const int c = el / (M.nloc * 8), rem = el % (M.nloc * 8);
    const int nl = rem / 8, j = rem % 8;
    const int k = c * 8 + j;
    const int n = rank * M.nloc + nl;
    float v = 0.f;
    const LayerDesc& ld = md.mlp[M.src_mlp].layer[M.src_layer];
    int ks = -1;
    if (k >= M.seg_tc[0] && k < M.seg_tc[0] + M.seg_len[0]) ks = M.seg_src[0] + (k - M.seg_tc[0]);
    else if (k >= M.seg_tc[1] && k < M.seg_tc[1] + M.seg_len[1]) ks = M.seg_src[1] + (k - M.seg_tc[1]);
    if (ks >= 0 && ks < ld.K && n < ld.N) v = wm[ld.w_off + (size_t)ks * ld.Npad + n];
    reinterpret_cast<__half*>(wimg + ((size_t)e * 2 + rank) * tl.image_bytes)[off_b / 2] = __float2half_rn(v);
  }
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < tl.par_floats; idx += gridDim.x * blockDim.x) {
    float v = 0.f;
    for (int m = 0; m < 9; ++m) {
      const LayerDesc& ld = md.mlp[tl.mat[m].src_mlp].layer[tl.mat[m].src_layer];
      for (int w = 0; w < 3; ++w) {
        const int o = tl.par_off[m][w];
        if (o >= 0 && idx >= o && idx < o + tl.par_len[m]) {
          const int j = idx - o;
          if (j < ld.N) v = wm[(w == 0 ? ld.b_off : w == 1 ? ld.g_off : ld.be_off) + j];
        }
      }
    }
    wpar[(size_t)e * tl.par_floats + idx] = v;
  }
}

}  // namespace

bool make_tc_layout(const ModelDesc& md, int num_sms, TcLayout* out) {
  TcLayout t{};
  if (md.kind != CEEUS_MODEL_GNN || (md.Hd != 64 && md.Hd != 128) || md.L != 2) return false;
  if (md.D > 16 || md.A > 16) return false;
  const int NA = md.D + md.S, NE = md.N * (md.N - 1);
  t.HD = md.Hd;
  t.KE1 = round_up(2 * NA + md.A + md.U, 16);
  t.KX = round_up(NA + md.A + md.U, 16);
  if (t.KE1 > md.Hd) return false;  // the edge tile's layer-1 input shares the K = HD chunk range
  const int Hd = md.Hd;
  t.T_r = 128 / NE < 128 / (md.N + 1) ? 128 / NE : 128 / (md.N + 1);
  if (t.T_r < 1) return false;
  auto mat = [&](int m, int K, int N, int mlp, int layer) {
    t.mat[m].K = K; t.mat[m].N = N; t.mat[m].nloc = N / 2; t.mat[m].src_mlp = mlp; t.mat[m].src_layer = layer;
    t.mat[m].seg_tc[0] = 0; t.mat[m].seg_len[0] = K; t.mat[m].seg_src[0] = 0;
    t.mat[m].seg_tc[1] = 0; t.mat[m].seg_len[1] = 0; t.mat[m].seg_src[1] = 0;
  };
  mat(0, t.KE1, Hd, 0, 0); t.mat[0].seg_len[0] = 2 * NA + md.A + md.U;
  mat(1, Hd, Hd, 0, 1);
  mat(2, Hd, Hd, 0, 2);
  mat(3, t.KX + Hd, Hd, 1, 0);
  t.mat[3].seg_len[0] = NA + md.A + md.U;
  t.mat[3].seg_tc[1] = t.KX; t.mat[3].seg_len[1] = Hd; t.mat[3].seg_src[1] = NA + md.A + md.U;
  mat(4, Hd, Hd, 1, 1);
  mat(5, Hd, 16, 1, 2);
  mat(6, t.KX + Hd, Hd, 2, 0);
  t.mat[6].seg_tc[0] = NA; t.mat[6].seg_len[0] = md.A + md.U; t.mat[6].seg_src[0] = 0;
  t.mat[6].seg_tc[1] = t.KX; t.mat[6].seg_len[1] = Hd; t.mat[6].seg_src[1] = md.A + md.U;
  mat(7, Hd, Hd, 2, 1);
  mat(8, Hd, 16, 2, 2);
  int off = 0;
  for (int m = 0; m < 9; ++m) {
    t.mat[m].off = off;
    off += (t.mat[m].K / 8) * t.mat[m].nloc * 16;
  }
  t.image_bytes = off;
  int po = 0;
  for (int m = 0; m < 9; ++m) {
    const bool hidden = (m % 3) != 2;
    t.par_len[m] = (m == 5 || m == 8) ? 16 : Hd;
    t.par_off[m][0] = po; po += t.par_len[m];
    if (hidden) {
      t.par_off[m][1] = po; po += Hd;
      t.par_off[m][2] = po; po += Hd;
    } else {
      t.par_off[m][1] = t.par_off[m][2] = -1;
    }
  }
  t.par_floats = po;
  // shared memory carve-up
  int sm = 0;
  t.sm_w = sm; sm += round_up(t.image_bytes, 1024);
  const int slot_bytes = (t.KX + Hd) / 8 * 2048;
  t.sm_slot[0] = sm; sm += slot_bytes;
  t.sm_slot[1] = sm; sm += slot_bytes;
  t.sm_par = sm; sm += round_up(t.par_floats * 4, 16);
  t.sraw_bytes = round_up(t.T_r * md.sd * 4, 16);
  const int snrm_bytes = round_up(t.T_r * (md.A + md.U + md.N * NA) * 2, 16);
  t.sm_state[0] = sm; sm += t.sraw_bytes + snrm_bytes;
  t.sm_state[1] = sm; sm += t.sraw_bytes + snrm_bytes;
  t.sm_rinfo = sm; sm += 2 * 128 * 8;
  t.sm_bar = sm; sm += 64;
  t.smem_bytes = sm;
  if (t.smem_bytes > 227 * 1024) return false;
  const int pairs = num_sms / 2;
  t.ppm = pairs / md.E;
  if (t.ppm < 1) return false;
  *out = t;
  return true;
}

cudaError_t launch_tc_pack(const float* w32, const ModelDesc& md, const TcLayout& tl, uint8_t* wimg, float* wpar, cudaStream_t st) {
  dim3 grid(64, md.E);
  tc_pack_kernel<<<grid, 256, 0, st>>>(w32, md, tl, wimg, wpar);
  return cudaGetLastError();
}

cudaError_t launch_rollout_gnn_tc(const ModelDesc& md, const RolloutArgs& ra, const TcLayout& tl, const uint8_t* wimg,
                                  const float* wpar, int* err_flag, long long* dbg, cudaStream_t st) {
  cudaError_t e;
  const dim3 grid(2 * tl.ppm * md.E);
  if (md.Hd == 128) {
    e = cudaFuncSetAttribute(rollout_gnn_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, tl.smem_bytes);
    if (e != cudaSuccess) return e;
    rollout_gnn_tc_kernel<128><<<grid, kTcThreads, tl.smem_bytes, st>>>(md, ra, tl, wimg, wpar, err_flag, dbg);
  } else {
    e = cudaFuncSetAttribute(rollout_gnn_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tl.smem_bytes);
    if (e != cudaSuccess) return e;
    rollout_gnn_tc_kernel<64><<<grid, kTcThreads, tl.smem_bytes, st>>>(md, ra, tl, wimg, wpar, err_flag, dbg);
  }
  return cudaGetLastError();
}

}  // namespace ceeus
