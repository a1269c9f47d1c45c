// CEEUS_PREC_TC for the dense-MLP ensemble: the whole horizon of ParallelNNDeterministicEnsemble.predict_n_steps
// (mbrl/models/mlp_ensemble.py:502-611; model = MLPParallelEnsemble, simple.py:18-101: [Linear -> act] x L -> Linear,
// yaml construction/common/mlp_ensemble.yaml: 62 -> 256 -> 256 -> 256 -> 58, SiLU, no LayerNorm) in one persistent launch.
//
// Same machinery as rollout_tc.cu: CTA pairs (tcgen05 cta_group::2, M = 256), one ensemble member per pair, all L+1 weight
// matrices resident in shared memory as fp16 (split by output column between the two CTAs: 169 KB per CTA for the yaml model),
// A operands in TMEM (tcgen05.mma TS form), biases in a constant-one K column.  One tile of 128 trajectories per CTA; two
// threads per row, each owning half of the 256 accumulator columns (no LayerNorm, so the epilogue is purely element-wise).
// TMEM: accumulator [0, HD), A operand [HD, HD + HD/2), constant-one block [HD + HD/2, +8).
// SiLU is evaluated as h + h * tanh(h), h = x / 2, with tanh.approx.f32 (one MUFU per element, 2^-11 relative).
#include <cuda_fp16.h>

#include "common.cuh"
#include "cost_device.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"
#include "tc_rollout_common.cuh"

namespace ceeus {

namespace {
using namespace tc;

constexpr int kMlpThreads = 256;
constexpr int kMaxOW = 64;  // output columns per thread (NOUT / 2; state dim <= 128)

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool SILU>
__device__ __forceinline__ float mlp_act(float x, int act) {
  if constexpr (SILU) {
    const float h = 0.5f * x;
    return fmaf(h, tanh_approx(h), h);
  } else {
    return act_fn(x, act);
  }
}

template <int HD, bool SILU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
    rollout_mlp_tc_kernel(const ModelDesc md, const RolloutArgs ra, const TcMlpLayout tl, const uint8_t* __restrict__ wimg,
                          int* __restrict__ err, const costdev::FusedCost fc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const bool fused = fc.pred != nullptr;  // planner launch: predictions -> compact scratch, costs reduced in the tail
  constexpr int W = HD / 2;  // accumulator columns per thread in a hidden layer
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int U = md.U, sd = md.sd, O = md.O, in = sd + U, L = tl.L, K1 = tl.K1;
  const int SNS = K1 + 8;  // snrm row stride in halves: 16 bytes of padding keep the per-row 16-byte accesses conflict free

  const TcTile tt = tc_tile(tl.pairs, md.E, pair, ra.n, 1, 128, 32);  // one 128-trajectory tile per CTA
  const int e = tt.e, lp = tt.lp, per_worker = tt.per_worker, rounds = tt.rounds;
  const int T = tt.T;  // trajectories per tile, <= 128
  const int worker = lp * 2 + (int)rank;

  uint8_t* sW = smem + tl.sm_w;
  float* sNrm = reinterpret_cast<float*>(smem + tl.sm_nrm);
  float* s_state = reinterpret_cast<float*>(smem + tl.sm_state);  // [128][sd] raw fp32 state
  __half* snrm = reinterpret_cast<__half*>(smem + tl.sm_snrm);    // [128][SNS] normalised [state | action | 1 | 0..]
  float4* s_tab = reinterpret_cast<float4*>(smem + tl.sm_tab);    // [NOUT] output-layer table (std_out, mean_out, mean_in, 1/std_in)
  float* s_act = reinterpret_cast<float*>(smem + tl.sm_act);      // [128][8] next step's raw action
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + tl.sm_bar);
  uint64_t* wbar = bars;
  uint64_t* a_ready = bars + 1;
  uint64_t* d_ready = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  if (tid == 0) *abort_flag = 0;
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(wbar, 1);
      mbar_init(a_ready, 16);  // 8 warps x 2 CTAs
      mbar_init(d_ready, 1);
      fence_barrier_init();
      mbar_expect_tx(wbar, (uint32_t)tl.image_bytes);
      bulk_g2s(sW, wimg + ((size_t)e * 2 + rank) * tl.image_bytes, (uint32_t)tl.image_bytes, wbar);
    }
    __syncwarp();
    tmem_alloc<2>(tmem_slot, 512);
    tmem_relinquish<2>();
  }
  // normalisers: input group as (mean, 1/std), output group as (mean, std)
  for (int i = tid; i < tl.nrm_floats; i += kMlpThreads) {
    const float v = __ldg(ra.norm + i);
    sNrm[i] = (i >= md.nrm.in_all + in && i < md.nrm.out_all) ? 1.f / v : v;
  }
  for (int d = tid; d < tl.NOUT; d += kMlpThreads)
    s_tab[d] = d < sd ? make_float4(__ldg(ra.norm + md.nrm.out_all + sd + d), __ldg(ra.norm + md.nrm.out_all + d),
                                    __ldg(ra.norm + md.nrm.in_all + d), 1.f / __ldg(ra.norm + md.nrm.in_all + in + d))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
  tc_fence_before();
  __syncthreads();
  Ctx cx;
  cx.abort = abort_flag; cx.err = err;
  if (warp == 0 && lane == 0) wait_bar(cx, wbar, 0, 1);
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  if (*tmem_slot != 0u && tid == 0) atomicExch(err, 7);
  cx.a_ready_leader[0] = mapa(smem_u32(a_ready), 0);
  cx.a_ready_leader[1] = cx.a_ready_leader[0];

  const int hf = warp >> 2;   // column half
  const int q = warp & 3;     // TMEM lane quadrant
  const int row = tid & 127;  // tile row == TMEM lane == trajectory of the tile
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const uint32_t t_acc = lane_base, t_a = lane_base + HD, t_one = lane_base + HD + HD / 2;
  const float* __restrict__ nrm = sNrm;
  const int act = md.act;
  const bool issuer_warp = rank == 0 && warp == 3;  // the warp of the last rows: idle (no epilogue of its own) whenever the tile is not full
  uint32_t aph = 0, dph = 0;
  const uint32_t wbase = smem_u32(sW);
  auto mma_layer = [&](int m) {
    if (!issuer_warp) return;
    const bool ok = wait_bar_spin(cx, a_ready, aph, 2);
    aph ^= 1u;
    tc_fence_after();
    if (ok) {
      const uint32_t elected = elect_one();
      const TcMlpMat& mm = tl.mat[m];
      const uint32_t idesc = make_idesc_f16(256, mm.N);
      const uint32_t lbo_b = (uint32_t)mm.nloc * 16;
      const uint64_t db0 = make_smem_desc(wbase + (uint32_t)mm.off, lbo_b, 128);
      const uint32_t db_step = (2u * lbo_b) >> 4;
      const int kall = mm.K / 16;
#pragma unroll 2
      for (int ks = 0; ks < kall; ++ks)
        umma_f16_ts_cg2_elect(0u, (uint32_t)HD + (uint32_t)(8 * ks), db0 + (uint64_t)((uint32_t)ks * db_step), idesc, ks > 0 ? 1u : 0u, elected);
      umma_commit_cg2_mc_elect(d_ready, 0x3, elected);
    }
    __syncwarp();
  };
  auto cta_sync = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
  // the constant-one K block behind the hidden activations (bias rows of layers 1..L): written once
  tmem_st4(t_one, make_uint4(0x00003c00u, 0u, 0u, 0u));  // both threads of a row write the same value
  tmem_st4(t_one + 4, make_uint4(0u, 0u, 0u, 0u));
  tmem_wait_st();

  uint32_t satm = 0u;  // record of saturated fp16 operand conversions of this thread (tc_common.cuh: sat_track)
  for (int rd = 0; rd < rounds; ++rd) {
    const int w_lo = min(ra.n, worker * per_worker), w_hi = min(ra.n, w_lo + per_worker);
    const int p0 = w_lo + rd * T;
    const int Gc = max(0, min(T, w_hi - p0));
    const bool live = row < Gc;
    cta_sync();
    // ---- tile init: start observation -> state, snrm, traj[0] ----
    for (int idx = tid; idx < (128 * SNS) / 8; idx += kMlpThreads) reinterpret_cast<uint4*>(snrm)[idx] = make_uint4(0u, 0u, 0u, 0u);
    for (int idx = tid; idx < Gc * O; idx += kMlpThreads) {
      const int b = idx / O, c = idx % O;
      const float v = __ldg(ra.start_obs + (size_t)((p0 + b) / ra.traj_per_start) * O + c);
      if (!fused) ra.traj[((size_t)e * ra.n + p0 + b) * O + c] = v;
      if (c < sd) s_state[b * sd + c] = v;
    }
    cta_sync();
    for (int idx = tid; idx < Gc * in; idx += kMlpThreads) {
      const int b = idx / in, c = idx % in;
      const float x = c < sd ? s_state[b * sd + c] : __ldg(ra.actions + ((size_t)(p0 + b) * ra.H) * U + (c - sd));
      satm = sat_track1(satm, snrm[b * SNS + c] = half_sat((x - nrm[md.nrm.in_all + c]) * nrm[md.nrm.in_all + in + c]));
    }
    if (tid < 128 && live) snrm[tid * SNS + in] = __float2half_rn(1.f);
    // the goal part of every predicted observation (env.obs_postproc) is constant along the horizon: written here, once
    for (int idx = tid; idx < (fused ? 0 : Gc * md.G); idx += kMlpThreads) {
      const int b = idx / md.G, c = idx % md.G;
      const float gv = __ldg(ra.goal + (size_t)((p0 + b) / ra.traj_per_goal) * md.G + c);
      for (int hh = 1; hh <= ra.H; ++hh) ra.traj[(((size_t)hh * md.E + e) * ra.n + p0 + b) * O + sd + c] = gv;
    }
    // this thread's half of the raw state lives in registers for the whole horizon
    float st[kMaxOW];
#pragma unroll
    for (int c = 0; c < kMaxOW; ++c) {
      const int d = hf * (tl.NOUT / 2) + c;
      st[c] = (live && c < tl.NOUT / 2 && d < sd) ? s_state[row * sd + d] : 0.f;
    }
    cta_sync();

    for (int h = 0; h < ra.H; ++h) {
      if (hf == 1 && live && h + 1 < ra.H) {  // prefetch the next step's action; consumed in the output epilogue
        for (int j = 0; j < U; ++j)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(s_act + row * 8 + j)),
                       "l"(ra.actions + ((size_t)(p0 + row) * ra.H + h + 1) * U + j)
                       : "memory");
      }
      // A operand of layer 0: this row's normalised input, the two halves write alternate 16-byte chunks
      {
        const uint4* sb = reinterpret_cast<const uint4*>(snrm + row * SNS);
        for (int c = hf; c < K1 / 8; c += 2) tmem_st4(t_a + 4 * c, sb[c]);
      }
      signal_a(cx, 0, lane);
      mma_layer(0);
      // row `row` of this step's predictions: full trajectories [H+1,E,n,O] or the trajectory's block of the planner scratch
      float* outr = fused ? fc.pred + (size_t)(p0 + row) * fc.ps + (size_t)h * fc.hs + (size_t)e * fc.es
                          : ra.traj + (((size_t)(h + 1) * md.E + e) * ra.n + p0 + row) * O;

#pragma unroll 1
      for (int m = 0; m <= L; ++m) {
        wait_bar(cx, d_ready, dph, 10 + m);
        dph ^= 1u;
        tc_fence_after();
        const bool warp_on = __any_sync(FULL, live);
        if (m < L) {
          // ---------------- hidden layer: activation -> fp16 -> next A operand ----------------
          if (warp_on) {
#pragma unroll
            for (int c0 = 0; c0 < W; c0 += 32) {
              float x[32];
              uint32_t hc[16];
              tmem_ld32(t_acc + hf * W + c0, *reinterpret_cast<uint32_t(*)[32]>(&x[0]));
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                hc[j / 2] = pack_half2(mlp_act<SILU>(x[j], act), mlp_act<SILU>(x[j + 1], act));
                // no LayerNorm bounds these activations: record a saturated conversion (SiLU >= -0.28: positive side only)
                satm = SILU ? sat_track_pos(satm, hc[j / 2]) : sat_track(satm, hc[j / 2]);
              }
              tmem_st16(t_a + (hf * W + c0) / 2, hc);
            }
          }
          signal_a(cx, 0, lane);
          mma_layer(m + 1);
        } else {
          // ---------------- output layer: delta -> de-normalise, residual update (mlp_ensemble.py:530-573) ----------------
          const int OW = tl.NOUT / 2;  // output columns per thread
          if (warp_on) {
#pragma unroll
            for (int cc = 0; cc < kMaxOW / 16; ++cc) {
              const int c0 = 16 * cc;
              if (c0 < OW) {
                uint32_t v[16];
                tmem_ld16(t_acc + hf * OW + c0, v);
                tmem_wait_ld();
                if (live) {
                  const int d0 = hf * OW + c0;
                  uint32_t pk[8];
#pragma unroll
                  for (int j = 0; j < 16; j += 2) {
                    // padded columns (d >= sd) have an all-zero table entry: state and normalised copy stay exactly 0
                    const float4 ta = s_tab[d0 + j], tb = s_tab[d0 + j + 1];
                    const float na = st[c0 + j] + fmaf(ta.x, __uint_as_float(v[j]), ta.y);
                    const float nb = st[c0 + j + 1] + fmaf(tb.x, __uint_as_float(v[j + 1]), tb.y);
                    st[c0 + j] = na; st[c0 + j + 1] = nb;
                    if (d0 + j < sd) outr[d0 + j] = na;
                    if (d0 + j + 1 < sd) outr[d0 + j + 1] = nb;
                    pk[j / 2] = pack_half2((na - ta.z) * ta.w, (nb - tb.z) * tb.w);
                    satm = sat_track(satm, pk[j / 2]);
                  }
                  __half* dst = snrm + row * SNS + d0;
                  if (d0 + 16 <= sd) {
                    *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(dst + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                  } else {  // the chunk that runs into the action / constant-one columns: element-wise
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                      if (d0 + j < sd) dst[j] = reinterpret_cast<const __half*>(pk)[j];
                  }
                }
              }
            }
          }
          if (hf == 1 && live && h + 1 < ra.H) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            for (int j = 0; j < U; ++j)
              satm = sat_track1(satm, snrm[row * SNS + sd + j] = half_sat((s_act[row * 8 + j] - nrm[md.nrm.in_all + sd + j]) * nrm[md.nrm.in_all + in + sd + j]));
          }
          tc_fence_before();
        }
      }
      cta_sync();  // snrm of the next step is complete
    }
    if (sat_hit(satm)) atomicOr(err, kTcRangeFlag);  // an fp16 operand left the representable range somewhere in this round
  }
  // ---- fused cost tail (planner launches), after the round loop so that nothing of the step loop is live here ----
  // This member's rollouts of the CTA's tiles are complete: warp w arrives for the trajectories w, w + 8, ... of every tile and
  // reduces those whose last ensemble member this was (disagreement + task cost + ensemble mean / std, cost_device.cuh), while the
  // predictions are still in L2.
  if (fused && fc.done != nullptr && fc.dbg_mode != 2) {
    __threadfence();
    cta_sync();
    float* envbuf = (128 * sd) / 8 >= md.E * ra.H ? s_state + warp * ((128 * sd) / 8) : nullptr;  // the state tile is free now
    // phase 0: arrive for the warp's trajectories; phase 1: reduce the ones assigned to this member (p % E == e) once every member
    // has arrived (all CTAs of this persistent grid are resident, so the wait cannot deadlock)
    for (int phase = 0; phase < 2; ++phase)
      for (int rd = 0; rd < rounds; ++rd) {
        const int w_lo = min(ra.n, worker * per_worker), w_hi = min(ra.n, w_lo + per_worker);
        const int p0 = w_lo + rd * T;
        const int Gc = max(0, min(T, w_hi - p0));
        for (int b = warp; b < Gc; b += kMlpThreads / 32) {
          const int p = p0 + b;
          if (phase == 0) {
            if (lane == 0) costdev::fused_arrive_nowait(fc, p);
          } else if (p % md.E == e) {
            costdev::fused_wait(fc, p, lane);
            const float* srow = ra.start_obs + (size_t)(p / ra.traj_per_start) * O;
            const float* grow = ra.goal + (size_t)(p / ra.traj_per_goal) * md.G;
            if (md.E <= 5) costdev::fused_reduce<5, 8>(fc, p, lane, srow, grow, envbuf);
            else costdev::fused_reduce<kMaxEnsemble, 4>(fc, p, lane, srow, grow, envbuf);
          }
        }
      }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) tmem_dealloc<2>(0u, 512);
}

// fp32 packed weights -> fp16 B-operand images with the bias in the constant-one K row
__global__ void mlp_tc_pack_kernel(const float* __restrict__ w32, const ModelDesc md, const TcMlpLayout tl, uint8_t* __restrict__ wimg) {
  const int e = blockIdx.y;
  const float* wm = w32 + (size_t)e * md.member_stride;
  int total = 0;
  for (int m = 0; m <= tl.L; ++m) total += tl.mat[m].K * tl.mat[m].N;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int m = 0, base = 0;
    for (int t = 0; t <= tl.L; ++t) {
      if (idx >= base + tl.mat[t].K * tl.mat[t].N) base += tl.mat[t].K * tl.mat[t].N;
      else { m = t; break; }
    }
    const TcMlpMat& M = tl.mat[m];
    const int k = (idx - base) / M.N, n = (idx - base) % M.N;
    const LayerDesc& ld = md.mlp[0].layer[m];
    const int kreal = m == 0 ? md.sd + md.U : tl.HD;  // the constant-one column follows the real inputs
    float v = 0.f;
    if (n < ld.N) {
      if (k < kreal) v = wm[ld.w_off + (size_t)k * ld.Npad + n];
      else if (k == kreal) v = wm[ld.b_off + n];
    }
    const int rk = n / M.nloc, nl = n % M.nloc;
    uint8_t* img = wimg + ((size_t)e * 2 + rk) * tl.image_bytes;
    reinterpret_cast<__half*>(img + M.off)[((size_t)(k / 8) * M.nloc + nl) * 8 + (k % 8)] = __float2half_rn(v);
  }
}

template <int HD, bool SILU>
cudaError_t launch_mlp_t(const ModelDesc& md, const RolloutArgs& ra, const TcMlpLayout& tl, const uint8_t* wimg, int* err_flag,
                         const costdev::FusedCost& fc, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(rollout_mlp_tc_kernel<HD, SILU>, cudaFuncAttributeMaxDynamicSharedMemorySize, tl.smem_bytes);
  if (e != cudaSuccess) return e;
  rollout_mlp_tc_kernel<HD, SILU><<<dim3(2 * tl.pairs), kMlpThreads, tl.smem_bytes, st>>>(md, ra, tl, wimg, err_flag, fc);
  return cudaGetLastError();
}

}  // namespace

bool make_mlp_tc_layout(const ModelDesc& md, int num_sms, TcMlpLayout* out) {
  TcMlpLayout t{};
  if (md.kind != CEEUS_MODEL_MLP || (md.Hd != 128 && md.Hd != 256) || md.layer_norm || md.L < 1 || md.L + 1 > kMaxLayers) return false;
  if (md.U > 8 || md.sd > 128) return false;
  t.HD = md.Hd;
  t.L = md.L;
  t.K1 = round_up(md.sd + md.U + 1, 16);
  t.NOUT = round_up(md.sd, 32);
  if (t.K1 > md.Hd || t.NOUT > md.Hd) return false;
  int off = 0;
  for (int m = 0; m <= md.L; ++m) {
    TcMlpMat& M = t.mat[m];
    M.K = m == 0 ? t.K1 : md.Hd + 16;
    M.N = m == md.L ? t.NOUT : md.Hd;
    M.nloc = M.N / 2;
    M.off = off;
    off += (M.K / 8) * M.nloc * 16;
  }
  t.image_bytes = off;
  t.nrm_floats = 2 * (md.sd + md.U) + 2 * md.sd;
  int sm = 0;
  t.sm_w = sm; sm += round_up(t.image_bytes, 1024);
  t.sm_nrm = sm; sm += round_up(t.nrm_floats * 4, 16);
  t.sm_state = sm; sm += round_up(128 * md.sd * 4, 16);
  t.sm_snrm = sm; sm += 128 * (t.K1 + 8) * 2;
  t.sm_tab = sm; sm += t.NOUT * 16;
  t.sm_act = sm; sm += 128 * 8 * 4;
  t.sm_bar = sm; sm += 64;
  t.smem_bytes = sm;
  if (t.smem_bytes > 227 * 1024) return false;
  t.pairs = num_sms / 2;
  if (t.pairs < md.E) return false;
  *out = t;
  return true;
}

cudaError_t launch_mlp_tc_pack(const float* w32, const ModelDesc& md, const TcMlpLayout& tl, uint8_t* wimg, cudaStream_t st) {
  mlp_tc_pack_kernel<<<dim3(128, md.E), 256, 0, st>>>(w32, md, tl, wimg);
  return cudaGetLastError();
}

cudaError_t launch_rollout_mlp_tc(const ModelDesc& md, const RolloutArgs& ra, const TcMlpLayout& tl, const uint8_t* wimg, int* err_flag,
                                  const costdev::FusedCost& fc, cudaStream_t st) {
  const bool silu = md.act == CEEUS_ACT_SILU;
  if (md.Hd == 256) return silu ? launch_mlp_t<256, true>(md, ra, tl, wimg, err_flag, fc, st) : launch_mlp_t<256, false>(md, ra, tl, wimg, err_flag, fc, st);
  return silu ? launch_mlp_t<128, true>(md, ra, tl, wimg, err_flag, fc, st) : launch_mlp_t<128, false>(md, ra, tl, wimg, err_flag, fc, st);
}

}  // namespace ceeus
