// CEEUS_PREC_TC: tensor-core (tcgen05) GNN rollout -- the whole horizon of GNNForwardEnsembleModel.predict_n_steps
// (mbrl/models/gnn_ensemble.py:876-983, gnn_modules.py:173-278) in one persistent launch.
//
// Execution model (DESIGN.md "rollout_gnn_tc_kernel"):
//   * grid = CTA PAIRS (cluster of 2, tcgen05 cta_group::2, M = 256).  A pair owns one ensemble member; every B operand
//     of the member (8 fp16 matrices, ~100 KB per CTA thanks to the split by output column) is resident in shared memory
//     for the whole kernel, staged by one bulk TMA copy.
//   * per CTA two (Hd = 128) or four (Hd = 64) warpgroups ("slots"), each owning 2*HD TMEM columns: the fp32 accumulator
//     (HD columns) and two fp16 A-operand regions R1, R2 (HD/2 columns each).  The A operand of EVERY MMA comes from TMEM
//     (tcgen05.mma "TS" form): the epilogue threads write the next layer's input with tcgen05.st, so activations never
//     touch shared memory.  Thread == tile row == TMEM lane in every epilogue.  One warp per slot of the leader CTA issues
//     the MMAs; completion is multicast-committed to both CTAs.
//   * tile rows are "node major": row == (trajectory, node), a warp owns 32 / N whole trajectories.  A slot takes its group of
//     trajectories through all H steps; per step N-1 edge passes (E1, E2; row (b, i) carries edge i -> j), the node tile
//     (N1, N2, N3) on the same rows and the global tile (G1, G2, G3) on the rows (b, 0).  The node aggregate is a running
//     fp16 sum in the row's own TMEM lane, the global aggregate a segmented warp-shuffle sum: no block barrier in the step loop.
//   * algebra done at pack time (tc_logical_kernel): the edge output layer has no LayerNorm / activation and the
//     aggregation is linear, so  aggr_j(h2_j W3 + b3) W_agg = aggr_j(h2_j) (W3 W_agg) + b3 W_agg  -- W3 is folded into the
//     first node / global layer and the E3 stage disappears; mean-aggregation scales are folded too.  Weights feeding a
//     LayerNorm are centred over the output columns, so the MMA result already has zero row mean and the epilogue only
//     needs the second moment; the LayerNorm affine is folded into the weights where every gamma != 0.  Biases ride in a
//     constant-one K column (or, for E2 of the N >= 3 graphs, are pre-loaded into the accumulator).
//   * fp32 object states live in REGISTERS of the row-owner threads (node row: D dynamic features, row (b, 0): agent);
//     shared memory holds the normalised fp16 copy the A tiles are built from and the warp's staged next observations,
//     which leave as one coalesced copy per step while an MMA stage runs.
//   * operands are fp16 (10-bit mantissa, TF32 class) with fp32 accumulation; states, residual update, normalisers and
//     LayerNorm statistics stay fp32.  Conversions saturate (never inf) and a saturation is recorded in the status word
//     (kTcRangeFlag).  Parity bar: 1e-3 (BASELINE.json north_star, tf32/bf16 mode).
//   * second tile row layout (template parameter GROW, TcLayout::grow, picked per handle by tc_pick_grow): a trajectory owns N node
//     rows + 1 GLOBAL row; node MLP and global MLP run K-stacked as one stage each (NG1, NG2, NG3), their raw-input and bias K
//     blocks come from shared-memory A tiles (SS-form MMAs accumulating into the same tile as the TS-form ones): 2(N-1) + 3
//     stages per model step instead of 2(N-1) + 6, for launches whose step time is the stage chain, not the tile capacity.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "cost_device.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"
#include "tc_rollout_common.cuh"

namespace ceeus {

namespace {
using namespace tc;
static_assert(kTcRangeFlag == CEEUS_TC_RANGE, "the kernels' range bit is the one include/ceeus_b200.h documents");

// Warpgroups ("slots") per CTA: each owns 2 * HD TMEM columns, so the 512 columns hold two slots at Hd = 128 (256 threads, up to
// 255 registers each: the epilogue keeps a whole 128-column row in registers) and four at Hd = 64 (512 threads, 128 registers).
// Warp 0 of every slot also issues that slot's MMAs (leader CTA): there are no dedicated issuer warps.
template <int HD>
struct TcSlots { static constexpr int value = 512 / (2 * HD); };
constexpr int kFold = 100000;    // kmap code >= kFold: folded aggregate row (index code - kFold)

__device__ __forceinline__ void wg_sync(int s) { asm volatile("bar.sync %0, 128;" ::"r"(1 + s) : "memory"); }

template <int HD>
__device__ __forceinline__ void load_acc(uint32_t taddr, float (&x)[HD]) {
#pragma unroll
  for (int c0 = 0; c0 < HD; c0 += 32) tmem_ld32(taddr + c0, *reinterpret_cast<uint32_t(*)[32]>(&x[c0]));
}

// HD/2 packed fp16 pairs -> TMEM columns [taddr, taddr + HD/2)
template <int HD>
__device__ __forceinline__ void store_hidden(uint32_t taddr, const uint32_t (&h)[HD / 2]) {
#pragma unroll
  for (int c0 = 0; c0 < HD / 2; c0 += 32) tmem_st32(taddr + c0, *reinterpret_cast<const uint32_t(*)[32]>(&h[c0]));
}

// One-pass hidden-layer epilogue on the whole accumulator row held in registers (255 registers per thread are available with
// 256 threads per CTA).  FOLDED LayerNorm (fold != 0, ReLU): sign(gamma) lives in this layer's weight columns and |gamma| in the
// consumer's weight rows (pack time), so  relu(gamma * xhat + beta) = |gamma| * relu(sign(gamma) * xhat + beta / |gamma|)  costs
// one FFMA per element and one parameter vector (pbe = beta / |gamma|).  General path: xhat * gamma + beta, any activation.
template <int HD, bool RELU, bool FOLD>
__device__ __forceinline__ void epi_row(uint32_t tacc, const float* __restrict__ pb, const float* __restrict__ pg,
                                        const float* __restrict__ pbe, bool ln, bool fold, int act, uint32_t (&h)[HD / 2],
                                        uint32_t& satm) {
  float x[HD];
#pragma unroll
  for (int c0 = 0; c0 < HD; c0 += 32) tmem_ld32(tacc + c0, *reinterpret_cast<uint32_t(*)[32]>(&x[c0]));
  tmem_wait_ld();
  if (pb != nullptr) {
#pragma unroll
    for (int j = 0; j < HD; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(pb + j);
      upk2(fadd2(pk2(x[j], x[j + 1]), pk2(b.x, b.y)), x[j], x[j + 1]);
      upk2(fadd2(pk2(x[j + 2], x[j + 3]), pk2(b.z, b.w)), x[j + 2], x[j + 3]);
    }
  }
  if (ln) {
    // second moment with packed fp32 FMAs (FFMA2): four independent accumulator pairs
    uint64_t v2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
    for (int j = 0; j < HD; j += 2) {
      const uint64_t xp = pk2(x[j], x[j + 1]);
      v2[(j >> 1) & 3] = ffma2(xp, xp, v2[(j >> 1) & 3]);
    }
    float v8[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) upk2(v2[j], v8[2 * j], v8[2 * j + 1]);
    const float v = ((v8[0] + v8[1]) + (v8[2] + v8[3])) + ((v8[4] + v8[5]) + (v8[6] + v8[7]));
    const float rstd = rsqrtf(v * (1.f / HD) + 1e-5f);
    if (FOLD || (RELU && fold)) {
      const uint64_t r2 = pk2(rstd, rstd);
#pragma unroll
      for (int j = 0; j < HD; j += 4) {
        const float4 be = *reinterpret_cast<const float4*>(pbe + j);
        float a0, a1, a2, a3;
        upk2(ffma2(pk2(x[j], x[j + 1]), r2, pk2(be.x, be.y)), a0, a1);
        upk2(ffma2(pk2(x[j + 2], x[j + 3]), r2, pk2(be.z, be.w)), a2, a3);
        h[j / 2] = pack_half2_relu(a0, a1);
        h[j / 2 + 1] = pack_half2_relu(a2, a3);
      }
      return;
    }
    if constexpr (FOLD) return;  // (unreachable: FOLD kernels only run members whose LayerNorm affine is folded)
#pragma unroll
    for (int j = 0; j < HD; j += 4) {
      const float4 g = *reinterpret_cast<const float4*>(pg + j);
      const float4 be = *reinterpret_cast<const float4*>(pbe + j);
      x[j] = fmaf(x[j] * rstd, g.x, be.x); x[j + 1] = fmaf(x[j + 1] * rstd, g.y, be.y);
      x[j + 2] = fmaf(x[j + 2] * rstd, g.z, be.z); x[j + 3] = fmaf(x[j + 3] * rstd, g.w, be.w);
    }
  }
  if constexpr (RELU) {
#pragma unroll
    for (int j = 0; j < HD; j += 2) h[j / 2] = pack_half2_relu(x[j], x[j + 1]);
  } else {
#pragma unroll 4
    for (int j = 0; j < HD; j += 2) h[j / 2] = pack_half2(act_fn(x[j], act), act_fn(x[j + 1], act));
  }
  // without a LayerNorm nothing bounds the hidden activations: record a saturated conversion (with one, |xhat| <= sqrt(HD))
  if (!ln) {
#pragma unroll
    for (int j = 0; j < HD / 2; ++j) satm = sat_track(satm, h[j]);
  }
}

// MMA issue of one stage (leader CTA, one warp): wait for both CTAs' A operands, issue the K steps, commit to both CTAs.
// (Out of line it was slower: the call spills the 255-register frame of the issuing warp.)  The stage
// kind is a compile-time constant: K extents are fixed by the operand geometry (KE1 = 48, KXn = 32, KXg = 16), so descriptors
// and TMEM addresses are uniform-register arithmetic and the K loop is fully unrolled.
template <int HD, int KIND>
__device__ __forceinline__ void tc_issue_stage(volatile int* abort_flag, int* err, uint64_t* a_bar, uint32_t aph, uint64_t* d_bar,
                                            uint32_t wdesc_addr, uint32_t m_acc, int n2) {
  Ctx cx;
  cx.abort = abort_flag; cx.err = err;
  const bool ok = wait_bar_spin(cx, a_bar, aph, 2);
  tc_fence_after();
  if (ok) {  // warp-converged: one elected lane issues, operands stay warp uniform
    const uint32_t elected = elect_one();
    constexpr bool outl = KIND == TC_N3 || KIND == TC_G3;
    constexpr int kx = (KIND == TC_E1 ? 48 : KIND == TC_N1 ? 32 : KIND == TC_G1 ? 16 : HD) / 16;
    constexpr int kmax = KIND == TC_E1 ? 3 : outl ? HD / 16 : KIND == TC_N1 ? 2 + HD / 16 : 1 + HD / 16;
    const int kall = KIND == TC_E2 ? HD / 16 + (n2 ? 1 : 0) : kmax;  // E2: + the constant-one block only when R2 is free (N == 2)
    constexpr uint32_t idesc = outl ? make_idesc_f16(256, 16) : make_idesc_f16(256, HD);
    constexpr uint32_t lbo_b = (uint32_t)(outl ? 8 : HD / 2) * 16;
    const uint64_t db0 = make_smem_desc(wdesc_addr, lbo_b, 128);
    constexpr uint32_t db_step = (2u * lbo_b) >> 4;  // descriptor start-address field advances by two 16-byte-chunk rows per K step
    const uint32_t m_r1 = m_acc + HD, m_r2 = m_r1 + HD / 2;
#pragma unroll
    for (int ks = 0; ks < kmax; ++ks) {
      if (ks < kall) {
        const uint32_t a_col = ks < kx ? m_r1 + (uint32_t)(8 * ks) : m_r2 + (uint32_t)(8 * (ks - kx));
        // E2 with N >= 3: the accumulator was pre-initialised with the bias by the E1 epilogue (no constant-one block: R2 is busy)
        const uint32_t accum = ks > 0 ? 1u : (KIND == TC_E2 && !n2) ? 1u : 0u;
        umma_f16_ts_cg2_elect(m_acc, a_col, db0 + (uint64_t)((uint32_t)ks * db_step), idesc, accum, elected);
      }
    }
    umma_commit_cg2_mc_elect(d_bar, 0x3, elected);
  }
  __syncwarp();
}

// "Global row" tiling (TcLayout::grow): stage kinds of one model step.  The node MLP (rows (b, i)) and the global MLP (rows (b, N))
// run as ONE stage each: the two layers are stacked along K -- a row presents its input in the K block of its own role and zeros in
// the other role's block, so one accumulation chain applies N-weights to node rows and G-weights to global rows.
enum { GK_E1 = 0, GK_E2, GK_NG1, GK_NG2, GK_NG3 };

// MMA issue of one stage of the global-row tiling.  TMEM A blocks: R1 = [m_acc + HD, +HD/2) and R2 = [m_acc + 3 HD / 2, +HD/2).
//   E1 : R1[0:48 halves) = [hdr | x_i | x_j]                                  x E1
//   E2 : R1 = hidden of E1                                                     x E2[0:HD)      + one(node rows) x E2[HD:HD+16) (bias)
//   NG1: smem [hdr_N | x_i] x N1[0:32), smem [hdr_G] x G1[0:16), R2 (node aggregate) x N1[32:), R1 (global aggregate) x G1[16:)
//   NG2: R2 x N2[0:HD), R1 x G2[0:HD), one(node) x N2[HD:+16), one(global) x G2[HD:+16)
//   NG3: R2 x N3, R1 x G3 (16 output columns)
// Node rows keep R1 = 0 and global rows R2 = 0 from the last E2 epilogue to the end of the step.
template <int HD, int KIND>
__device__ __forceinline__ void tc_issue_grow(volatile int* abort_flag, int* err, uint64_t* a_bar, uint32_t aph, uint64_t* d_bar,
                                              uint32_t wbase, const TcLayout& tl, uint32_t m_acc, uint32_t xa_addr, uint32_t one_addr) {
  Ctx cx;
  cx.abort = abort_flag; cx.err = err;
  const bool ok = wait_bar_spin(cx, a_bar, aph, 2);
  tc_fence_after();
  if (ok) {
    const uint32_t elected = elect_one();
    constexpr int KH = HD / 16;  // K steps of a hidden row
    constexpr uint32_t idesc_h = make_idesc_f16(256, HD), idesc_o = make_idesc_f16(256, 16);
    constexpr uint32_t lbo_h = (uint32_t)(HD / 2) * 16, lbo_o = 8u * 16u;
    constexpr uint32_t st_h = (2u * lbo_h) >> 4, st_o = (2u * lbo_o) >> 4;  // descriptor start-address advance per K step
    constexpr uint32_t st_a = (2u * 2048u) >> 4;                            // shared-memory A tiles: 2048 bytes per 8-half K chunk
    const uint32_t m_r1 = m_acc + HD, m_r2 = m_r1 + HD / 2;
    uint32_t acc = 0u;
    auto ts = [&](uint32_t a_col, uint64_t db, uint32_t idesc) { umma_f16_ts_cg2_elect(m_acc, a_col, db, idesc, acc, elected); acc = 1u; };
    auto ss = [&](uint64_t da, uint64_t db, uint32_t idesc) { umma_f16_ss_cg2_elect(m_acc, da, db, idesc, acc, elected); acc = 1u; };
    auto bh = [&](int m) { return make_smem_desc(wbase + (uint32_t)tl.mat[m].off, lbo_h, 128); };
    auto bo = [&](int m) { return make_smem_desc(wbase + (uint32_t)tl.mat[m].off, lbo_o, 128); };
    if constexpr (KIND == GK_E1) {
      const uint64_t db = bh(TC_E1);
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) ts(m_r1 + 8u * ks, db + (uint64_t)(ks * st_h), idesc_h);
    } else if constexpr (KIND == GK_E2) {
      const uint64_t db = bh(TC_E2);
#pragma unroll
      for (int ks = 0; ks < KH; ++ks) ts(m_r1 + 8u * ks, db + (uint64_t)(ks * st_h), idesc_h);
      ss(make_smem_desc(one_addr, 2048, 128), db + (uint64_t)(KH * st_h), idesc_h);
    } else if constexpr (KIND == GK_NG1) {
      const uint64_t dn = bh(TC_N1), dg = bh(TC_G1), da = make_smem_desc(xa_addr, 2048, 128);
      ss(da, dn, idesc_h);
      ss(da + st_a, dn + st_h, idesc_h);
      ss(da + 2 * st_a, dg, idesc_h);
#pragma unroll
      for (int j = 0; j < KH; ++j) ts(m_r2 + 8u * j, dn + (uint64_t)((2 + j) * st_h), idesc_h);
#pragma unroll
      for (int j = 0; j < KH; ++j) ts(m_r1 + 8u * j, dg + (uint64_t)((1 + j) * st_h), idesc_h);
    } else if constexpr (KIND == GK_NG2) {
      const uint64_t dn = bh(TC_N2), dg = bh(TC_G2), d1 = make_smem_desc(one_addr, 2048, 128);
#pragma unroll
      for (int j = 0; j < KH; ++j) ts(m_r2 + 8u * j, dn + (uint64_t)(j * st_h), idesc_h);
#pragma unroll
      for (int j = 0; j < KH; ++j) ts(m_r1 + 8u * j, dg + (uint64_t)(j * st_h), idesc_h);
      ss(d1, dn + (uint64_t)(KH * st_h), idesc_h);
      ss(d1 + st_a, dg + (uint64_t)(KH * st_h), idesc_h);
    } else {
      const uint64_t dn = bo(TC_N3), dg = bo(TC_G3);
#pragma unroll
      for (int j = 0; j < KH; ++j) ts(m_r2 + 8u * j, dn + (uint64_t)(j * st_o), idesc_o);
#pragma unroll
      for (int j = 0; j < KH; ++j) ts(m_r1 + 8u * j, dg + (uint64_t)(j * st_o), idesc_o);
    }
    umma_commit_cg2_mc_elect(d_bar, 0x3, elected);
  }
  __syncwarp();
}

// Group of (pair, cluster rank, slot) in round rd: out of line on purpose -- evaluated once per round, and inlined its operands
// (the whole TcTile) would stay live across the 255-register step loop.  Returns the packed pair (p0, Gc).
__device__ __noinline__ int2 tc_group_of(int pairs, int E, int pair, int n, int NS, int Tmax, int tpw, int rank, int s, int rd,
                                         int even_rounds) {
  const TcTile t = tc_tile(pairs, E, pair, n, NS, Tmax, tpw);
  int p0, Gc;
  tc_group(t, n, NS, tpw, rank, s, rd, even_rounds, &p0, &Gc);
  return make_int2(p0, Gc);
}

// ------------------------------------------------------------------------------------------------------------
// Row layout ("node major"): TMEM lane == tile row == (trajectory b, node i); a warp owns 32 / N whole trajectories, so every
// exchange between the rows of a trajectory stays inside one warp (__syncwarp, no block barriers in the step loop).
//   * edge pass jj (0 .. N-2): row (b, i) carries the edge i -> j, j = jj + (jj >= i): E1, E2.  The sum over the outgoing edges
//     of node i is a running fp16 sum in the row's own TMEM lane (R2), i.e. the node aggregate needs no communication at all.
//   * the sum over all edges of a trajectory (global aggregate) is formed once per step from the N node aggregates through a
//     warp-private staging buffer and parked in shared memory until the global tile needs it.
//   * node tile rows == the same lanes; global tile row of trajectory b == lane (b, 0).
template <int HD, bool RELU, bool DBG, bool FOLD, bool GROW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 * TcSlots<HD>::value, 1)
    rollout_gnn_tc_kernel(const ModelDesc md, const RolloutArgs ra, const TcLayout tl, const uint8_t* __restrict__ wimg,
                          const float* __restrict__ wpar, int* __restrict__ err, long long* __restrict__ dbg,
                          const costdev::FusedCost fc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const bool fused = fc.pred != nullptr;  // planner launch: predictions -> compact scratch, costs reduced in the tail
  constexpr int CH = HD / 8;  // 16-byte chunks per hidden row
  constexpr int NS = TcSlots<HD>::value;
  constexpr int kTcThreads = 128 * NS;
  constexpr int kAU = 16, kNA = 16;  // halves: header [agent | action | 1 | 0..], node block [dyn | stat | 0..]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int N = md.N, A = md.A, D = md.D, S = md.S, U = md.U, O = md.O;

  // ---- member / trajectory range of this pair: pairs are dealt to members as evenly as the count allows ----
  // Uneven split: the MMA-issuing warp is the last-filled warp of the leader CTA's slots.  Where the even split would keep it busy
  // and the same number of rounds also fits with that warp left empty (leader slots 3 warps' worth of trajectories, the other
  // CTA's slots 4), the issuer stays a pure issuer that reacts at once.  q = trajectories per warp and round.
  const TcTile tt = tc_tile(tl.pairs, md.E, pair, ra.n, NS, tl.Tmax, tl.tpw);
  const int e = tt.e, rounds = tt.rounds;
  const int n_stage = 2 * (N - 1) + (GROW ? 3 : 6);
  const int R = N + (GROW ? 1 : 0);  // tile rows per trajectory

  uint8_t* sW = smem + tl.sm_w;
  float* sPar = reinterpret_cast<float*>(smem + tl.sm_par);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + tl.sm_bar);
  uint64_t* wbar = bars;        // weights landed
  uint64_t* a_ready = bars + 1;       // [NS] (used in the leader CTA)
  uint64_t* d_ready = bars + 1 + NS;  // [NS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * NS);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  // ---- one-time setup ----
  if (tid == 0) *abort_flag = 0;
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(wbar, 1);
      for (int i = 0; i < NS; ++i) {
        mbar_init(&a_ready[i], 8);
        mbar_init(&d_ready[i], 1);
      }
      fence_barrier_init();
      mbar_expect_tx(wbar, (uint32_t)tl.image_bytes);
      bulk_g2s(sW, wimg + ((size_t)e * 2 + rank) * tl.image_bytes, (uint32_t)tl.image_bytes, wbar);
    }
    __syncwarp();
    tmem_alloc<2>(tmem_slot, 512);
    tmem_relinquish<2>();
  }
  for (int i = tid; i < tl.par_floats; i += kTcThreads) sPar[i] = __ldg(wpar + (size_t)e * tl.par_floats + i);
  // normalisers -> shared memory; the std of the INPUT groups is stored as its reciprocal (x_n = (x - mean) * inv_std)
  float* sNrm = reinterpret_cast<float*>(smem + tl.sm_nrm);
  for (int i = tid; i < tl.nrm_floats; i += kTcThreads) {
    const float v = __ldg(ra.norm + i);
    const bool inv = (i >= md.nrm.in_agent + md.A && i < md.nrm.in_dyn) || (i >= md.nrm.in_dyn + md.D && i < md.nrm.in_stat) ||
                     (i >= md.nrm.in_stat + md.S && i < md.nrm.in_action) || (i >= md.nrm.in_action + md.U && i < md.nrm.out_agent);
    sNrm[i] = inv ? 1.f / v : v;
  }
  if constexpr (GROW) {
    // constant A tile of the bias K blocks: chunks 0-1 = [1, 0 ..] on node rows, chunks 2-3 = [1, 0 ..] on global rows; the
    // slots' [hdr | x_i | hdr] tiles start out zero (a row only ever writes the chunks of its own role)
    uint8_t* sOne = smem + tl.sm_one;
    for (int i = tid; i < 4 * 128; i += kTcThreads) {
      const int chunk = i >> 7, r = i & 127, ln = r & 31;
      const int bwr = ln / R, nir = ln - bwr * R;
      const bool okr = bwr < tl.tpw, isg = nir == N;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (okr && ((chunk == 0 && !isg) || (chunk == 2 && isg))) v.x = 0x00003c00u;
      *reinterpret_cast<uint4*>(sOne + chunk * 2048 + r * 16) = v;
    }
    for (int sl = 0; sl < NS; ++sl)
      for (int i = tid; i < 6 * 128; i += kTcThreads)
        *reinterpret_cast<uint4*>(smem + tl.sm_slot[sl] + tl.so_xa + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    float* sNegI = reinterpret_cast<float*>(smem + tl.sm_neg);
    for (int i = tid; i < HD; i += kTcThreads) sNegI[i] = -60000.f;
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  // output-layer tables (float4-friendly, 16 entries each): new = old + c1 * acc + c0 with c1 = std_out, c0 = std_out * bias +
  // mean_out; the re-normalised fp16 copy is (new - mean_in) * inv_std_in.  [0]: node rows (dyn), [1]: global rows (agent)
  float* sOut = reinterpret_cast<float*>(smem + tl.sm_out);
  for (int i = tid; i < 2 * 4 * 16; i += kTcThreads) {
    const int which = i / 64, arr = (i / 16) % 4, j = i % 16;
    const int dim = which == 0 ? md.D : md.A;
    const int o_out = which == 0 ? md.nrm.out_dyn : md.nrm.out_agent, o_in = which == 0 ? md.nrm.in_dyn : md.nrm.in_agent;
    const float* pbias = sPar + tl.mat[which == 0 ? TC_N3 : TC_G3].par_bias;
    float v = 0.f;
    if (j < dim) {
      const float sdv = sNrm[o_out + dim + j];
      v = arr == 0 ? sdv : arr == 1 ? fmaf(sdv, pbias[j], sNrm[o_out + j]) : arr == 2 ? sNrm[o_in + j] : sNrm[o_in + dim + j];
    }
    sOut[i] = v;
  }
  __syncthreads();
  Ctx cx;
  cx.abort = abort_flag; cx.err = err;
  if (warp == 0 && lane == 0) wait_bar(cx, wbar, 0, 1);  // this CTA's weight image has landed
  __syncthreads();
  cluster_sync();  // both CTAs: barriers initialised, TMEM allocated, weights resident
  tc_fence_after();
  // all 512 columns are allocated, so the base can only be column 0 / lane 0; using the literal keeps every TMEM address
  // warp uniform for the compiler (checked once)
  if (*tmem_slot != 0u && tid == 0) atomicExch(err, 7);
  constexpr uint32_t tmem_base = 0u;
  for (int i = 0; i < NS; ++i) cx.a_ready_leader[i] = mapa(smem_u32(&a_ready[i]), 0);

  {
    // =========================== warpgroup: tile builder + epilogue ===========================
    const int s = warp >> 2;       // slot
    const int q = warp & 3;        // TMEM lane quadrant
    const int row = tid & 127;     // tile row == TMEM lane
    uint8_t* slot = smem + tl.sm_slot[s];
    float* s_obs = reinterpret_cast<float*>(slot + tl.so_obs + q * tl.obs_bytes);  // this warp's next observations [tpw][O] fp32
    uint8_t* s_glob = slot + tl.so_glob;  // [Tmax][HD] fp16 sums over all edges of a trajectory
    __half* snrm = reinterpret_cast<__half*>(slot + tl.so_snrm);  // [Tmax][SNs] normalised [hdr | node blocks]
    float* s_tail = reinterpret_cast<float*>(slot + tl.so_tail);  // [Tmax][tail] static features + goal (fp32)
    const int SNs = tl.SNs, tail = tl.tail, tpw = tl.tpw;
    const float* __restrict__ nrm = sNrm;
    const bool ln = md.layer_norm != 0;
    const int act = md.act;
    // LayerNorm affine folded into the weights: decided per member at pack time; the FOLD instantiation is launched when every
    // member folds (the host reads the flags after packing) and does not contain the general LayerNorm code at all
    const bool fold = FOLD || sPar[tl.par_fold] != 0.f;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 2 * HD);
    const uint32_t t_acc = lane_base, t_r1 = lane_base + HD, t_r2 = lane_base + HD + HD / 2;
    uint32_t dph = 0;
    int dbg_n = 0;
    // profiling flag (ceeus_tc_debug_flag): bit 0 = slot 1 idles, bits 1..9 = probed thread, bits 12..15 = first probed round,
    // bits 16.. = probed CTA
    const bool dbg_on = DBG && dbg != nullptr && (long long)blockIdx.x == (dbg[499] >> 16) && tid == (int)((dbg[499] >> 1) & 0x1ff);
    int dbg_rd = 0;
    auto stamp = [&](int tag) {
      if constexpr (!DBG) return;  // the timeline probes exist only in the profiling instantiation
      if (dbg_on && dbg_rd >= (int)((dbg[499] >> 12) & 0xf) && dbg_n < 250) { dbg[2 * dbg_n] = tag; dbg[2 * dbg_n + 1] = clock64(); ++dbg_n; }
    };

    // ---- static row roles ----
    // The warps of a slot are filled with trajectories in order; odd slots fill from the last scheduler backwards, so that at
    // partial fill the busy warps of neighbouring slots sit on different schedulers.
    const int qf = (s & 1) ? 3 - q : q;
    const int bw = lane / R, ni = lane - bw * R;  // trajectory inside the warp, node (GROW: ni == N is the trajectory's global row)
    const bool lane_ok = bw < tpw;
    const bool is_g = GROW && (ni == N || !lane_ok);  // idle lanes behave like (dead) global rows
    const int b = qf * tpw + min(bw, tpw - 1);  // group-local trajectory of this row (clamped for the idle lanes)
    const int nix = GROW ? min(ni, N - 1) : ni;  // (global rows ride through the edge stages on finite filler data)
    const float* sNeg = reinterpret_cast<const float*>(smem + tl.sm_neg);
    uint8_t* xa_row = slot + tl.so_xa + row * 16;

    // ---- MMA issue: the last-filled warp of the warpgroup in the leader CTA waits for both CTAs' A operands and issues ----
    const bool issuer_warp = rank == 0 && qf == 3;
    uint32_t aph = 0;
    const uint32_t wbase = smem_u32(sW);
    const uint32_t m_acc = tmem_base + (uint32_t)(s * 2 * HD);
    auto mma_stage = [&](int kind) {
      if (!issuer_warp) return;
      stamp(70 + kind);
      if constexpr (GROW) {
        const uint32_t xa = smem_u32(slot + tl.so_xa), one = smem_u32(smem + tl.sm_one);
        switch (kind) {
          case GK_E1: tc_issue_grow<HD, GK_E1>(abort_flag, err, &a_ready[s], aph, &d_ready[s], wbase, tl, m_acc, xa, one); break;
          case GK_E2: tc_issue_grow<HD, GK_E2>(abort_flag, err, &a_ready[s], aph, &d_ready[s], wbase, tl, m_acc, xa, one); break;
          case GK_NG1: tc_issue_grow<HD, GK_NG1>(abort_flag, err, &a_ready[s], aph, &d_ready[s], wbase, tl, m_acc, xa, one); break;
          case GK_NG2: tc_issue_grow<HD, GK_NG2>(abort_flag, err, &a_ready[s], aph, &d_ready[s], wbase, tl, m_acc, xa, one); break;
          default: tc_issue_grow<HD, GK_NG3>(abort_flag, err, &a_ready[s], aph, &d_ready[s], wbase, tl, m_acc, xa, one); break;
        }
      } else {
        const uint32_t wd = wbase + (uint32_t)tl.mat[kind].off;
        const int n2 = N == 2;
        switch (kind) {
          case TC_E1: tc_issue_stage<HD, TC_E1>(abort_flag, err, &a_ready[s], aph, &d_ready[s], wd, m_acc, n2); break;
          case TC_E2: tc_issue_stage<HD, TC_E2>(abort_flag, err, &a_ready[s], aph, &d_ready[s], wd, m_acc, n2); break;
          case TC_N1: tc_issue_stage<HD, TC_N1>(abort_flag, err, &a_ready[s], aph, &d_ready[s], wd, m_acc, n2); break;
          case TC_N2: tc_issue_stage<HD, TC_N2>(abort_flag, err, &a_ready[s], aph, &d_ready[s], wd, m_acc, n2); break;
          case TC_N3: tc_issue_stage<HD, TC_N3>(abort_flag, err, &a_ready[s], aph, &d_ready[s], wd, m_acc, n2); break;
          case TC_G1: tc_issue_stage<HD, TC_G1>(abort_flag, err, &a_ready[s], aph, &d_ready[s], wd, m_acc, n2); break;
          case TC_G2: tc_issue_stage<HD, TC_G2>(abort_flag, err, &a_ready[s], aph, &d_ready[s], wd, m_acc, n2); break;
          default: tc_issue_stage<HD, TC_G3>(abort_flag, err, &a_ready[s], aph, &d_ready[s], wd, m_acc, n2); break;
        }
      }
      aph ^= 1u;
      stamp(90 + kind);
    };
    // stages whose successor takes its bias from a constant-one K block parked in R2 (bit = stage kind)
    uint32_t cb_mask = 0;
    for (int k = 0; k + 1 < TC_NMAT; ++k)
      if (tl.mat[k + 1].K > tl.mat[k + 1].kx && tl.mat[k + 1].kx == HD) cb_mask |= 1u << k;
    auto kind_of = [&](int st) { return st < 2 * (N - 1) ? (st & 1) : 2 + (st - 2 * (N - 1)); };
    auto parp = [&](int off) -> const float* { return off >= 0 ? sPar + off : nullptr; };

    // fp32 state of the row (node row: D dynamic features, row (b, 0): agent) and the prefetched next action.  Hd = 128 (255
    // registers per thread): registers.  Hd = 64 (128 registers per thread, four slots): the state is read back from the warp's
    // staged observation rows s_obs and the action is prefetched with cp.async into s_actn -- 40 registers less, no spills
    // (measured: Hd = 64 13 % faster this way, Hd = 128 7 % slower).
    constexpr bool kStateInSmem = HD == 64;
    float* s_actn = reinterpret_cast<float*>(slot + tl.so_actn) + (size_t)row * 8;
    float dyn[16], agent[16], act_next[8];
    uint32_t satm = 0u;  // record of saturated fp16 operand conversions of this thread (sat_track)
#pragma unroll
    for (int j = 0; j < 16; ++j) { dyn[j] = 0.f; agent[j] = 0.f; }  // padded entries take part in the vectorised output update

    for (int rd = 0; rd < rounds; ++rd) {
      if constexpr (DBG) dbg_rd = rd;
      // first trajectory / live trajectories of this group (Gc == 0: pure lock-step filler)
      const int2 grp = tc_group_of(tl.pairs, md.E, pair, ra.n, NS, tl.Tmax, tpw, (int)rank, s, rd, tl.even_rounds);
      const int p0 = grp.x;
      int Gc = grp.y;
      if (DBG && dbg != nullptr && (dbg[499] & 1) && s == 1) Gc = 0;  // profiling aid: slot 1 idles, slot 0 runs without SIMT contention
      const bool live = lane_ok && b < Gc;
      const bool live_n = GROW ? (live && !is_g) : live;              // row carries a node of a live trajectory
      const bool live_g = GROW ? (live && is_g) : (live && ni == 0);  // row carries the agent / global features of a live trajectory
      const bool warp_on = qf * tpw < Gc;  // this warp owns at least one live trajectory
      const uint4* sb = reinterpret_cast<const uint4*>(snrm + b * SNs);  // this row's trajectory: [hdr(2) | node blocks (2 each)]
      wg_sync(s);
      // ---- group init: start observation -> registers, snrm, traj[0]; the trailing observation part for every step ----
      for (int idx = row; idx < (tl.Tmax * SNs) / 8; idx += 128) reinterpret_cast<uint4*>(snrm)[idx] = make_uint4(0u, 0u, 0u, 0u);
      if (!fused)
        for (int idx = row; idx < Gc * O; idx += 128) {
          const int bb = idx / O, c = idx % O;
          const float v = __ldg(ra.start_obs + (size_t)((p0 + bb) / ra.traj_per_start) * O + c);
          ra.traj[((size_t)e * ra.n + p0 + bb) * O + c] = v;
        }
      for (int idx = row; idx < Gc * tail; idx += 128) {
        const int bb = idx / tail, c = idx % tail;
        s_tail[idx] = c < N * S ? __ldg(ra.start_obs + (size_t)((p0 + bb) / ra.traj_per_start) * O + A + N * D + c)
                                : __ldg(ra.goal + (size_t)((p0 + bb) / ra.traj_per_goal) * md.G + (c - N * S));
      }
      wg_sync(s);
      // static features + goal (env.obs_postproc) are constant along the horizon: placed once in the staged observation rows
      for (int idx = lane; idx < tpw * tail; idx += 32) {
        const int tb = idx / tail, c = idx % tail;
        s_obs[tb * O + (A + N * D) + c] = qf * tpw + tb < Gc ? s_tail[(qf * tpw + tb) * tail + c] : 0.f;
      }
      if (live_n) {
        const float* so = ra.start_obs + (size_t)((p0 + b) / ra.traj_per_start) * O;
        __half* dst = snrm + b * SNs + kAU + ni * kNA;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < D) {
            const float v0 = __ldg(so + A + ni * D + j);
            dyn[j] = v0;
            s_obs[bw * O + A + ni * D + j] = v0;
            dst[j] = half_sat((v0 - nrm[md.nrm.in_dyn + j]) * nrm[md.nrm.in_dyn + D + j]);
            satm = sat_track1(satm, dst[j]);
          }
        for (int j = 0; j < S; ++j)
          satm = sat_track1(satm, dst[D + j] = half_sat((__ldg(so + A + N * D + ni * S + j) - nrm[md.nrm.in_stat + j]) * nrm[md.nrm.in_stat + S + j]));
      }
      if (live_g) {
        const float* so = ra.start_obs + (size_t)((p0 + b) / ra.traj_per_start) * O;
        __half* dst = snrm + b * SNs;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < A) {
            const float v0 = __ldg(so + j);
            if constexpr (GROW) dyn[j] = v0; else agent[j] = v0;  // (GROW: a row has ONE role, its state lives in dyn[])
            s_obs[bw * O + j] = v0;
            satm = sat_track1(satm, dst[j] = half_sat((v0 - nrm[md.nrm.in_agent + j]) * nrm[md.nrm.in_agent + A + j]));
          }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < U) {
            const float a0 = __ldg(ra.actions + ((size_t)(p0 + b) * ra.H) * U + j);
            satm = sat_track1(satm, dst[A + j] = half_sat((a0 - nrm[md.nrm.in_action + j]) * nrm[md.nrm.in_action + U + j]));
          }
        dst[A + U] = __float2half_rn(1.f);
      }
      wg_sync(s);

      if constexpr (GROW) {
        // TMEM A blocks start out zero: global rows keep R2 = 0 for the whole group, and nothing uninitialised (NaN patterns)
        // can reach a block that another role's weights are applied to
        if (warp_on) {
          uint32_t z[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) z[j] = 0u;
#pragma unroll
          for (int c0 = 0; c0 < HD; c0 += 32) tmem_st32(t_r1 + c0, z);
          tmem_wait_st();
        }
      }
      // raw-input K blocks of the NG1 stage (shared-memory A tile): node rows [hdr | x_i], global rows [hdr]
      auto write_xa = [&]() {
        const uint4 h0 = sb[0], h1 = sb[1];
        if (is_g) {
          *reinterpret_cast<uint4*>(xa_row + 4 * 2048) = h0;
          *reinterpret_cast<uint4*>(xa_row + 5 * 2048) = h1;
        } else {
          *reinterpret_cast<uint4*>(xa_row) = h0;
          *reinterpret_cast<uint4*>(xa_row + 2048) = h1;
          *reinterpret_cast<uint4*>(xa_row + 2 * 2048) = sb[2 + 2 * nix];
          *reinterpret_cast<uint4*>(xa_row + 3 * 2048) = sb[3 + 2 * nix];
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
      };
      // A operand of edge pass jj: [hdr | x_i | x_j] of this row's edge -> R1 (KE1 = 48 halves = 24 columns)
      auto build_edge = [&](int jj) {
        const int nj = min(jj + (jj >= nix ? 1 : 0), N - 1);
        const uint4 h0 = sb[0], h1 = sb[1], xi0 = sb[2 + 2 * nix], xi1 = sb[3 + 2 * nix], xj0 = sb[2 + 2 * nj], xj1 = sb[3 + 2 * nj];
        const uint32_t v[16] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w, xi0.x, xi0.y, xi0.z, xi0.w, xi1.x, xi1.y, xi1.z, xi1.w};
        const uint32_t w[8] = {xj0.x, xj0.y, xj0.z, xj0.w, xj1.x, xj1.y, xj1.z, xj1.w};
        tmem_st16(t_r1, v);
        tmem_st8(t_r1 + 16, w);
      };
      // state after model step hh -> traj[hh + 1]: the warp's trajectories are contiguous in global memory, so the staged rows
      // go out as one coalesced copy (issued while an MMA stage runs, off the dependency chain)
      const int n_traj_w = max(0, min(tpw, Gc - qf * tpw));  // live trajectories of this warp
      const int n_out = n_traj_w * O;
      auto flush_outputs = [&](int hh) {
        if (fused) {
          // planner: only the predicted dims, into the trajectory's own block [H][E][Op] of the compact scratch
          float* og = fc.pred + (size_t)(p0 + qf * tpw) * fc.ps + (size_t)hh * fc.hs + (size_t)e * fc.es;
          const int n_pred = n_traj_w * fc.Op;
#pragma unroll 4
          for (int i = lane; i < n_pred; i += 32) {
            const int tb = (int)__umulhi((unsigned)i, fc.Op_magic), c = i - tb * fc.Op;
            og[(size_t)tb * fc.ps + c] = s_obs[tb * O + c];
          }
          return;
        }
        float* og = ra.traj + (((size_t)(hh + 1) * md.E + e) * ra.n + p0 + qf * tpw) * O;
#pragma unroll 4
        for (int i = lane; i < n_out; i += 32) og[i] = s_obs[i];
      };

      // node output: delta -> de-normalise, residual update (gnn_ensemble.py:302-334, 935-939); row of node ni
      auto node_update = [&](const uint32_t (&v)[16]) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(snrm + b * SNs + kAU + ni * kNA);
        float* og = s_obs + bw * O + A + ni * D;
        const float4* tb = reinterpret_cast<const float4*>(sOut);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          if (4 * j4 < D) {
            const float4 c1 = tb[j4], c0 = tb[4 + j4], mi = tb[8 + j4], is = tb[12 + j4];
            // (padded entries: c1 = c0 = 0 and the old value is taken as 0)
            float o0, o1, o2, o3;
            if constexpr (kStateInSmem) {
              o0 = 4 * j4 < D ? og[4 * j4] : 0.f; o1 = 4 * j4 + 1 < D ? og[4 * j4 + 1] : 0.f;
              o2 = 4 * j4 + 2 < D ? og[4 * j4 + 2] : 0.f; o3 = 4 * j4 + 3 < D ? og[4 * j4 + 3] : 0.f;
            } else {
              o0 = dyn[4 * j4]; o1 = dyn[4 * j4 + 1]; o2 = dyn[4 * j4 + 2]; o3 = dyn[4 * j4 + 3];
            }
            const float n0 = o0 + fmaf(c1.x, __uint_as_float(v[4 * j4]), c0.x);
            const float n1 = o1 + fmaf(c1.y, __uint_as_float(v[4 * j4 + 1]), c0.y);
            const float n2 = o2 + fmaf(c1.z, __uint_as_float(v[4 * j4 + 2]), c0.z);
            const float n3 = o3 + fmaf(c1.w, __uint_as_float(v[4 * j4 + 3]), c0.w);
            if constexpr (!kStateInSmem) { dyn[4 * j4] = n0; dyn[4 * j4 + 1] = n1; dyn[4 * j4 + 2] = n2; dyn[4 * j4 + 3] = n3; }
            if (4 * j4 < D) og[4 * j4] = n0;
            if (4 * j4 + 1 < D) og[4 * j4 + 1] = n1;
            if (4 * j4 + 2 < D) og[4 * j4 + 2] = n2;
            if (4 * j4 + 3 < D) og[4 * j4 + 3] = n3;
            // padded entries have c1 = c0 = mi = is = 0 -> exact zeros, like the zero padding of the node block
            const uint32_t w01 = pack_half2((n0 - mi.x) * is.x, (n1 - mi.y) * is.y), w23 = pack_half2((n2 - mi.z) * is.z, (n3 - mi.w) * is.w);
            satm = sat_track(sat_track(satm, w01), w23);
            if (4 * j4 + 1 < D || S == 0) dst[2 * j4] = w01;
            else reinterpret_cast<unsigned short*>(dst)[4 * j4] = (unsigned short)(w01 & 0xffffu);
            if (4 * j4 + 3 < D || (S == 0 && 4 * j4 + 2 < D)) dst[2 * j4 + 1] = w23;
            else if (4 * j4 + 2 < D) reinterpret_cast<unsigned short*>(dst)[4 * j4 + 2] = (unsigned short)(w23 & 0xffffu);
          }
        }
      
      };
      // global output: agent delta, next action; row carrying the agent features
      auto agent_update = [&](const uint32_t (&v)[16], int h) {
        __half* dsth = snrm + b * SNs;
        uint32_t* dst = reinterpret_cast<uint32_t*>(dsth);
        float* og = s_obs + bw * O;
        const float4* tb = reinterpret_cast<const float4*>(sOut + 64);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          if (4 * j4 < A) {
            const float4 c1 = tb[j4], c0 = tb[4 + j4], mi = tb[8 + j4], is = tb[12 + j4];
            float o0, o1, o2, o3;
            if constexpr (kStateInSmem) {
              o0 = 4 * j4 < A ? og[4 * j4] : 0.f; o1 = 4 * j4 + 1 < A ? og[4 * j4 + 1] : 0.f;
              o2 = 4 * j4 + 2 < A ? og[4 * j4 + 2] : 0.f; o3 = 4 * j4 + 3 < A ? og[4 * j4 + 3] : 0.f;
            } else {
              o0 = agent[4 * j4]; o1 = agent[4 * j4 + 1]; o2 = agent[4 * j4 + 2]; o3 = agent[4 * j4 + 3];
            }
            const float n0 = o0 + fmaf(c1.x, __uint_as_float(v[4 * j4]), c0.x);
            const float n1 = o1 + fmaf(c1.y, __uint_as_float(v[4 * j4 + 1]), c0.y);
            const float n2 = o2 + fmaf(c1.z, __uint_as_float(v[4 * j4 + 2]), c0.z);
            const float n3 = o3 + fmaf(c1.w, __uint_as_float(v[4 * j4 + 3]), c0.w);
            if constexpr (!kStateInSmem) { agent[4 * j4] = n0; agent[4 * j4 + 1] = n1; agent[4 * j4 + 2] = n2; agent[4 * j4 + 3] = n3; }
            if (4 * j4 < A) og[4 * j4] = n0;
            if (4 * j4 + 1 < A) og[4 * j4 + 1] = n1;
            if (4 * j4 + 2 < A) og[4 * j4 + 2] = n2;
            if (4 * j4 + 3 < A) og[4 * j4 + 3] = n3;
            // the agent part is followed by the action in the header: only whole pairs may be written as 32 bits
            const uint32_t w01 = pack_half2((n0 - mi.x) * is.x, (n1 - mi.y) * is.y), w23 = pack_half2((n2 - mi.z) * is.z, (n3 - mi.w) * is.w);
            satm = sat_track(sat_track(satm, w01), w23);
            if (4 * j4 + 1 < A) dst[2 * j4] = w01;
            else reinterpret_cast<unsigned short*>(dsth)[4 * j4] = (unsigned short)(w01 & 0xffffu);
            if (4 * j4 + 3 < A) dst[2 * j4 + 1] = w23;
            else if (4 * j4 + 2 < A) reinterpret_cast<unsigned short*>(dsth)[4 * j4 + 2] = (unsigned short)(w23 & 0xffffu);
          }
        }
        if (h + 1 < ra.H) {
          if constexpr (kStateInSmem) asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < U) {
              const float an = kStateInSmem ? s_actn[j] : act_next[j];
              satm = sat_track1(satm, dsth[A + j] = half_sat((an - nrm[md.nrm.in_action + j]) * nrm[md.nrm.in_action + U + j]));
            }
        }
      
      };

      // Global-row tiling: every row has exactly one role, so the two output updates above run as ONE pass with the role's tables
      // and destinations (node rows: dyn block of node ni; global rows: agent part of the header + next action) instead of two
      // divergent passes through the NG3 epilogue.
      auto row_update = [&](const uint32_t (&v)[16], int h) {
        const int dim = is_g ? A : D;
        __half* dsth = snrm + b * SNs + (is_g ? 0 : kAU + ni * kNA);
        uint32_t* dst = reinterpret_cast<uint32_t*>(dsth);
        float* og = s_obs + bw * O + (is_g ? 0 : A + ni * D);
        const float4* tb = reinterpret_cast<const float4*>(sOut + (is_g ? 64 : 0));
        const bool tight = !is_g && S == 0;  // a node block without static features ends in zero padding: whole words may be written
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          if (4 * j4 < dim) {
            const float4 c1 = tb[j4], c0 = tb[4 + j4], mi = tb[8 + j4], is = tb[12 + j4];
            float o0, o1, o2, o3;
            if constexpr (kStateInSmem) {
              o0 = 4 * j4 < dim ? og[4 * j4] : 0.f; o1 = 4 * j4 + 1 < dim ? og[4 * j4 + 1] : 0.f;
              o2 = 4 * j4 + 2 < dim ? og[4 * j4 + 2] : 0.f; o3 = 4 * j4 + 3 < dim ? og[4 * j4 + 3] : 0.f;
            } else {
              o0 = dyn[4 * j4]; o1 = dyn[4 * j4 + 1]; o2 = dyn[4 * j4 + 2]; o3 = dyn[4 * j4 + 3];
            }
            const float n0 = o0 + fmaf(c1.x, __uint_as_float(v[4 * j4]), c0.x);
            const float n1 = o1 + fmaf(c1.y, __uint_as_float(v[4 * j4 + 1]), c0.y);
            const float n2 = o2 + fmaf(c1.z, __uint_as_float(v[4 * j4 + 2]), c0.z);
            const float n3 = o3 + fmaf(c1.w, __uint_as_float(v[4 * j4 + 3]), c0.w);
            if constexpr (!kStateInSmem) { dyn[4 * j4] = n0; dyn[4 * j4 + 1] = n1; dyn[4 * j4 + 2] = n2; dyn[4 * j4 + 3] = n3; }
            if (4 * j4 < dim) og[4 * j4] = n0;
            if (4 * j4 + 1 < dim) og[4 * j4 + 1] = n1;
            if (4 * j4 + 2 < dim) og[4 * j4 + 2] = n2;
            if (4 * j4 + 3 < dim) og[4 * j4 + 3] = n3;
            const uint32_t w01 = pack_half2((n0 - mi.x) * is.x, (n1 - mi.y) * is.y), w23 = pack_half2((n2 - mi.z) * is.z, (n3 - mi.w) * is.w);
            satm = sat_track(sat_track(satm, w01), w23);
            if (4 * j4 + 1 < dim || tight) dst[2 * j4] = w01;
            else reinterpret_cast<unsigned short*>(dsth)[4 * j4] = (unsigned short)(w01 & 0xffffu);
            if (4 * j4 + 3 < dim || (tight && 4 * j4 + 2 < dim)) dst[2 * j4 + 1] = w23;
            else if (4 * j4 + 2 < dim) reinterpret_cast<unsigned short*>(dsth)[4 * j4 + 2] = (unsigned short)(w23 & 0xffffu);
          }
        }
        if (is_g && h + 1 < ra.H) {
          if constexpr (kStateInSmem) asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < U) {
              const float an = kStateInSmem ? s_actn[j] : act_next[j];
              satm = sat_track1(satm, dsth[A + j] = half_sat((an - nrm[md.nrm.in_action + j]) * nrm[md.nrm.in_action + U + j]));
            }
        }
      };

      for (int h = 0; h < ra.H; ++h) {
        stamp(100);
        if (live_g && h + 1 < ra.H) {  // prefetch the next step's action; consumed in the G3 epilogue
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < U) {
              const float* src = ra.actions + ((size_t)(p0 + b) * ra.H + h + 1) * U + j;
              if constexpr (kStateInSmem) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(s_actn + j)), "l"(src) : "memory");
              else act_next[j] = __ldg(src);
            }
        }
        if (warp_on) {
          if constexpr (GROW) write_xa();
          build_edge(0);
        }
        signal_a(cx, s, lane);
        mma_stage(TC_E1);
        if (!GROW && h > 0) flush_outputs(h - 1);  // (global-row tiling: in the shadow of the longest MMA stage, NG1)

#pragma unroll 1
        for (int st = 0; st < n_stage; ++st) {
          const int kind = kind_of(st);
          const int jj = st >> 1;  // edge pass (kinds E1 / E2)
          stamp(10 + kind);
          if constexpr (GROW) {
            // ============ global-row tiling: E1 / E2 per edge pass, then NG1, NG2, NG3 (kind = GK_*) ============
            // parameters of this row's role: node rows take the node MLP's LayerNorm, global rows the global MLP's
            const int mi = kind == GK_E1 ? TC_E1 : kind == GK_E2 ? TC_E2 : (is_g ? TC_G1 : TC_N1) + (kind - GK_NG1);
            const TcMat& m = tl.mat[mi];
            const float* pg = parp(m.par_gamma);
            const float* pbe = parp(m.par_beta);
            const bool lnm = ln && m.ln;
            // global rows ride through the edge stages on filler data and must hand ZEROS to the node-aggregate weights (R2): where
            // the epilogue is relu(x * rstd + beta') a beta' of -60000 does it for free, otherwise the row is cleared explicitly
            const bool neg_trick = lnm && (FOLD || (RELU && fold));
            if (kind == GK_E2 && is_g && neg_trick) pbe = sNeg;
            wait_bar(cx, &d_ready[s], dph, 10 + kind);
            dph ^= 1u;
            tc_fence_after();
            stamp(20 + kind);
            if (kind != GK_NG3) {
              if (warp_on) {
                uint32_t hreg[HD / 2];
                uint32_t sat_row = 0u;
                epi_row<HD, RELU, FOLD>(t_acc, nullptr, pg, pbe, lnm, fold, act, hreg, sat_row);
                if (live && !(is_g && kind <= GK_E2)) satm = __vmaxu2(satm, sat_row);  // (filler / dead rows do not count)
                stamp(110 + kind);
                if (kind == GK_E1) {
                  store_hidden<HD>(t_r1, hreg);
                } else if (kind == GK_E2) {
                  if (!neg_trick && is_g) {
#pragma unroll
                    for (int j = 0; j < HD / 2; ++j) hreg[j] = 0u;
                  }
                  // node aggregate: running sum over this row's outgoing edges in the row's own TMEM lane (R2); W3 is folded
                  if (jj > 0) {
#pragma unroll
                    for (int c0 = 0; c0 < HD / 2; c0 += 32) {
                      uint32_t cur[32];
                      tmem_ld32(t_r2 + c0, cur);
                      tmem_wait_ld();
#pragma unroll
                      for (int j = 0; j < 32; ++j) {
                        const __half2 sum = __hadd2(*reinterpret_cast<const __half2*>(&hreg[c0 + j]), *reinterpret_cast<const __half2*>(&cur[j]));
                        hreg[c0 + j] = *reinterpret_cast<const uint32_t*>(&sum);
                      }
                    }
                  }
                  store_hidden<HD>(t_r2, hreg);
                  if (jj == N - 2) {
                    // global aggregate: sum of the trajectory's N node aggregates (adjacent lanes, segmented shuffle tree into lane
                    // (b, 0)), handed to the trajectory's global row (lane (b, N)) as ITS aggregate block R1; node rows clear R1
                    for (int off = 1; off < N; off <<= 1) {
                      const bool take = ni + off < N;
#pragma unroll
                      for (int j = 0; j < HD / 2; ++j) {
                        uint32_t o = __shfl_down_sync(FULL, hreg[j], off);
                        o = take ? o : 0u;
                        const __half2 sum = __hadd2(*reinterpret_cast<const __half2*>(&hreg[j]), *reinterpret_cast<const __half2*>(&o));
                        hreg[j] = *reinterpret_cast<const uint32_t*>(&sum);
                      }
                    }
#pragma unroll
                    for (int j = 0; j < HD / 2; ++j) {
                      const uint32_t o = __shfl_up_sync(FULL, hreg[j], N);
                      hreg[j] = is_g ? o : 0u;
                    }
                    store_hidden<HD>(t_r1, hreg);
                  } else {
                    build_edge(jj + 1);
                  }
                } else {
                  // NG1 / NG2: hidden row -> the block of the row's own role (node rows R2, global rows R1), zeros -> the other
                  const uint32_t mg = is_g ? 0xffffffffu : 0u, mn = ~mg;
#pragma unroll
                  for (int c0 = 0; c0 < HD / 2; c0 += 32) {
                    uint32_t t2[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) t2[j] = hreg[c0 + j] & mn;
                    tmem_st32(t_r2 + c0, t2);
                  }
#pragma unroll
                  for (int c0 = 0; c0 < HD / 2; c0 += 32) {
                    uint32_t t1[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) t1[j] = hreg[c0 + j] & mg;
                    tmem_st32(t_r1 + c0, t1);
                  }
                }
              }
              stamp(30 + kind);
              signal_a(cx, s, lane);
              stamp(120 + kind);
              mma_stage(kind_of(st + 1));
              if (kind == GK_E2 && jj == N - 2 && h > 0) flush_outputs(h - 1);  // the NG1 MMAs (3 + 2 HD / 16 K steps) run meanwhile
            } else {
              // NG3: 16 output columns -- node rows: delta of the dynamic features, global rows: agent delta + next action
              if (warp_on) {
                uint32_t v[16];
                tmem_ld16(t_acc, v);
                tmem_wait_ld();
                if (live) row_update(v, h);
              }
            }
            continue;
          }
          // stage parameters are fetched before the wait, so their (indexed constant bank) latency hides behind the MMA
          const TcMat& m = tl.mat[kind];
          const float* pg = parp(m.par_gamma);
          const float* pbe = parp(m.par_beta);
          const bool lnm = ln && m.ln;
          const bool cb = (cb_mask >> kind) & 1u;
          wait_bar(cx, &d_ready[s], dph, 10 + kind);
          dph ^= 1u;
          tc_fence_after();
          stamp(20 + kind);
          if (kind != TC_N3 && kind != TC_G3) {
            // ------------------ hidden layer: LayerNorm + activation -> fp16 ------------------
            if (warp_on) {
              uint32_t hreg[HD / 2];
              // (no bias operand: hidden-layer biases ride in a constant-one K column, or -- E2 with N >= 3 -- in the accumulator)
              uint32_t sat_row = 0u;
              epi_row<HD, RELU, FOLD>(t_acc, nullptr, pg, pbe, lnm, fold, act, hreg, sat_row);
              if (live && (kind < TC_G1 || ni == 0)) satm = __vmaxu2(satm, sat_row);  // (the global tile only lives on the rows (b, 0))
              stamp(110 + kind);
              if (kind != TC_E2) {
                store_hidden<HD>(t_r1, hreg);
                if (kind == TC_E1 && N != 2) {
                  // E2's bias goes into the accumulator this thread has just read out; the E2 MMA accumulates on top of it
                  const float* pb2 = sPar + tl.mat[TC_E2].par_bias;
#pragma unroll
                  for (int c0 = 0; c0 < HD; c0 += 32) {
                    uint32_t bv[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      const float4 b4 = *reinterpret_cast<const float4*>(pb2 + c0 + 4 * j);
                      bv[4 * j] = __float_as_uint(b4.x); bv[4 * j + 1] = __float_as_uint(b4.y);
                      bv[4 * j + 2] = __float_as_uint(b4.z); bv[4 * j + 3] = __float_as_uint(b4.w);
                    }
                    tmem_st32(t_acc + c0, bv);
                  }
                }
                // the next layer's bias rides in a constant-one K column that lives in R2 (free once this layer's MMA is done)
                if (cb) tmem_st4(t_r2, make_uint4(0x00003c00u, 0u, 0u, 0u)), tmem_st4(t_r2 + 4, make_uint4(0u, 0u, 0u, 0u));
              } else {
                // ---------- aggregation of the hidden edge activations (models/utils.py:42-51; W3 folded) ----------
                // node aggregate: running sum over this row's outgoing edges, kept in the row's own TMEM lane (R2)
                if (jj > 0) {
#pragma unroll
                  for (int c0 = 0; c0 < HD / 2; c0 += 32) {
                    uint32_t cur[32];
                    tmem_ld32(t_r2 + c0, cur);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                      const __half2 sum = __hadd2(*reinterpret_cast<const __half2*>(&hreg[c0 + j]), *reinterpret_cast<const __half2*>(&cur[j]));
                      hreg[c0 + j] = *reinterpret_cast<const uint32_t*>(&sum);
                    }
                  }
                }
                store_hidden<HD>(t_r2, hreg);
                if (jj == N - 2) {
                  // global aggregate: sum of the N node aggregates of a trajectory (adjacent lanes) by a segmented shuffle tree;
                  // lane (b, 0) parks it in s_glob until the global tile is built
                  for (int off = 1; off < N; off <<= 1) {
                    const bool take = ni + off < N;
#pragma unroll
                    for (int j = 0; j < HD / 2; ++j) {
                      uint32_t o = __shfl_down_sync(FULL, hreg[j], off);
                      o = take ? o : 0u;
                      const __half2 sum = __hadd2(*reinterpret_cast<const __half2*>(&hreg[j]), *reinterpret_cast<const __half2*>(&o));
                      hreg[j] = *reinterpret_cast<const uint32_t*>(&sum);
                    }
                  }
                  if (live_g) {
#pragma unroll
                    for (int c = 0; c < CH; ++c)
                      *reinterpret_cast<uint4*>(s_glob + (size_t)b * (HD * 2) + ((c ^ (b & 7)) * 16)) =
                          make_uint4(hreg[4 * c], hreg[4 * c + 1], hreg[4 * c + 2], hreg[4 * c + 3]);
                  }
                }
              }
            }
            stamp(30 + kind);
            if (kind == TC_E2 && warp_on) {
              // next A operand: the following edge pass, or the x part of the node tile ([hdr | x_i], KXn = 32 halves)
              if (jj + 1 < N - 1) {
                build_edge(jj + 1);
              } else {
                const uint4 h0 = sb[0], h1 = sb[1], xi0 = sb[2 + 2 * ni], xi1 = sb[3 + 2 * ni];
                const uint32_t v[16] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w, xi0.x, xi0.y, xi0.z, xi0.w, xi1.x, xi1.y, xi1.z, xi1.w};
                tmem_st16(t_r1, v);
              }
            }
            signal_a(cx, s, lane);
            stamp(120 + kind);
            mma_stage(kind_of(st + 1));
          } else if (kind == TC_N3) {
            // ------------------ node output: delta -> de-normalise, residual update (gnn_ensemble.py:302-334, 935-939) ------------------
            if (warp_on) {
              uint32_t v[16];
              tmem_ld16(t_acc, v);
              tmem_wait_ld();
              if (live_n) node_update(v);
              // A operand of the global tile (row of trajectory b == lane (b, 0)): [hdr] -> R1, sum over all edges -> R2
              {
                const uint4 h0 = sb[0], h1 = sb[1];
                const uint32_t v8[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                tmem_st8(t_r1, v8);
                uint32_t hreg[HD / 2];
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                  const uint4 g4 = *reinterpret_cast<const uint4*>(s_glob + (size_t)b * (HD * 2) + ((c ^ (b & 7)) * 16));
                  hreg[4 * c] = g4.x; hreg[4 * c + 1] = g4.y; hreg[4 * c + 2] = g4.z; hreg[4 * c + 3] = g4.w;
                }
                store_hidden<HD>(t_r2, hreg);
              }
            }
            signal_a(cx, s, lane);
            mma_stage(TC_G1);
          } else {
            // ------------------ global output: agent delta, next action ------------------
            if (warp_on) {
              uint32_t v[16];
              tmem_ld16(t_acc, v);
              tmem_wait_ld();
              if (live_g) agent_update(v, h);
            }
          }
        }
        __syncwarp();  // snrm of the next step is complete (all rows of a trajectory live in this warp)
        stamp(60);
      }
      flush_outputs(ra.H - 1);
      if (sat_hit(satm)) atomicOr(err, kTcRangeFlag);  // an fp16 operand left the representable range somewhere in this round
    }
    // ---- fused cost tail (planner launches) ----
    // This member's rollouts of the warp's trajectories (all rounds) are complete in global memory.  For every one of them the
    // warp bumps the trajectory's arrival counter; the warp that brings it to E reduces the trajectory (disagreement + task cost
    // + ensemble mean / std, cost_device.cuh) while the predictions are still in L2.  Placed AFTER the round loop on purpose:
    // nothing of the 255-register step loop is live here (inside the loop it cost 600 bytes of spills and 1.8x the time).
    if (fused && fc.done != nullptr && fc.dbg_mode != 2) {
      __threadfence();
      __syncwarp();
      float* envbuf = tpw * O >= md.E * ra.H ? s_obs : nullptr;  // the warp's observation staging is free now
      // phase 0: arrive for every trajectory of this warp; phase 1: reduce the ones assigned to this member (p % E == e) once all
      // members have arrived (every CTA of this persistent grid is resident, so the wait cannot deadlock)
      for (int phase = 0; phase < 2; ++phase)
        for (int rd = 0; rd < rounds; ++rd) {
          const int2 grp = tc_group_of(tl.pairs, md.E, pair, ra.n, NS, tl.Tmax, tpw, (int)rank, s, rd, tl.even_rounds);
          const int p0 = grp.x, Gc = grp.y;
          const int n_traj_w = max(0, min(tpw, Gc - qf * tpw));
          if (phase == 0) {  // lane tb arrives for trajectory tb (the stores were fenced above)
            if (lane < n_traj_w) costdev::fused_arrive_nowait(fc, p0 + qf * tpw + lane);
            continue;
          }
          constexpr int FB = HD == 128 ? 8 : 4;  // loads in flight per lane and member: what the register budget allows
          for (int tb = 0; tb < n_traj_w; ++tb) {
            const int p = p0 + qf * tpw + tb;
            if (p % md.E != e) continue;
            costdev::fused_wait(fc, p, lane);
            const float* srow = ra.start_obs + (size_t)(p / ra.traj_per_start) * O;
            const float* grow = ra.goal + (size_t)(p / ra.traj_per_goal) * md.G;
            if (md.E <= 5) costdev::fused_reduce<5, FB>(fc, p, lane, srow, grow, envbuf);
            else costdev::fused_reduce<kMaxEnsemble, 4>(fc, p, lane, srow, grow, envbuf);
          }
        }
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) tmem_dealloc<2>(tmem_base, 512);
}


#ifdef CEEUS_DEBUG_TOOLS
// ---- epilogue micro-benchmark: the hidden-layer epilogue alone (no barriers), 1..3 warps per scheduler; optionally one
// extra warp keeps the tensor core busy with back-to-back MMAs into other TMEM columns (with_mma) ----
template <int NWG>
__global__ void __launch_bounds__(128 * NWG + 32, 1) epi_bench_kernel(long long* out, int reps, int with_bias, int with_mma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  __shared__ volatile int stop;
  __shared__ __align__(16) float par[3 * 128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 3 * 128; i += blockDim.x) par[i] = 0.5f + 0.001f * i;
  for (int i = tid; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (tid == 0) stop = 0;
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc<1>(&tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 4 * NWG) {
    const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    uint32_t a[32];
    for (int j = 0; j < 32; ++j) a[j] = __float_as_uint(0.01f * (tid + j));
    for (int c0 = 0; c0 < 128; c0 += 32) tmem_st32(base + c0, a);
    tmem_wait_st();
    asm volatile("bar.sync 1, %0;" ::"r"(128 * NWG) : "memory");
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      uint32_t hreg[64];
      uint32_t satm = 0u;
      epi_row<128, true, false>(base, (with_bias & 1) ? par : nullptr, par + 128, par + 256, true, (with_bias & 2) != 0, CEEUS_ACT_RELU, hreg, satm);
      store_hidden<128>(base, hreg);
      tmem_wait_st();
    }
    const long long t1 = clock64();
    asm volatile("bar.sync 1, %0;" ::"r"(128 * NWG) : "memory");
    if (tid == 0) { out[0] = (t1 - t0) / reps; stop = 1; }
  } else if (with_mma && lane == 0) {
    const uint32_t idesc = make_idesc_f16(128, 128);
    const uint64_t da = make_smem_desc(smem_u32(smem), 2048, 128);
    const uint64_t db = make_smem_desc(smem_u32(smem) + 32768, 2048, 128);
    uint32_t ph = 0;
    long long n = 0;
    while (!stop) {
      for (int i = 0; i < 8; ++i) umma_f16_ss<1>(tmem_slot + 384, da, db, idesc, i > 0 ? 1u : 0u);
      umma_commit_cg1(&bar);
      mbar_wait(&bar, ph, nullptr, 0);
      ph ^= 1u;
      n += 8;
    }
    out[1] = n;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem_slot, 512);
}

#endif  // CEEUS_DEBUG_TOOLS

// ---- weight packing ------------------------------------------------------------------------------------------
__device__ __forceinline__ const LayerDesc& mat_layer(const ModelDesc& md, int m) {
  switch (m) {
    case TC_E1: return md.mlp[0].layer[0];
    case TC_E2: return md.mlp[0].layer[1];
    case TC_N1: return md.mlp[1].layer[0];
    case TC_N2: return md.mlp[1].layer[1];
    case TC_N3: return md.mlp[1].layer[2];
    case TC_G1: return md.mlp[2].layer[0];
    case TC_G2: return md.mlp[2].layer[1];
    default: return md.mlp[2].layer[2];
  }
}

// LayerNorm layer whose (hidden) output feeds the K rows of stage matrix m (-1: the rows are raw inputs)
__device__ __forceinline__ int producer_of(int m) {
  switch (m) {
    case TC_E2: return TC_E1;
    case TC_N1: case TC_G1: return TC_E2;  // folded aggregate rows
    case TC_N2: return TC_N1;
    case TC_N3: return TC_N2;
    case TC_G2: return TC_G1;
    case TC_G3: return TC_G2;
    default: return -1;
  }
}

// One block per member: can the LayerNorm affine be folded into the weights?  Needs ReLU, LayerNorm, and for every hidden
// feature gamma != 0 and a moderate beta / |gamma| (the folded activation is scaled by 1 / |gamma| before the fp16 rounding).
__global__ void tc_fold_flag_kernel(const float* __restrict__ w32, const ModelDesc md, const TcLayout tl, float* __restrict__ wpar) {
  const int e = blockIdx.x;
  const float* wm = w32 + (size_t)e * md.member_stride;
  __shared__ int bad;
  if (threadIdx.x == 0) bad = (md.act != CEEUS_ACT_RELU || !md.layer_norm) ? 1 : 0;
  __syncthreads();
  const int lnl[6] = {TC_E1, TC_E2, TC_N1, TC_N2, TC_G1, TC_G2};
  for (int idx = threadIdx.x; idx < 6 * md.Hd; idx += blockDim.x) {
    const LayerDesc& ld = mat_layer(md, lnl[idx / md.Hd]);
    const float g = fabsf(wm[ld.g_off + idx % md.Hd]), b = fabsf(wm[ld.be_off + idx % md.Hd]);
    if (md.layer_norm && (!(g > 1e-4f) || !(b <= 1024.f * g) || !(g < 1e4f))) atomicExch(&bad, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) wpar[(size_t)e * tl.par_floats + tl.par_fold] = bad ? 0.f : 1.f;
}

// scratch[e][m][k][n] (fp32): the logical B matrix of every stage in tile-K order, with the edge output layer folded into
// the aggregate rows of N1 / G1, mean-aggregation scales applied and first-layer biases placed in the constant-one row.
__global__ void tc_logical_kernel(const float* __restrict__ w32, const ModelDesc md, const TcLayout tl, const int* __restrict__ kmap,
                                  const float* __restrict__ wpar, float* __restrict__ scratch) {
  const int e = blockIdx.y;
  const float* wm = w32 + (size_t)e * md.member_stride;
  const bool fold = wpar[(size_t)e * tl.par_floats + tl.par_fold] != 0.f;
  float* out = scratch + (size_t)e * tl.scratch_floats;
  const LayerDesc& e3 = md.mlp[0].layer[2];
  const int NE = md.N * (md.N - 1);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < tl.scratch_floats; idx += gridDim.x * blockDim.x) {
    int m = 0, base = 0;
    for (int t = 0; t < TC_NMAT; ++t) {
      if (idx >= base + tl.mat[t].K * tl.mat[t].N) base += tl.mat[t].K * tl.mat[t].N;
      else { m = t; break; }
    }
    const TcMat& M = tl.mat[m];
    const int k = (idx - base) / M.N, n = (idx - base) % M.N;
    const LayerDesc& ld = mat_layer(md, m);
    const int code = kmap[M.kmap_off + k];
    const bool node = m == TC_N1;
    const int agg_off = node ? (md.D + md.S + md.A + md.U) : (md.A + md.U);  // first aggregate row of the layer input
    const float cnt = node ? (float)(md.N - 1) : (float)NE;
    float v = 0.f;
    if (n < ld.N) {
      if (code >= kFold) {
        const int a = code - kFold;
        float acc = 0.f;
        for (int c = 0; c < md.Hd; ++c) acc = fmaf(wm[e3.w_off + (size_t)a * e3.Npad + c], wm[ld.w_off + (size_t)(agg_off + c) * ld.Npad + n], acc);
        v = md.aggr_mean ? acc / cnt : acc;
      } else if (code == -2) {
        v = wm[ld.b_off + n];
        if (m == TC_N1 || m == TC_G1) {
          float acc = 0.f;
          for (int c = 0; c < md.Hd; ++c) acc = fmaf(wm[e3.b_off + c], wm[ld.w_off + (size_t)(agg_off + c) * ld.Npad + n], acc);
          v += md.aggr_mean ? acc : acc * cnt;
        }
      } else if (code >= 0) {
        v = wm[ld.w_off + (size_t)code * ld.Npad + n];
      }
      // folded LayerNorm: |gamma| of the producing layer scales the consumer's K row
      const int pm = producer_of(m);
      if (fold && pm >= 0 && (code >= kFold || (code >= 0 && m != TC_N1 && m != TC_G1)))
        v *= fabsf(wm[mat_layer(md, pm).g_off + (code >= kFold ? code - kFold : code)]);
    }
    out[idx] = v;
  }
}

// logical fp32 matrices -> centred (LayerNorm layers) fp16 B-operand images + fp32 parameter blocks
__global__ void tc_image_kernel(const float* __restrict__ w32, const ModelDesc md, const TcLayout tl, const float* __restrict__ scratch,
                                uint8_t* __restrict__ wimg, float* __restrict__ wpar) {
  const int e = blockIdx.y;
  const float* wm = w32 + (size_t)e * md.member_stride;
  const float* L = scratch + (size_t)e * tl.scratch_floats;
  const bool fold = wpar[(size_t)e * tl.par_floats + tl.par_fold] != 0.f;
  int rows_total = 0;
  for (int t = 0; t < TC_NMAT; ++t) rows_total += tl.mat[t].K;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < rows_total; idx += gridDim.x * blockDim.x) {
    int m = 0, rbase = 0, fbase = 0;
    for (int t = 0; t < TC_NMAT; ++t) {
      if (idx >= rbase + tl.mat[t].K) { rbase += tl.mat[t].K; fbase += tl.mat[t].K * tl.mat[t].N; }
      else { m = t; break; }
    }
    const TcMat& M = tl.mat[m];
    const int k = idx - rbase;
    const float* rowp = L + fbase + (size_t)k * M.N;
    const int nreal = mat_layer(md, m).N;
    float mean = 0.f;
    if (M.ln && md.layer_norm) {
      for (int n = 0; n < nreal; ++n) mean += rowp[n];
      mean /= (float)nreal;
    }
    const LayerDesc& ldm = mat_layer(md, m);
    for (int n = 0; n < M.N; ++n) {
      float v = n < nreal ? rowp[n] - mean : 0.f;
      if (fold && M.ln && n < nreal && wm[ldm.g_off + n] < 0.f) v = -v;  // sign(gamma) into the (centred) output column
      const int rk = n / M.nloc, nl = n % M.nloc;
      uint8_t* img = wimg + ((size_t)e * 2 + rk) * tl.image_bytes;
      reinterpret_cast<__half*>(img + M.off)[((size_t)(k / 8) * M.nloc + nl) * 8 + (k % 8)] = __float2half_rn(v);
    }
  }
  // parameter block
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < TC_NMAT * 3 * 128; idx += gridDim.x * blockDim.x) {
    const int m = idx / (3 * 128), w = (idx / 128) % 3, j = idx % 128;
    const TcMat& M = tl.mat[m];
    const LayerDesc& ld = mat_layer(md, m);
    const int off = w == 0 ? M.par_bias : w == 1 ? M.par_gamma : M.par_beta;
    const int len = (m == TC_N3 || m == TC_G3) ? 16 : tl.HD;
    if (off < 0 || j >= len) continue;
    float v = 0.f;
    if (j < ld.N) {
      if (w == 0) {
        v = wm[ld.b_off + j];
        if (M.ln && md.layer_norm) {
          float mean = 0.f;
          for (int n = 0; n < ld.N; ++n) mean += wm[ld.b_off + n];
          v -= mean / (float)ld.N;
          if (fold && wm[ld.g_off + j] < 0.f) v = -v;
        }
      } else {
        v = wm[(w == 1 ? ld.g_off : ld.be_off) + j];
        if (fold && w == 2) v /= fabsf(wm[ld.g_off + j]);  // beta / |gamma|
      }
    }
    wpar[(size_t)e * tl.par_floats + off + j] = v;
  }
}

}  // namespace

#ifdef CEEUS_DEBUG_TOOLS
cudaError_t launch_epi_bench(int nwg, int with_bias, int with_mma, long long* out_dev) {
  const int smem = 65536;
  switch (nwg) {
    case 1:
      cudaFuncSetAttribute(epi_bench_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      epi_bench_kernel<1><<<1, 128 + 32, smem>>>(out_dev, 50, with_bias, with_mma);
      break;
    case 2:
      cudaFuncSetAttribute(epi_bench_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      epi_bench_kernel<2><<<1, 256 + 32, smem>>>(out_dev, 50, with_bias, with_mma);
      break;
    default:
      cudaFuncSetAttribute(epi_bench_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      epi_bench_kernel<3><<<1, 384 + 32, smem>>>(out_dev, 50, with_bias, with_mma);
      break;
  }
  return cudaGetLastError();
}

#endif  // CEEUS_DEBUG_TOOLS

bool make_tc_layout(const ModelDesc& md, int num_sms, TcLayout* out, int* kmap_host, int grow) {
  TcLayout t{};
  if (md.kind != CEEUS_MODEL_GNN || (md.Hd != 64 && md.Hd != 128) || md.L != 2) return false;
  if (md.D > 16 || md.A > 16 || md.U > 8 || md.N < 2) return false;
  const int N = md.N, NA = md.D + md.S, Hd = md.Hd;
  t.HD = Hd;
  t.NS = 512 / (2 * Hd);
  // fixed operand geometry: 16-half header [agent | action | 1], 16-half node blocks -> KE1 = 48, KXn = 32, KXg = 16
  if (md.A + md.U + 1 > 16 || NA > 16 || N > 32) return false;
  t.AUp = 16;
  t.NAp = 16;
  t.SNs = t.AUp + N * t.NAp;
  t.KE1 = t.AUp + 2 * t.NAp;
  t.KXn = t.AUp + t.NAp;
  t.KXg = t.AUp;
  if (t.KE1 > Hd) return false;  // R1 holds the edge input (KE1/2 columns) and the hidden row (Hd/2)
  t.grow = grow ? 1 : 0;
  {
    const char* ov = getenv("CEEUS_TC_EVEN_ROUNDS");  // development / A-B switch: the even split of round 1
    t.even_rounds = (ov && ov[0] == '1') ? 1 : 0;
  }
  t.tpw = 32 / (N + t.grow);  // whole trajectories per warp: every row of a trajectory (+ its global row) lives in one warp
  if (t.tpw < 1) return false;
  t.Tmax = 4 * t.tpw;   // trajectories per group (one 128-row tile per CTA and slot)
  t.tail = N * md.S + md.G;
  auto mat = [&](int m, int kx, int kagg, int n, int lnf) {
    t.mat[m].K = kx + kagg; t.mat[m].kx = kx; t.mat[m].N = n; t.mat[m].nloc = n / 2; t.mat[m].ln = lnf;
  };
  mat(TC_E1, t.KE1, 0, Hd, 1);
  mat(TC_E2, Hd, (N == 2 || grow) ? 16 : 0, Hd, 1);  // + constant-one K block (bias): in R2 where R2 is free during the MMA, or (grow) a shared-memory A tile
  mat(TC_N1, t.KXn, Hd, Hd, 1);
  mat(TC_N2, Hd, 16, Hd, 1);
  mat(TC_N3, Hd, 0, 16, 0);
  mat(TC_G1, t.KXg, Hd, Hd, 1);
  mat(TC_G2, Hd, 16, Hd, 1);
  mat(TC_G3, Hd, 0, 16, 0);
  int off = 0, po = 0, ko = 0, sf = 0;
  for (int m = 0; m < TC_NMAT; ++m) {
    TcMat& M = t.mat[m];
    M.off = off;
    off += (M.K / 8) * M.nloc * 16;
    M.kmap_off = ko;
    ko += M.K;
    sf += M.K * M.N;
    const bool first = m == TC_E1 || m == TC_N1 || m == TC_G1;
    const bool outl = m == TC_N3 || m == TC_G3;
    M.par_bias = -1; M.par_gamma = -1; M.par_beta = -1;
    if (!first && M.K == M.kx) { M.par_bias = po; po += outl ? 16 : Hd; }  // otherwise the bias is a K row
    if (!outl) { M.par_gamma = po; po += Hd; M.par_beta = po; po += Hd; }
  }
  t.image_bytes = off;
  t.par_fold = po; po += 4;
  t.par_floats = po;
  t.kmap_len = ko;
  t.scratch_floats = sf;
  if (kmap_host) {
    for (int i = 0; i < ko; ++i) kmap_host[i] = -1;
    const int A = md.A, U = md.U;
    // header [agent | action | 1]: the reference orders the layer inputs edge [x_i, x_j, agent, action] (gnn_modules.py:123-127),
    // node [x_i, agent, action, agg] (:213-223), global [agent, action, agg] (:250-254)
    auto hdr = [&](int m, int src_agent) {
      int* km = kmap_host + t.mat[m].kmap_off;
      for (int k = 0; k < A + U; ++k) km[k] = src_agent + k;
      km[A + U] = -2;
    };
    hdr(TC_E1, 2 * NA);
    for (int f = 0; f < NA; ++f) {
      kmap_host[t.mat[TC_E1].kmap_off + t.AUp + f] = f;
      kmap_host[t.mat[TC_E1].kmap_off + t.AUp + t.NAp + f] = NA + f;
    }
    hdr(TC_N1, NA);
    for (int f = 0; f < NA; ++f) kmap_host[t.mat[TC_N1].kmap_off + t.AUp + f] = f;
    for (int c = 0; c < Hd; ++c) kmap_host[t.mat[TC_N1].kmap_off + t.KXn + c] = kFold + c;
    hdr(TC_G1, 0);
    for (int c = 0; c < Hd; ++c) kmap_host[t.mat[TC_G1].kmap_off + t.KXg + c] = kFold + c;
    for (int m : {TC_E2, TC_N2, TC_N3, TC_G2, TC_G3}) {
      for (int k = 0; k < Hd; ++k) kmap_host[t.mat[m].kmap_off + k] = k;
      if (t.mat[m].K > Hd) kmap_host[t.mat[m].kmap_off + Hd] = -2;  // bias row behind the constant-one column
    }
  }
  // shared memory carve-up
  {
    int sm = 0;
    t.sm_w = sm; sm += round_up(t.image_bytes, 1024);
    t.sm_par = sm; sm += round_up(t.par_floats * 4, 16);
    t.nrm_floats = md.nrm.out_dyn + 2 * md.D;
    t.sm_nrm = sm; sm += round_up(t.nrm_floats * 4, 16);
    t.sm_out = sm; sm += 2 * 4 * 16 * 4;
    int so = 0;
    t.obs_bytes = round_up(t.tpw * md.O * 4, 16);  // per warp: staged next observations of its trajectories
    t.so_obs = so; so += 4 * t.obs_bytes;
    t.so_glob = so; so += grow ? 0 : round_up(t.Tmax, 8) * Hd * 2;  // (grow: the global aggregate goes lane to lane into the global row's TMEM block)
    if (grow) so = round_up(so, 128);
    t.so_xa = so; so += grow ? 6 * 2048 : 0;
    t.so_snrm = so; so += round_up(t.Tmax * t.SNs * 2, 16);
    t.so_tail = so; so += round_up(t.Tmax * t.tail * 4, 16);
    t.so_actn = so; so += 128 * 8 * 4;  // per row: the next step's raw action (cp.async prefetch)
    so = round_up(so, 128);
    for (int i = 0; i < t.NS; ++i) { t.sm_slot[i] = sm; sm += so; }
    t.sm_bar = sm; sm += 128;
    t.sm_one = sm; sm += grow ? 4 * 2048 : 0;
    t.sm_neg = sm; sm += grow ? Hd * 4 : 0;
    t.smem_bytes = sm;
    if (t.smem_bytes > 227 * 1024) return false;
  }
  t.pairs = num_sms / 2;
  if (t.pairs < md.E) return false;
  *out = t;
  return true;
}

// Row layout for launches over n trajectories.  Measured on B200: a stage of the global-row tiling costs about
// 1.3x a node-major stage (two K-stacked layers per MMA stage, role-masked operand stores), and a busy issuing warp costs either
// layout ~30 % in the rounds where it carries trajectories; the global-row tiling (2(N-1) + 3 instead of 2(N-1) + 6 stages per
// step) is taken where rounds x stages wins after both factors.
int tc_pick_grow(const ModelDesc& md, int num_sms, int n) {
  if (md.N + 1 > 32) return 0;
  TcLayout t0{}, t1{};
  if (!make_tc_layout(md, num_sms, &t1, nullptr, 1)) return 0;
  if (!make_tc_layout(md, num_sms, &t0, nullptr, 0)) return 1;
  const int N = md.N;
  // slowest member: rounds weighted 1.3 where the issuing warp carries trajectories (all rounds of an even split with more than
  // three warps' worth per round; the trailing rounds of the mixed split)
  auto worst = [&](const TcLayout& t) {
    double w = 0.0;
    for (int pair = 0; pair < t.pairs; ++pair) {
      const TcTile tt = tc_tile(t.pairs, md.E, pair, n, t.NS, t.Tmax, t.tpw);
      const int busy = tt.uneven == 2 ? tt.rounds - tt.k_pure : (tt.T > 3 * t.tpw && tt.uneven == 0) ? tt.rounds : 0;
      w = std::max(w, (tt.rounds - busy) + 1.3 * busy);
    }
    return w;
  };
  const double c0 = worst(t0) * (2 * (N - 1) + 6);
  const double c1 = worst(t1) * (2 * (N - 1) + 3) * 1.3;
  return c1 < 0.95 * c0 ? 1 : 0;
}

cudaError_t launch_tc_pack(const float* w32, const ModelDesc& md, const TcLayout& tl, const int* kmap_dev, float* scratch,
                           uint8_t* wimg, float* wpar, cudaStream_t st) {
  dim3 grid(128, md.E);
  tc_fold_flag_kernel<<<md.E, 256, 0, st>>>(w32, md, tl, wpar);
  tc_logical_kernel<<<grid, 256, 0, st>>>(w32, md, tl, kmap_dev, wpar, scratch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  tc_image_kernel<<<dim3(16, md.E), 256, 0, st>>>(w32, md, tl, scratch, wimg, wpar);
  return cudaGetLastError();
}

template <int HD, bool RELU>
static cudaError_t launch_tc_n(const ModelDesc& md, const RolloutArgs& ra, const TcLayout& tl, const uint8_t* wimg,
                               const float* wpar, int* err_flag, long long* dbg, bool all_fold, const costdev::FusedCost& fc,
                               cudaStream_t st) {
  const bool fold = RELU && all_fold;
  auto kern = tl.grow ? (fold ? rollout_gnn_tc_kernel<HD, RELU, false, RELU, true> : rollout_gnn_tc_kernel<HD, RELU, false, false, true>)
                      : (fold ? rollout_gnn_tc_kernel<HD, RELU, false, RELU, false> : rollout_gnn_tc_kernel<HD, RELU, false, false, false>);
#ifdef CEEUS_DEBUG_TOOLS  // the timeline-probe instantiations exist only in the debug build
  if (dbg != nullptr)
    kern = tl.grow ? (fold ? rollout_gnn_tc_kernel<HD, RELU, true, RELU, true> : rollout_gnn_tc_kernel<HD, RELU, true, false, true>)
                   : (fold ? rollout_gnn_tc_kernel<HD, RELU, true, RELU, false> : rollout_gnn_tc_kernel<HD, RELU, true, false, false>);
#endif
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tl.smem_bytes);
  if (e != cudaSuccess) return e;
  kern<<<dim3(2 * tl.pairs), 128 * TcSlots<HD>::value, tl.smem_bytes, st>>>(md, ra, tl, wimg, wpar, err_flag, dbg, fc);
  return cudaGetLastError();
}

cudaError_t launch_rollout_gnn_tc(const ModelDesc& md, const RolloutArgs& ra, const TcLayout& tl, const uint8_t* wimg,
                                  const float* wpar, int* err_flag, long long* dbg, bool all_fold, const costdev::FusedCost& fc,
                                  cudaStream_t st) {
  const bool relu = md.act == CEEUS_ACT_RELU;
#ifdef CEEUS_TC_FAST_COMPILE  // development aid (register / spill checks): only the ReLU instantiations
  if (md.Hd == 128) return launch_tc_n<128, true>(md, ra, tl, wimg, wpar, err_flag, dbg, all_fold, fc, st);
  return launch_tc_n<64, true>(md, ra, tl, wimg, wpar, err_flag, dbg, all_fold, fc, st);
#endif
  if (md.Hd == 128) return relu ? launch_tc_n<128, true>(md, ra, tl, wimg, wpar, err_flag, dbg, all_fold, fc, st)
                                : launch_tc_n<128, false>(md, ra, tl, wimg, wpar, err_flag, dbg, all_fold, fc, st);
  return relu ? launch_tc_n<64, true>(md, ra, tl, wimg, wpar, err_flag, dbg, all_fold, fc, st)
              : launch_tc_n<64, false>(md, ra, tl, wimg, wpar, err_flag, dbg, all_fold, fc, st);
}

}  // namespace ceeus
