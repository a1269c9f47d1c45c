// Trajectory cost of ONE trajectory evaluated by ONE warp: trajectory_cost_fn (mbrl/controllers/mpc_torch.py:116-137,
// mbrl/controllers/mpc_torch_curiosity.py:45-80) + the ensemble mean / std (mpc_torch.py:327-332) + env.cost_fn
// (mbrl/environments/fpp_construction_env.py:404-511, playground_env_wgoals.py:108-157).
//
// Shared by the stand-alone cost_kernel (caller-supplied trajectories, ceeus_costs / externally evaluated task costs) and by
// the tail of every rollout kernel: the planner's rollouts never leave the launch -- the warp that completes a trajectory's
// LAST ensemble member (a per-trajectory arrival counter) reduces it on the spot, while the predictions are still in L2.
#pragma once
#include <math_constants.h>

#include "common.cuh"
#include "kernels.cuh"

namespace ceeus {
namespace costdev {

__device__ __forceinline__ float warp_allsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// |a - b| of two 3-vectors behind L2 loads (scalars only: pointer-addressed local arrays would live on the stack)
__device__ __forceinline__ float dist3(const float* a, const float* b) {
  const float x = __ldcg(a) - __ldcg(b), y = __ldcg(a + 1) - __ldcg(b + 1), z = __ldcg(a + 2) - __ldcg(b + 2);
  return sqrtf(x * x + y * y + z * z);
}

// L2 loads: the rows may have been written by other SMs earlier in the same launch (never through this SM's L1)
__device__ __forceinline__ float ld_l2(const float* p) { return __ldcg(p); }

// Where the predictions of one trajectory live.  Full layout (traj [H+1,E,n,O]): row(h, e) = traj + ((h+1) E + e) n O + p O,
// goal = that row + sd.  Compact layout (planner scratch [n][H][E][Op], only the predicted dims): row(h, e) = blk + (h E + e) Op,
// goal = the environment's goal row.
struct TrajView {
  const float* row0;   // predicted row of (h = 0, e = 0)
  long long hs, es;    // strides (floats) between horizon steps / ensemble members
  const float* goal;   // goal row of every (h, e): goal + h ghs + e ges
  long long ghs, ges;
  const float* start;  // start observation row [O] (with its own goal dims): next-block id of the tower tasks
};

struct CostDims { int H, E, O, N, A, D, S, G, sd, Ostd; };

// env.cost_fn on one predicted observation row (`row`: agent at 0, block i at A + D i; `goal`: the goal dims).
__device__ inline float env_cost_row(const float* __restrict__ row, const float* __restrict__ goal, const CostDims& a,
                                     const CostDesc& cd, int kstar) {
  if (cd.env_kind == CEEUS_ENV_PLAYGROUND) {
    float s = 0.f;
    for (int i = 0; i < a.N; ++i)
      for (int c = 0; c < 2; ++c) {
        const float d = fmaxf(fabsf(ld_l2(row + a.A + a.D * i + c) - ld_l2(goal + 2 * i + c)), 0.23f);
        s += d * d;
      }
    return sqrtf(s);
  }
  if (cd.env_kind == CEEUS_ENV_ROBODESK) {
    // Robodesk.cost_fn = -reward (robodesk_env.py:53-135); observation indices are the reference's (agent 24, objects of 13
    // dynamic features in the order drawer, slide, red / green / blue button, ball, upright block, flat block)
    const float gx = ld_l2(row), gy = ld_l2(row + 1), gz = ld_l2(row + 2);
    auto dist = [&](float x, float y, float z) { const float a = gx - x, b = gy - y, c = gz - z; return sqrtf(a * a + b * b + c * c); };
    float rew;
    if (cd.cost_case == CEEUS_RD_OPEN_SLIDE) {               // :75-86
      const float t = __fsub_rn(__fadd_rn(ld_l2(row + 37), 0.3f), 0.55f);
      rew = cd.sparse ? (t >= 0.f ? 1.f : 0.f) : __fsub_rn(t, __fmul_rn(0.05f, dist(ld_l2(row + 37), ld_l2(row + 38), ld_l2(row + 39))));
    } else if (cd.cost_case == CEEUS_RD_OPEN_DRAWER) {       // :88-101
      const float t = __fsub_rn(-0.2f, __fsub_rn(ld_l2(row + 25), 0.85f));
      rew = cd.sparse ? (t >= 0.f ? 1.f : 0.f) : __fsub_rn(t, __fmul_rn(0.05f, dist(ld_l2(row + 24), ld_l2(row + 25), ld_l2(row + 26))));
    } else if (cd.cost_case <= CEEUS_RD_PUSH_BLUE) {         // :103-123 (button z at 52 / 65 / 78, static z 0.76)
      const int zi = 52 + 13 * (cd.cost_case - CEEUS_RD_PUSH_RED);
      const float t = __fsub_rn(-0.0005238f, __fsub_rn(ld_l2(row + zi), 0.76f));
      rew = cd.sparse ? (t >= 0.001f ? 1.f : 0.f) : __fsub_rn(t, __fmul_rn(0.05f, dist(ld_l2(row + zi - 2), ld_l2(row + zi - 1), 0.76f)));
    } else {                                                 // :125-137 (block x at 89 / 102 / 115)
      const int k = cd.cost_case - CEEUS_RD_BALL_OFF_TABLE, bx = 89 + 13 * k;
      const float z0 = k == 0 ? 0.79963282f : k == 1 ? 0.84978449f : 0.77478449f;
      const float z = ld_l2(row + bx + 2);
      rew = cd.sparse ? (z < 0.6f ? 1.f : 0.f)
                      : __fsub_rn(__fsub_rn(1.f, __fdiv_rn(z, z0)), __fmul_rn(0.001f, dist(ld_l2(row + bx), ld_l2(row + bx + 1), z)));
    }
    return -rew;
  }
  if (cd.cost_case == CEEUS_CASE_FLIP || cd.cost_case == CEEUS_CASE_SLIDE) {
    // per-coordinate distances (:297-355): achieved goal = euler angles (Flip, dims 3..5 of a block) or positions (Slide)
    const int ao = cd.cost_case == CEEUS_CASE_FLIP ? 3 : 0;
    float n_sparse = 0.f, dense_exp = 0.f;
    for (int i = 0; i < a.N; ++i) {
      bool any = false;
      float es = 0.f;
      for (int c = 0; c < 3; ++c) {
        const float v = fmaxf(fabsf(ld_l2(row + a.A + a.D * i + ao + c) - ld_l2(goal + 3 * i + c)), cd.buffer_xyz[c]) * cd.coord_weights[c];
        any = any || v > cd.cost_thres[c];
        es += 1.f - expf(-0.5f * v);
      }
      n_sparse += any ? 1.f : 0.f;
      dense_exp += es * (1.f / 3.f);
    }
    float g = 0.f;
    if (cd.shaped && cd.cost_case == CEEUS_CASE_FLIP) {  // :457-458
      const float x = ld_l2(row) - cd.gripper_init[0], y = ld_l2(row + 1) - cd.gripper_init[1], z = ld_l2(row + 2) - cd.gripper_init[2];
      g = sqrtf(x * x + y * y + z * z);
    }
    if (cd.sparse) return n_sparse + g * 0.001f;
    return dense_exp * 1e-3f + n_sparse + g * 0.01f;
  }
  int unsolved = 0;
  float dense = 0.f;
  for (int i = 0; i < a.N; ++i) {
    const float d = dist3(row + a.A + a.D * i, goal + 3 * i);
    unsolved += d > cd.threshold ? 1 : 0;
    dense += fmaxf(d, cd.buffer_threshold);
  }
  float g = 0.f;
  if (cd.shaped) {
    g = dist3(row, kstar < a.N ? row + a.A + a.D * kstar : goal + 3 * a.N);  // gripper to block k* / to the gripper goal
    if (!(a.N - unsolved <= kstar)) g = 0.f;
  }
  if (cd.sparse) return (float)unsolved + (g > cd.threshold ? 0.5f : 0.f);
  if (cd.cost_case == CEEUS_CASE_TOWER) return (float)unsolved + g * 0.01f;
  return dense + g * 0.01f;
}

// One warp, one trajectory: lanes stride over the observation dims for the disagreement term, lane e evaluates the task cost of
// member e.  EM: compile-time bound of the member loops.  env_ext: [E,H] task costs evaluated by the caller, or null.
// Lane e (< E) returns costs_per_model[e] in *cpm_lane; every lane returns the trajectory cost.
// envbuf: optional warp-private scratch of >= E * H floats: the task costs of all (member, step) pairs are then evaluated by all 32
// lanes side by side before the horizon loop instead of by E lanes one step at a time.
template <int EM, int FB = 4>
__device__ __forceinline__ float trajectory_cost_warp(const CostDims& a, const CostDesc& cd, const TrajView& v,
                                                      const float* __restrict__ env_ext, int lane, float* cpm_lane,
                                                      float* envbuf = nullptr) {
  const bool need_env = !cd.curiosity || cd.extrinsic;
  int kstar = 0;
  if (need_env && env_ext == nullptr && cd.env_kind == CEEUS_ENV_CONSTRUCTION && cd.shaped && cd.cost_case != CEEUS_CASE_FLIP &&
      cd.cost_case != CEEUS_CASE_SLIDE) {  // get_next_block_id, :365-383 (start obs, with the goal dims it carries)
    int unsolved = 0;
    for (int i = 0; i < a.N; ++i) {
      unsolved += dist3(v.start + a.A + a.D * i, v.start + a.sd + 3 * i) > cd.threshold ? 1 : 0;
    }
    kstar = a.N - unsolved;
  }
  const bool env_pre = need_env && env_ext == nullptr && envbuf != nullptr;
  if (env_pre) {
    __syncwarp();
    for (int i = lane; i < a.E * a.H; i += 32) {
      const int e = i / a.H, h = i - e * a.H;
      envbuf[i] = env_cost_row(v.row0 + (long long)h * v.hs + e * v.es, v.goal + (long long)h * v.ghs + e * v.ges, a, cd, kstar);
    }
    __syncwarp();
  }
  const float inv_em1 = 1.f / (float)(a.E - 1);
  float acc = cd.cost_along == CEEUS_COST_BEST ? CUDART_INF_F : 0.f;
  if (cd.curiosity && !need_env && cd.cost_along == CEEUS_COST_SUM) {
    // ---- pure disagreement, summed along the horizon (curious-i-cem.yaml): cost = - sum over ALL (h, d) of std_e ----
    // No per-step structure is needed, so the lanes stride over the flat (h, d) index, FB elements each per batch with all
    // FB x E loads issued before the first use; one warp reduction at the very end.
    constexpr int FB = 4;
    const int n_el = a.H * a.Ostd;
    float part = 0.f;
    int h = lane / a.Ostd, d = lane - h * a.Ostd;  // element of this lane in the current stride of 32
    for (int i0 = 0; i0 < n_el; i0 += 32 * FB) {
      float x[FB][EM];
      bool on[FB];
#pragma unroll
      for (int k = 0; k < FB; ++k) {
        on[k] = i0 + 32 * k + lane < n_el;
        const float* src = v.row0 + (long long)min(h, a.H - 1) * v.hs + d;
#pragma unroll
        for (int e = 0; e < EM; ++e) x[k][e] = ld_l2(src + (long long)min(e, a.E - 1) * v.es);
        d += 32;
        while (d >= a.Ostd) { d -= a.Ostd; ++h; }
      }
#pragma unroll
      for (int k = 0; k < FB; ++k) {
        float mean = 0.f, m2 = 0.f;  // Welford, as torch.std does
#pragma unroll
        for (int e = 0; e < EM; ++e)
          if (e < a.E) {
            const float dl = x[k][e] - mean;
            mean = fmaf(dl, 1.f / (float)(e + 1), mean);
            m2 = fmaf(dl, x[k][e] - mean, m2);
          }
        if (on[k]) part += sqrtf(m2 * inv_em1);
      }
    }
    acc = -warp_allsum(part);
  } else {
  // HB horizon steps at a time: all HB x E loads of a 32-feature column block are issued before any is used, so a warp keeps
  // ~40 sectors in flight instead of one dependent load chain
  constexpr int HB = 4;
  for (int h0 = 0; h0 < a.H; h0 += HB) {
    float bonus[HB];
#pragma unroll
    for (int hb = 0; hb < HB; ++hb) bonus[hb] = 0.f;
    if (cd.curiosity) {
      // two 32-feature column blocks per pass (dims >= Ostd are identical across the members: their std is exactly 0).
      // All 2 x HB x E loads are UNCONDITIONAL (indices clamped, results masked afterwards): behind predicates the compiler
      // issued them one dependent load at a time, and a rollout warp has no other warps to hide that latency behind.
      for (int d0 = 0; d0 < a.Ostd; d0 += 64) {
        const int d = d0 + lane;
        const int da = min(d, a.Ostd - 1), db = min(d + 32, a.Ostd - 1);
        float x[2][HB][EM];
#pragma unroll
        for (int hb = 0; hb < HB; ++hb)
#pragma unroll
          for (int e = 0; e < EM; ++e) {
            const float* src = v.row0 + (long long)min(h0 + hb, a.H - 1) * v.hs + (long long)min(e, a.E - 1) * v.es;
            x[0][hb][e] = ld_l2(src + da);
            x[1][hb][e] = ld_l2(src + db);
          }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const bool on = c == 0 ? d < a.Ostd : d + 32 < a.Ostd;
#pragma unroll
          for (int hb = 0; hb < HB; ++hb) {
            // Welford, as torch.std does (exactly 0 for dims identical across members); 1/(e+1) as a compile-time constant
            // multiply: <= 1 ulp from torch's division (IEEE divisions made this instruction-bound)
            float mean = 0.f, m2 = 0.f;
#pragma unroll
            for (int e = 0; e < EM; ++e)
              if (e < a.E) {
                const float dl = x[c][hb][e] - mean;
                mean = fmaf(dl, 1.f / (float)(e + 1), mean);
                m2 = fmaf(dl, x[c][hb][e] - mean, m2);
              }
            if (on && h0 + hb < a.H) bonus[hb] += sqrtf(m2 * inv_em1);
          }
        }
      }
    }
#pragma unroll
    for (int hb = 0; hb < HB; ++hb) {
      const int h = h0 + hb;
      if (h < a.H) {
        const float bs = cd.curiosity ? warp_allsum(bonus[hb]) : 0.f;
        float c = 0.f;
        if (lane < a.E) {
          float envc = 0.f;
          if (need_env)
            envc = env_ext != nullptr ? __ldg(env_ext + (size_t)lane * a.H + h)
                   : env_pre          ? envbuf[lane * a.H + h]
                                      : env_cost_row(v.row0 + (long long)h * v.hs + lane * v.es,
                                                     v.goal + (long long)h * v.ghs + lane * v.ges, a, cd, kstar);
          c = cd.curiosity ? (cd.extrinsic ? __fadd_rn(-bs, __fmul_rn(cd.extrinsic_scale, envc)) : -bs) : envc;
        }
        if (cd.cost_along == CEEUS_COST_BEST) {
          if (h >= 1) acc = fminf(acc, c);  // amin over h >= 1 (mpc_torch.py:131)
        } else {
          acc += c;
        }
      }
    }
  }
  }
  *cpm_lane = acc;
  float s = 0.f;
  for (int e = 0; e < a.E; ++e) s += __shfl_sync(0xffffffffu, acc, e);
  const float mean = s / (float)a.E;
  float out = mean;
  if (cd.use_std) {
    float vv = 0.f;
    for (int e = 0; e < a.E; ++e) {
      const float d = __shfl_sync(0xffffffffu, acc, e) - mean;
      vv = fmaf(d, d, vv);
    }
    out = mean + sqrtf(vv / (float)(a.E - 1));
  }
  return out;
}

// ---- fused tail of the rollout kernels -----------------------------------------------------------------------------------
// Planner scratch `pred`: only the dims the model predicts, element (p, h, e, d) at pred + p ps + h hs + e es + d.
//   trajectory-major (GNN kernels): ps = pstride (block [H][E][Op] padded to whole 128-byte lines), hs = E Op, es = Op -- the
//     reduction reads one contiguous block per trajectory and may drop it from L2 afterwards;
//   step-major (dense-MLP kernels, whose threads write their rows straight from registers): ps = Op, hs = E n Op, es = n Op.
struct FusedCost {
  float* pred;             // null: not fused (the launch writes full trajectories instead)
  long long ps, hs, es;    // strides in floats (see above)
  long long pstride;       // trajectory-major only: floats per trajectory block (multiple of 32), 0 otherwise
  int Op;                  // predicted dims per row (GNN: A + N D, dense MLP: sd)
  unsigned Op_magic;       // floor(2^32 / Op) + 1: i / Op == __umulhi(i, Op_magic) for the small i the rollout kernels divide
  int* done;               // [n] arrival counters (self-resetting)
  float* cpm;              // [n,E]
  float* costs;            // [n]
  int discard;             // drop the consumed block from L2 without a write-back (discard.global.L2)
  int dbg_mode;            // development aid (CEEUS_TAIL_DEBUG): 1 = arrive only, 2 = no tail at all
  CostDims dims;
  CostDesc cd;
};

// ---- protocol ---------------------------------------------------------------------------------------------------------------
// fused_arrive: called by one whole warp once ITS member's rollout of trajectory p is complete in global memory (every lane has
//   executed __threadfence() after its stores); returns true in the warp that brought the arrival count to E.
// fused_wait:   spins (with back-off) until all E members of trajectory p have arrived.  Only for kernels whose CTAs are all
//   co-resident (the persistent tensor-core kernels): there the reduction of p is assigned STATICALLY to member p % E, so the work
//   is spread over every warp of the grid instead of piling up on the slowest member.
// fused_reduce: the reduction itself; resets the counter for the next launch.
__device__ __forceinline__ bool fused_arrive(const FusedCost& f, int p, int lane) {
  int old = 0;
  if (lane == 0) old = atomicAdd(f.done + p, 1);
  return __shfl_sync(0xffffffffu, old, 0) == f.dims.E - 1;
}
// arrival without the answer (statically assigned reductions): a fire-and-forget reduction, no round trip to L2
__device__ __forceinline__ void fused_arrive_nowait(const FusedCost& f, int p) {
  asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(f.done + p) : "memory");
}

__device__ __forceinline__ void fused_wait(const FusedCost& f, int p, int lane) {
  if (lane == 0) {
    const volatile int* c = f.done + p;
    unsigned ns = 64;
    while (*c < f.dims.E) {
      __nanosleep(ns);
      if (ns < 1024) ns <<= 1;
    }
  }
  __syncwarp();
}

template <int EM, int FB = 4>
__device__ __forceinline__ void fused_reduce(const FusedCost& f, int p, int lane, const float* start_row, const float* goal_row,
                                             float* envbuf /* warp-private, >= E * H floats, or null */) {
  __threadfence();  // acquire: the other members' rows
  if (lane == 0) f.done[p] = 0;  // ready for the next launch
  if (f.dbg_mode == 1) return;
  TrajView v;
  v.row0 = f.pred + (long long)p * f.ps;
  v.hs = f.hs; v.es = f.es;
  v.goal = goal_row; v.ghs = 0; v.ges = 0;
  v.start = start_row;
  float cpm = 0.f;
  const float c = trajectory_cost_warp<EM, FB>(f.dims, f.cd, v, nullptr, lane, &cpm, envbuf);
  if (lane < f.dims.E) f.cpm[(size_t)p * f.dims.E + lane] = cpm;
  if (lane == 0) f.costs[p] = c;
  if (f.discard && f.pstride > 0) {
    // the block is dead now: let L2 drop its (dirty) lines instead of writing them back to HBM
    const char* blk = reinterpret_cast<const char*>(v.row0);
    const long long bytes = f.pstride * 4;
    __syncwarp();
    for (long long off = (long long)lane * 128; off < bytes; off += 32 * 128)
      asm volatile("discard.global.L2 [%0], 128;" ::"l"(blk + off) : "memory");
  }
}

// "last arriver reduces" (kernels whose grid may exceed residency: waiting for another CTA could deadlock there)
template <int EM>
__device__ __forceinline__ void fused_cost_arrive(const FusedCost& f, int p, int lane, const float* start_row, const float* goal_row,
                                                  float* envbuf) {
  if (f.dbg_mode == 2) return;
  if (fused_arrive(f, p, lane)) fused_reduce<EM>(f, p, lane, start_row, goal_row, envbuf);
}

}  // namespace costdev
}  // namespace ceeus
