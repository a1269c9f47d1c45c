// iCEM sampling / cost / elite-selection / refit kernels (everything of TorchMpcICem.get_action that is not the
// model rollout).  All of them are tiny, HBM/latency-bound kernels: coalesced row accesses, warp-level reductions.
#include <math_constants.h>

#include "common.cuh"
#include "cost_device.cuh"
#include "kernels.cuh"

namespace ceeus {

namespace {

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC 2011); counters checked bit for bit in tests/test_gpu_sampler_props.py.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ float u01(unsigned x) { return (float)(((double)x + 0.5) * 2.3283064365386963e-10); }

// 4 standard normals for (pair, block): counter = (block, pair_lo, pair_hi, stream), key = seed
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned stream, unsigned long long pair,
                                                 unsigned block) {
  const uint4 x = philox4x32_10(make_uint4(block, (unsigned)pair, (unsigned)(pair >> 32), stream),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const float u0 = u01(x.x), u1 = u01(x.y), u2 = u01(x.z), u3 = u01(x.w);
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  const float two_pi = 6.283185307179586f;
  float4 z;
  z.x = r0 * cosf(two_pi * u1);
  z.y = r0 * sinf(two_pi * u1);
  z.z = r1 * cosf(two_pi * u3);
  z.w = r1 * sinf(two_pi * u3);
  return z;
}

// clip(y * std + mean) with the reference's op order: mul, add (two roundings), min(high), max(low)  mpc_torch.py:227-230
__device__ __forceinline__ float scale_clip(float y, float sd, float mu, float lo, float hi) {
  return fmaxf(fminf(__fadd_rn(__fmul_rn(y, sd), mu), hi), lo);
}

constexpr int kPairsPerBlock = 32;
constexpr int kSampleThreads = 256;

// sample_action_sequences (mpc_torch.py:204-236) + torch_powerlaw_psd_gaussian (colored_noise.py:115-218) in its
// closed form y = [zr | zi] @ M.  One block = 32 (trajectory, action-dim) pairs.
__global__ void __launch_bounds__(kSampleThreads) sample_kernel(const SampleArgs a) {
  extern __shared__ float sm[];
  const int nz = a.colored ? 2 * a.F : a.H;      // normals per pair
  const int nzp = (nz + 3) & ~3;
  float* zbuf = sm;                               // [32][nzp]
  float* M = zbuf + kPairsPerBlock * nzp;         // [2F][H]
  const int tid = threadIdx.x;
  const long long pair0 = (long long)blockIdx.x * kPairsPerBlock;
  const long long npairs = (long long)a.n * a.U;
  const unsigned sid = a.stream_id + (a.step_ctr != nullptr ? (unsigned)(*a.step_ctr) * a.streams_per_step : 0u);

  if (a.noise == nullptr) {
    if (a.colored)
      for (int i = tid; i < 2 * a.F * a.H; i += kSampleThreads) M[i] = __ldg(a.irdft + i);
    if (a.gauss != nullptr && !a.write_gauss) {
      for (int i = tid; i < kPairsPerBlock * nz; i += kSampleThreads) {
        const long long pr = pair0 + i / nz;
        zbuf[(i / nz) * nzp + i % nz] = pr < npairs ? __ldg(a.gauss + pr * nz + i % nz) : 0.f;
      }
    } else {
      const int nb = nzp / 4;
      for (int i = tid; i < kPairsPerBlock * nb; i += kSampleThreads) {
        const int lp = i / nb, b = i % nb;
        const long long pr = pair0 + lp;
        if (pr < npairs) {
          const long long r = pr / a.U;
          const int u = (int)(pr % a.U);
          const long long grow = (r / a.rows_per_env) * a.global_rows_per_env + a.first_row + r % a.rows_per_env;
          const float4 z = philox_normal4(a.seed, sid, (unsigned long long)(grow * a.U + u), (unsigned)b);
          *reinterpret_cast<float4*>(zbuf + lp * nzp + b * 4) = z;
        }
      }
    }
    __syncthreads();
    if (a.gauss != nullptr && a.write_gauss)
      for (int i = tid; i < kPairsPerBlock * nz; i += kSampleThreads) {
        const long long pr = pair0 + i / nz;
        if (pr < npairs) a.gauss[pr * nz + i % nz] = zbuf[(i / nz) * nzp + i % nz];
      }
  }
  for (int i = tid; i < kPairsPerBlock * a.H; i += kSampleThreads) {
    const int lp = i / a.H, t = i % a.H;
    const long long pr = pair0 + lp;
    if (pr >= npairs) continue;
    const long long r = pr / a.U;
    const int u = (int)(pr % a.U);
    float y;
    if (a.noise != nullptr) {
      y = __ldg(a.noise + pr * a.H + t);
    } else if (a.colored) {
      y = 0.f;
      const float* z = zbuf + lp * nzp;
      for (int k = 0; k < 2 * a.F; ++k) y = fmaf(z[k], M[k * a.H + t], y);
    } else {
      y = zbuf[lp * nzp + t];
    }
    const int env = (int)(r / a.rows_per_env);
    const int mi = (env * a.H + t) * a.U + u;
    a.actions[(r * a.H + t) * a.U + u] = scale_clip(y, __ldg(a.std + mi), __ldg(a.mean + mi), __ldg(a.low + u), __ldg(a.high + u));
  }
}

// mean-action row (mpc_torch.py:242-243) and elites shifted over time (:249-267, :304-310).  One block per env.
__global__ void fixup_kernel(const FixupArgs a) {
  const int env = blockIdx.x, tid = threadIdx.x;
  const int HU = a.H * a.U;
  float* act = a.actions + (size_t)env * a.P * HU;
  const float* mean = a.mean + (size_t)env * HU;
  const float* std = a.std + (size_t)env * HU;
  if (a.write_mean_row)
    for (int i = tid; i < HU; i += blockDim.x) act[i] = mean[i];
  if (a.shift_elites) {
    const unsigned sid = a.stream_id + (a.step_ctr != nullptr ? (unsigned)(*a.step_ctr) * a.streams_per_step : 0u);
    __syncthreads();
    const float* el = a.elite_actions + (size_t)env * a.k * HU;
    for (int i = tid; i < a.n_kept * HU; i += blockDim.x) {
      const int j = i / HU, q = i % HU, t = q / a.U, u = q % a.U;
      float v;
      if (t < a.H - 1) {
        v = el[j * HU + (t + 1) * a.U + u];
      } else {
        const long long pr = ((long long)env * a.n_kept + j) * a.U + u;
        float y;
        if (a.shift_noise != nullptr) {
          y = __ldg(a.shift_noise + pr * a.H + t);
        } else if (a.colored) {
          y = 0.f;
          const int nb = (2 * a.F + 3) / 4;
          for (int b = 0; b < nb; ++b) {
            const float4 z = philox_normal4(a.seed, sid, (unsigned long long)pr, (unsigned)b);
            const float zz[4] = {z.x, z.y, z.z, z.w};
            for (int c = 0; c < 4; ++c) {
              const int k = b * 4 + c;
              if (k < 2 * a.F) y = fmaf(zz[c], __ldg(a.irdft + k * a.H + t), y);
            }
          }
        } else {
          const float4 z = philox_normal4(a.seed, sid, (unsigned long long)pr, (unsigned)(t / 4));
          const float zz[4] = {z.x, z.y, z.z, z.w};
          y = zz[t % 4];
        }
        v = scale_clip(y, std[q], mean[q], __ldg(a.low + u), __ldg(a.high + u));
      }
      act[j * HU + q] = v;
    }
  }
}

// trajectory_cost_fn + ensemble mean / std on caller-supplied FULL trajectories [H+1,E,n,O] (ceeus_costs, externally evaluated
// task costs): one warp per trajectory, the arithmetic is cost_device.cuh's (shared with the fused tail of the rollout kernels).
template <int EM>
__global__ void __launch_bounds__(256) cost_kernel(const CostArgs a, const CostDesc cd) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= a.n) return;
  costdev::CostDims dm{a.H, a.E, a.O, a.N, a.A, a.D, a.S, a.G, a.sd, a.Ostd};
  costdev::TrajView v;
  v.hs = (long long)a.E * a.n * a.O;  // stride of one horizon slot
  v.es = (long long)a.n * a.O;
  v.row0 = a.traj + v.hs + (size_t)p * a.O;  // traj[1]: next_observations of step 0
  v.goal = v.row0 + a.sd; v.ghs = v.hs; v.ges = v.es;
  v.start = a.traj + (size_t)p * a.O;
  float cpm = 0.f;
  const float c = costdev::trajectory_cost_warp<EM>(dm, cd, v, a.env_cost_ext ? a.env_cost_ext + (size_t)p * a.E * a.H : nullptr, lane, &cpm);
  if (lane < a.E) a.cpm[(size_t)p * a.E + lane] = cpm;
  if (lane == 0) a.costs[p] = c;
}

// The same reduction on the planner's COMPACT scratch (cost_device.cuh: FusedCost; predicted dims only): one warp per
// trajectory, 64 warps per SM -- the high-occupancy alternative to the in-kernel tail of the rollout kernels.
template <int EM>
__global__ void __launch_bounds__(256) cost_compact_kernel(const costdev::FusedCost f, const float* __restrict__ start_obs,
                                                           const float* __restrict__ goal, int traj_per_start, int traj_per_goal,
                                                           int n) {
  extern __shared__ float envscratch[];  // [warps][E * H] task costs of all (member, step) pairs (0 bytes for pure curiosity)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = blockIdx.x * (blockDim.x >> 5) + warp;
  if (p >= n) return;
  const bool need_env = !f.cd.curiosity || f.cd.extrinsic;
  costdev::TrajView v;
  v.row0 = f.pred + (long long)p * f.ps;
  v.hs = f.hs; v.es = f.es;
  v.goal = goal + (size_t)(p / traj_per_goal) * f.dims.G; v.ghs = 0; v.ges = 0;
  v.start = start_obs + (size_t)(p / traj_per_start) * f.dims.O;
  float cpm = 0.f;
  const float c = costdev::trajectory_cost_warp<EM, 4>(f.dims, f.cd, v, nullptr, lane, &cpm,
                                                       need_env ? envscratch + warp * f.dims.E * f.dims.H : nullptr);
  if (lane < f.dims.E) f.cpm[(size_t)p * f.dims.E + lane] = cpm;
  if (lane == 0) f.costs[p] = c;
}

// ---------------------------------------------------------------------------------------------------------
// elite selection.  Order = (cost, index) ascending == torch.sort(stable=True) (mpc_torch.py:451-456); NaN -> +inf.
// ---------------------------------------------------------------------------------------------------------
struct Key {
  float c;
  int i;
};
__device__ __forceinline__ bool key_less(const Key& a, const Key& b) { return a.c < b.c || (a.c == b.c && a.i < b.i); }
__device__ __forceinline__ Key key_min(const Key& a, const Key& b) { return key_less(b, a) ? b : a; }

constexpr int kTopkThreads = 1024;

// this rank's k best fresh trajectories -> candidate records [cost, global index, costs_per_model[E], actions[H*U]]
__global__ void __launch_bounds__(kTopkThreads) topk_kernel(const TopkArgs a) {
  __shared__ Key wbest[32];
  __shared__ Key chosen;
  __shared__ Key picked[64];
  const int env = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float* costs = a.costs + (size_t)env * a.P;
  Key last{-CUDART_INF_F, -1};
  float* cand = a.cand + (size_t)env * a.k * a.rec;
  // the costs are read once: up to kTopkLocal per thread stay in registers for all k selection rounds (larger populations
  // fall back to re-reading the tail from L2)
  constexpr int kTopkLocal = 8;
  float cl[kTopkLocal];
#pragma unroll
  for (int j = 0; j < kTopkLocal; ++j) {
    const int i = tid + j * kTopkThreads;
    float c = i < a.P ? __ldg(costs + i) : CUDART_INF_F;
    if (c != c) c = CUDART_INF_F;
    cl[j] = c;
  }
  for (int r = 0; r < a.k; ++r) {
    Key best{CUDART_INF_F, 0x7fffffff};
#pragma unroll
    for (int j = 0; j < kTopkLocal; ++j) {
      const int i = tid + j * kTopkThreads;
      const Key k{cl[j], i};
      if (i < a.P && key_less(last, k)) best = key_min(best, k);
    }
    for (int i = tid + kTopkLocal * kTopkThreads; i < a.P; i += kTopkThreads) {
      float c = __ldg(costs + i);
      if (c != c) c = CUDART_INF_F;
      const Key k{c, i};
      if (key_less(last, k)) best = key_min(best, k);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Key other{__shfl_xor_sync(0xffffffffu, best.c, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
      best = key_min(best, other);
    }
    if (lane == 0) wbest[wid] = best;
    __syncthreads();
    if (wid == 0) {
      Key b = wbest[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Key other{__shfl_xor_sync(0xffffffffu, b.c, o), __shfl_xor_sync(0xffffffffu, b.i, o)};
        b = key_min(b, other);
      }
      if (lane == 0) chosen = b;
    }
    __syncthreads();
    last = chosen;
    if (tid == 0) picked[r] = last;
    __syncthreads();
  }
  // the k records are copied together once the selection is complete (one round trip to L2 instead of k dependent ones)
  const int body = a.E + a.HU;
  for (int t = tid; t < a.k * (2 + body); t += kTopkThreads) {
    const int r = t / (2 + body), c = t % (2 + body);
    const Key pk = picked[r];
    const bool valid = pk.i != 0x7fffffff;  // fewer than k finite candidates cannot happen for P >= k
    const int li = valid ? pk.i : 0;
    float v;
    if (c == 0) v = valid ? pk.c : CUDART_INF_F;
    else if (c == 1) v = __int_as_float((int)(a.first_index + (long long)li));
    else if (c < 2 + a.E) v = __ldg(a.cpm + ((size_t)env * a.P + li) * a.E + (c - 2));
    else v = __ldg(a.actions + ((size_t)env * a.P + li) * a.HU + (c - 2 - a.E));
    cand[(size_t)r * a.rec + c] = v;
  }
}

constexpr int kUpdThreads = 256;
constexpr int kMaxCand = 256;

// update_distributions (mpc_torch.py:445-478) over the union of all ranks' candidates and the kept elites (:319-321).
__global__ void __launch_bounds__(kUpdThreads) update_kernel(const UpdateArgs a, const float one_minus_alpha) {
  extern __shared__ float sm[];
  __shared__ Key keys[kMaxCand];
  __shared__ int sel[64];
  const int env = blockIdx.x, tid = threadIdx.x;
  const int HU = a.H * a.U;
  const int nc = a.world * a.k;
  const int M = nc + a.n_kept_in;
  const int body = a.E + HU;
  float* stage = sm;  // [k][body]
  const float* cand_all = a.cand_all;
  if (a.step_ctr != nullptr) {
    // peer-memory exchange: wait until every rank has delivered the records of THIS exchange (bounded: a rank that never plans
    // must surface as an error flag, not as a hung GPU), then read the half of the double buffer they were written to
    const unsigned long long x = *a.step_ctr * (unsigned long long)a.iters_per_step + (unsigned long long)a.iter;
    if (tid < a.world) {
      const volatile unsigned long long* f = a.flags + tid;
      unsigned long long spins = 0;
      while (*f < x + 1ull) {
        __nanosleep(100);
        if (++spins > (1ull << 24)) { if (a.err) atomicExch(a.err, 21); break; }
      }
    }
    __syncthreads();
    __threadfence_system();
    cand_all += (x & 1ull) * a.half_stride;
  }
  float* el_act = a.elite_actions + (size_t)env * a.k * HU;
  float* el_cost = a.elite_costs + (size_t)env * a.k;
  float* el_cpm = a.elite_cpm + (size_t)env * a.k * a.E;
  int* el_idx = a.elite_idx + (size_t)env * a.k;

  for (int t = tid; t < M; t += kUpdThreads) {
    Key k;
    if (t < nc) {
      const int w = t / a.k, j = t % a.k;
      const float* rec = cand_all + (((size_t)w * a.nE + env) * a.k + j) * a.rec;
      k.c = __ldcg(rec);  // (L2 loads: with the peer-memory exchange the records were stored by other GPUs)
      k.i = __float_as_int(__ldcg(rec + 1));
    } else {
      const int j = t - nc;
      k.c = el_cost[j];
      k.i = (int)(a.pop_fresh + j);
    }
    if (k.c != k.c) k.c = CUDART_INF_F;
    keys[t] = k;
  }
  __syncthreads();
  for (int t = tid; t < M; t += kUpdThreads) {
    const Key me = keys[t];
    int rank = 0;
    for (int j = 0; j < M; ++j) rank += key_less(keys[j], me) ? 1 : 0;
    if (rank < a.k) sel[rank] = t;
  }
  __syncthreads();
  for (int i = tid; i < a.k * body; i += kUpdThreads) {
    const int r = i / body, q = i % body;
    const int t = sel[r];
    float v;
    if (t < nc) {
      const int w = t / a.k, j = t % a.k;
      v = __ldcg(cand_all + (((size_t)w * a.nE + env) * a.k + j) * a.rec + 2 + q);
    } else {
      const int j = t - nc;
      v = q < a.E ? el_cpm[j * a.E + q] : el_act[j * HU + (q - a.E)];
    }
    stage[i] = v;
  }
  __syncthreads();
  for (int i = tid; i < a.k * body; i += kUpdThreads) {
    const int r = i / body, q = i % body;
    if (q < a.E) el_cpm[r * a.E + q] = stage[i];
    else el_act[r * HU + (q - a.E)] = stage[i];
  }
  for (int r = tid; r < a.k; r += kUpdThreads) {
    el_cost[r] = keys[sel[r]].c;
    el_idx[r] = keys[sel[r]].i;
  }
  // refit: mean/std over (elite, ensemble) -> Bessel n = k*E (:464-466); momentum alpha (:471-478)
  float* mean = a.mean + (size_t)env * HU;
  float* std = a.std + (size_t)env * HU;
  for (int q = tid; q < HU; q += kUpdThreads) {
    float s = 0.f;
    for (int r = 0; r < a.k; ++r) s += stage[r * body + a.E + q];
    const float m = s / (float)a.k;
    float v = 0.f;
    for (int r = 0; r < a.k; ++r) {
      const float d = stage[r * body + a.E + q] - m;
      v = fmaf(d, d, v);
    }
    const float sdv = sqrtf(v * (float)a.E / (float)(a.k * a.E - 1));
    mean[q] = __fadd_rn(__fmul_rn(one_minus_alpha, m), __fmul_rn(a.alpha, mean[q]));
    std[q] = __fadd_rn(__fmul_rn(one_minus_alpha, sdv), __fmul_rn(a.alpha, std[q]));
  }
}

__global__ void finish_kernel(const FinishArgs a, const int pick_action, const int shift_mean, const int reset_mean,
                              const float std_scale) {
  extern __shared__ float sm[];
  const int env = blockIdx.x, tid = threadIdx.x;
  const int HU = a.H * a.U;
  float* mean = a.mean + (size_t)env * HU;
  float* std = a.std + (size_t)env * HU;
  if (pick_action) {
    if (env == 0 && tid == 0 && a.step_ctr != nullptr) *a.step_ctr += 1ull;  // next MPC step draws from fresh Philox streams
    for (int u = tid; u < a.U; u += blockDim.x)
      a.action_out[env * a.U + u] = a.execute_best ? a.elite_actions[(size_t)env * a.k * HU + u] : mean[u];
  }
  if (shift_mean) {  // mean_[:-1] = mean_[1:], last entry kept (mpc_torch.py:413-417)
    for (int i = tid; i < HU; i += blockDim.x) sm[i] = i + a.U < HU ? mean[i + a.U] : mean[i];
    __syncthreads();
    for (int i = tid; i < HU; i += blockDim.x) mean[i] = sm[i];
  }
  for (int i = tid; i < HU; i += blockDim.x) {
    const int u = i % a.U;
    if (reset_mean) mean[i] = a.relative_init ? (a.high[u] + a.low[u]) * 2.0f : 0.f;  // reset_mean :189-195 (sic: *2)
    std[i] = a.relative_init ? (a.high[u] - a.low[u]) * std_scale : a.init_std;      // reset_std  :197-202
  }
}

__global__ void repack_kernel(const float* __restrict__ src, float* __restrict__ dst, int E, int rows, int cols,
                              long long member_stride, int dst_off, int dst_ld) {
  const long long total = (long long)E * rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    const int r = (int)((i / cols) % rows);
    const int e = (int)(i / ((long long)rows * cols));
    dst[e * member_stride + dst_off + (long long)r * dst_ld + c] = src[i];
  }
}

}  // namespace

// dynamic shared memory of sample_kernel: 32 rows of normals + (colored noise only) the [2F x H] irDFT matrix
size_t sample_smem_bytes(int H, int colored) {
  const int F = H / 2 + 1;
  const int nz = colored ? 2 * F : H;
  const int nzp = (nz + 3) & ~3;
  return ((size_t)kPairsPerBlock * nzp + (colored ? (size_t)2 * F * H : 0)) * sizeof(float);
}

cudaError_t launch_sample(const SampleArgs& a, cudaStream_t st) {
  const long long npairs = (long long)a.n * a.U;
  if (npairs == 0) return cudaSuccess;
  const size_t smem = sample_smem_bytes(a.H, a.colored);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int blocks = (int)((npairs + kPairsPerBlock - 1) / kPairsPerBlock);
  sample_kernel<<<blocks, kSampleThreads, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_fixup(const FixupArgs& a, cudaStream_t st) {
  fixup_kernel<<<a.nE, 128, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_costs(const CostArgs& a, const CostDesc& cd, cudaStream_t st) {
  if (a.n == 0) return cudaSuccess;
  const int wpb = 8;
  if (a.E <= 5) cost_kernel<5><<<(a.n + wpb - 1) / wpb, wpb * 32, 0, st>>>(a, cd);
  else cost_kernel<kMaxEnsemble><<<(a.n + wpb - 1) / wpb, wpb * 32, 0, st>>>(a, cd);
  return cudaGetLastError();
}

cudaError_t launch_costs_compact(const costdev::FusedCost& f, const float* start_obs, const float* goal, int traj_per_start,
                                 int traj_per_goal, int n, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const int wpb = 8;
  const bool need_env = !f.cd.curiosity || f.cd.extrinsic;
  const size_t smem = need_env ? (size_t)wpb * f.dims.E * f.dims.H * sizeof(float) : 0;
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  if (f.dims.E <= 5) cost_compact_kernel<5><<<(n + wpb - 1) / wpb, wpb * 32, smem, st>>>(f, start_obs, goal, traj_per_start, traj_per_goal, n);
  else cost_compact_kernel<kMaxEnsemble><<<(n + wpb - 1) / wpb, wpb * 32, smem, st>>>(f, start_obs, goal, traj_per_start, traj_per_goal, n);
  return cudaGetLastError();
}

cudaError_t launch_topk(const TopkArgs& a, cudaStream_t st) {
  if (a.k > 64) return cudaErrorInvalidValue;  // picked[] in shared memory; update_kernel has the same bound
  topk_kernel<<<a.nE, kTopkThreads, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_update(const UpdateArgs& a, cudaStream_t st) {
  if (a.world * a.k + a.n_kept_in > kMaxCand || a.k > 64) return cudaErrorInvalidValue;
  const size_t smem = (size_t)a.k * (a.E + a.H * a.U) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const float oma = (float)(1.0 - (double)a.alpha);
  update_kernel<<<a.nE, kUpdThreads, smem, st>>>(a, oma);
  return cudaGetLastError();
}

// one block per destination rank: coalesced stores of the records into the peer's buffer, system-scope fence, flag
__global__ void __launch_bounds__(256) exchange_kernel(const ExchangeArgs a) {
  const int dst = blockIdx.x;
  const unsigned long long x = *a.step_ctr * (unsigned long long)a.iters_per_step + (unsigned long long)a.iter;
  float* out = a.peer_buf[dst] + (x & 1ull) * a.half_stride + (size_t)a.rank * a.count;
  for (int i = threadIdx.x; i < a.count; i += blockDim.x) out[i] = a.cand[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile unsigned long long*>(a.peer_flags[dst] + a.rank) = x + 1ull;
    __threadfence_system();
  }
}

cudaError_t launch_exchange(const ExchangeArgs& a, cudaStream_t st) {
  exchange_kernel<<<a.world, 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_finish(const FinishArgs& a, cudaStream_t st) {
  const float std_scale = (float)((double)a.init_std / 2.0);
  finish_kernel<<<a.nE, 128, (size_t)a.H * a.U * sizeof(float), st>>>(a, 1, 1, 0, std_scale);
  return cudaGetLastError();
}

cudaError_t launch_reset_dist(const FinishArgs& a, cudaStream_t st) {
  const float std_scale = (float)((double)a.init_std / 2.0);
  finish_kernel<<<a.nE, 128, (size_t)a.H * a.U * sizeof(float), st>>>(a, 0, 0, 1, std_scale);
  return cudaGetLastError();
}

cudaError_t launch_repack(const float* src, float* dst, int E, int rows, int cols, long long member_stride, int dst_off,
                          int dst_ld, cudaStream_t st) {
  const long long total = (long long)E * rows * cols;
  if (total == 0) return cudaSuccess;
  const int blocks = (int)((total + 255) / 256 < 1024 ? (total + 255) / 256 : 1024);
  repack_kernel<<<blocks, 256, 0, st>>>(src, dst, E, rows, cols, member_stride, dst_off, dst_ld);
  return cudaGetLastError();
}

}  // namespace ceeus
