// Self-test of the tcgen05 building blocks the tensor-core rollout kernel is made of: one D = A * B^T tile with the
// same shared-memory layout, descriptors, TMEM access pattern, commit/mbarrier protocol (1-CTA and 2-CTA pair).
// Exposed through the C ABI as ceeus_selftest_umma so the GPU tests can pin every piece against a plain GEMM.
#include <cuda_fp16.h>

#include <string>

#include "common.cuh"
#include "debug_tools.h"
#include "tc_common.cuh"

namespace ceeus {
namespace {

using namespace tc;

// A [rows_total][K] fp16 row-major, Bt [N][K] fp16 row-major (= W^T), D [rows_total][N] fp32.
// CG = 1: one CTA, rows_total = 128.  CG = 2: cluster of 2, rows_total = 256, each CTA holds half of the N rows of Bt.
template <int CG>
__global__ void __launch_bounds__(160, 1) umma_probe_kernel(const __half* __restrict__ A, const __half* __restrict__ Bt,
                                                            float* __restrict__ D, int K, int N, int variant, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  const int nloc = N / CG;           // B rows held by this CTA
  const int kc = K / 8;              // 16-byte K chunks
  uint8_t* sA = smem;                // [kc][128][16B]
  uint8_t* sB = sA + (size_t)kc * 128 * 16;  // [kc][nloc][16B]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (size_t)kc * nloc * 16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

  for (int idx = tid; idx < 128 * kc; idx += blockDim.x) {
    const int r = idx % 128, c = idx / 128;
    *reinterpret_cast<uint4*>(sA + (size_t)c * 2048 + r * 16) =
        *reinterpret_cast<const uint4*>(A + (size_t)(rank * 128 + r) * K + c * 8);
  }
  for (int idx = tid; idx < nloc * kc; idx += blockDim.x) {
    const int n = idx % nloc, c = idx / nloc;
    *reinterpret_cast<uint4*>(sB + (size_t)c * nloc * 16 + n * 16) =
        *reinterpret_cast<const uint4*>(Bt + (size_t)(rank * nloc + n) * K + c * 8);
  }
  fence_proxy_async();
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (variant & 2) ncols = 512;
  if (warp == 4) {
    if (lane == 0) {
      mbar_init(bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<CG>(tmem_slot, ncols);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool a_in_tmem = (variant & 2) != 0;
  const uint32_t a_col0 = 256;  // TS variant: A tile lives in TMEM columns [256, 256 + K/2), two fp16 per 32-bit cell
  if (a_in_tmem) {
    if (warp < 4) {
      const __half* arow = A + (size_t)(rank * 128 + warp * 32 + lane) * K;
      for (int ks = 0; ks < K / 16; ++ks) {
        uint32_t v[8];
        const uint4 lo = *reinterpret_cast<const uint4*>(arow + ks * 16);
        const uint4 hi = *reinterpret_cast<const uint4*>(arow + ks * 16 + 8);
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
        tmem_st8(tmem_base + ((uint32_t)(warp * 32) << 16) + a_col0 + ks * 8, v);
      }
      tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync();
    tc_fence_after();
  }

  if (warp == 4 && lane == 0 && rank == 0) {
    uint32_t lbo_a = 2048, sbo_a = 128, lbo_b = nloc * 16, sbo_b = 128;
    if (variant & 1) {
      uint32_t t = lbo_a; lbo_a = sbo_a; sbo_a = t;
      t = lbo_b; lbo_b = sbo_b; sbo_b = t;
    }
    const uint32_t idesc = make_idesc_f16(128 * CG, N);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t da = make_smem_desc(smem_u32(sA) + ks * 2 * 2048, lbo_a, sbo_a);
      const uint64_t db = make_smem_desc(smem_u32(sB) + ks * 2 * nloc * 16, lbo_b, sbo_b);
      if (a_in_tmem) umma_f16_ts<CG>(tmem_base, tmem_base + a_col0 + ks * 8, db, idesc, ks > 0 ? 1u : 0u);
      else umma_f16_ss<CG>(tmem_base, da, db, idesc, ks > 0 ? 1u : 0u);
    }
    if (CG == 1) umma_commit_cg1(bar);
    else umma_commit_cg2_mc(bar, 0x3);
  }
  if (warp < 4) {
    const bool ok = mbar_wait(bar, 0, status, 100 + (int)rank);
    tc_fence_after();
    if (ok) {
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_wait_ld();
        float* out = D + (size_t)(rank * 128 + warp * 32 + lane) * N + c0;
        for (int j = 0; j < 32 && c0 + j < N; ++j) out[j] = __uint_as_float(v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (warp == 4) tmem_dealloc<CG>(tmem_base, ncols);
}

// TMEM <-> register-file micro-benchmark: clock64 around N x tcgen05.ld/st.32x32b.x32 (+ one wait) with 1, 4 or 8 warps active.
__global__ void __launch_bounds__(256, 1) tmem_bench_kernel(long long* out) {
  __shared__ uint32_t tmem_slot;
  __shared__ float sink[256];
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    tmem_alloc<1>(&tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t a[32], b[32], c[32], d[32];
  for (int j = 0; j < 32; ++j) { a[j] = tid + j; b[j] = tid * 3 + j; c[j] = tid * 5 + j; d[j] = tid * 7 + j; }
  // fill
  for (int c0 = 0; c0 < 512; c0 += 32) tmem_st32(base + c0, a);
  tmem_wait_st();
  float acc = 0.f;
  int slot = 0;
  auto run = [&](int nwarps, int nld, bool serial, bool store) {
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nwarps) {
      for (int i = 0; i < nld; i += 4) {
        const uint32_t col = (uint32_t)((i * 32) & 511);
        if (store) {
          tmem_st32(base + col, a); if (serial) tmem_wait_st();
          if (i + 1 < nld) { tmem_st32(base + ((col + 32) & 511), b); if (serial) tmem_wait_st(); }
          if (i + 2 < nld) { tmem_st32(base + ((col + 64) & 511), c); if (serial) tmem_wait_st(); }
          if (i + 3 < nld) { tmem_st32(base + ((col + 96) & 511), d); if (serial) tmem_wait_st(); }
        } else {
          tmem_ld32(base + col, a); if (serial) tmem_wait_ld();
          if (i + 1 < nld) { tmem_ld32(base + ((col + 32) & 511), b); if (serial) tmem_wait_ld(); }
          if (i + 2 < nld) { tmem_ld32(base + ((col + 64) & 511), c); if (serial) tmem_wait_ld(); }
          if (i + 3 < nld) { tmem_ld32(base + ((col + 96) & 511), d); if (serial) tmem_wait_ld(); }
        }
      }
      if (store) tmem_wait_st(); else tmem_wait_ld();
      acc += __uint_as_float(a[1] ^ b[2] ^ c[3] ^ d[4]);
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) out[slot] = t1 - t0;
    ++slot;
  };
  run(1, 1, false, false);   // 0: latency of one ld32 + wait
  run(1, 4, false, false);   // 1
  run(1, 16, false, false);  // 2
  run(1, 64, false, false);  // 3
  run(4, 4, false, false);   // 4
  run(4, 16, false, false);  // 5
  run(4, 64, false, false);  // 6
  run(8, 16, false, false);  // 7
  run(8, 64, false, false);  // 8
  run(1, 16, true, false);   // 9: serialised ld + wait
  run(4, 16, true, false);   // 10
  run(1, 1, false, true);    // 11: st latency
  run(4, 16, false, true);   // 12
  run(4, 64, false, true);   // 13
  run(8, 64, false, true);   // 14
  sink[tid] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem_slot, 512);
  if (sink[(tid + 1) & 255] == 1234.5f) out[31] = 1;
}

// Does a tcgen05.ld / tcgen05.st queue behind in-flight tcgen05.mma?  Warp 4 issues `nmma` MMAs (M=128, N=128, K=16, operands
// = whatever is in shared memory) into columns [256, 384); warps 0-3 time one ld32 (+wait) / st32 (+wait) of columns [0, 32)
// issued right after.  out[0] = ld latency, out[1] = st latency, out[2] = cycles until the MMA batch committed.
__global__ void __launch_bounds__(160, 1) tmem_mma_contention_kernel(long long* out, int nmma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 16384; i += 160) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  fence_proxy_async();
  if (warp == 4) {
    if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc<1>(&tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_slot;
  uint32_t a[32];
  for (int j = 0; j < 32; ++j) a[j] = tid + j;
  if (warp < 4) { tmem_st32(base + ((uint32_t)(warp * 32) << 16), a); tmem_wait_st(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const long long t0 = clock64();
  if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, 128);
      for (int i = 0; i < nmma; ++i) {
        const uint64_t da = make_smem_desc(smem_u32(smem), 2048, 128);
        const uint64_t db = make_smem_desc(smem_u32(smem) + 32768, 2048, 128);
        umma_f16_ss<1>(base + 256, da, db, idesc, i > 0 ? 1u : 0u);
      }
      umma_commit_cg1(&bar);
      mbar_wait(&bar, 0, nullptr, 0);
      out[2] = clock64() - t0;
    }
  } else {
    // small delay so that the MMAs are in flight
    for (int i = 0; i < 20; ++i) a[i & 31] += (uint32_t)clock64() & 1u;
    const long long t1 = clock64();
    tmem_ld32(base + ((uint32_t)(warp * 32) << 16), a);
    tmem_wait_ld();
    const long long t2 = clock64();
    tmem_st32(base + ((uint32_t)(warp * 32) << 16) + 64, a);
    tmem_wait_st();
    const long long t3 = clock64();
    if (tid == 0) { out[0] = t2 - t1; out[1] = t3 - t2; out[3] = t1 - t0; }
    if (a[3] == 0x12345678u) out[5] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<1>(base, 512);
}

}  // namespace
}  // namespace ceeus

extern "C" int ceeus_selftest_tmem_contention(int nmma, long long* out_host /* [8] */) {
  using namespace ceeus;
  if (!out_host) return CEEUS_ERR_INVALID;
  long long* d = nullptr;
  if (cudaMalloc(&d, 8 * sizeof(long long)) != cudaSuccess) return CEEUS_ERR_CUDA;
  cudaMemset(d, 0, 8 * sizeof(long long));
  cudaFuncSetAttribute(tmem_mma_contention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  tmem_mma_contention_kernel<<<1, 160, 65536>>>(d, nmma);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out_host, d, 8 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? CEEUS_OK : CEEUS_ERR_CUDA;
}

extern "C" int ceeus_selftest_tmem(long long* out_host /* [32] */) {
  using namespace ceeus;
  if (!out_host) return CEEUS_ERR_INVALID;
  long long* d = nullptr;
  if (cudaMalloc(&d, 32 * sizeof(long long)) != cudaSuccess) return CEEUS_ERR_CUDA;
  cudaMemset(d, 0, 32 * sizeof(long long));
  tmem_bench_kernel<<<1, 256>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out_host, d, 32 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? CEEUS_OK : CEEUS_ERR_CUDA;
}

extern "C" int ceeus_selftest_umma(int cta_group, const void* a_dev, const void* bt_dev, float* d_dev, int K, int N,
                                   int variant, int* status_host) {
  using namespace ceeus;
  if (!a_dev || !bt_dev || !d_dev || !status_host) return CEEUS_ERR_INVALID;
  if ((cta_group != 1 && cta_group != 2) || K % 16 != 0 || K < 16 || K > 256 || N % 16 != 0 || N < 16 || N > 256)
    return CEEUS_ERR_INVALID;
  int* d_status = nullptr;
  if (cudaMalloc(&d_status, sizeof(int)) != cudaSuccess) return CEEUS_ERR_CUDA;
  cudaMemset(d_status, 0, sizeof(int));
  const size_t smem = (size_t)(K / 8) * 128 * 16 + (size_t)(K / 8) * (N / cta_group) * 16 + 64;
  cudaError_t e;
  if (cta_group == 1) {
    e = cudaFuncSetAttribute(umma_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
      umma_probe_kernel<1><<<1, 160, smem>>>((const __half*)a_dev, (const __half*)bt_dev, d_dev, K, N, variant, d_status);
      e = cudaGetLastError();
    }
  } else {
    e = cudaFuncSetAttribute(umma_probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(2);
      cfg.blockDim = dim3(160);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, umma_probe_kernel<2>, (const __half*)a_dev, (const __half*)bt_dev, d_dev, K, N, variant, d_status);
    }
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  int st = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&st, d_status, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d_status);
  *status_host = e == cudaSuccess ? st : -(int)e;
  return e == cudaSuccess ? CEEUS_OK : CEEUS_ERR_CUDA;
}

// Host-only check of the tensor-core rollout tiling (kernels.cuh: tc_tile / tc_group): how often every trajectory of a launch over
// n trajectories is covered by the groups of all (pair, cluster rank, slot, round), per ensemble member.  cover_out[e * n + p] must
// come out as exactly 1.  Returns the largest group size seen (must be <= 4 * tpw), or a negative error.  No CUDA call.
#include "kernels.cuh"
extern "C" int ceeus_probe_tc_partition(int num_sms, int E, int n, int n_obj, int hidden, int grow, int32_t* cover_out /* [E*n] */,
                                        int32_t* issuer_rows_out /* [2]: trajectories on issuing warps in pure / in all rounds */) {
  using namespace ceeus;
  if (!cover_out || n < 1 || E < 1 || (hidden != 64 && hidden != 128) || n_obj + (grow ? 1 : 0) > 32) return CEEUS_ERR_INVALID;
  const int NS = 512 / (2 * hidden), tpw = 32 / (n_obj + (grow ? 1 : 0)), Tmax = 4 * tpw, pairs = num_sms / 2;
  for (int i = 0; i < E * n; ++i) cover_out[i] = 0;
  int gmax = 0, on_issuer_pure = 0, on_issuer_all = 0;
  for (int pair = 0; pair < pairs; ++pair) {
    const TcTile t = tc_tile(pairs, E, pair, n, NS, Tmax, tpw);
    for (int rank = 0; rank < 2; ++rank)
      for (int s = 0; s < NS; ++s)
        for (int rd = 0; rd < t.rounds; ++rd) {
          int p0, Gc;
          tc_group(t, n, NS, tpw, rank, s, rd, 0, &p0, &Gc);
          if (Gc > gmax) gmax = Gc;
          for (int p = p0; p < p0 + Gc; ++p) {
            if (p < 0 || p >= n) return CEEUS_ERR_STATE;
            ++cover_out[t.e * n + p];
          }
          // trajectories that land on the MMA-issuing warp (the last-filled warp of a leader-CTA slot)
          const int over = Gc - 3 * tpw;
          if (rank == 0 && over > 0) {
            on_issuer_all += over;
            if (t.uneven == 1 || (t.uneven == 2 && rd < t.k_pure)) on_issuer_pure += over;
          }
        }
  }
  if (issuer_rows_out) { issuer_rows_out[0] = on_issuer_pure; issuer_rows_out[1] = on_issuer_all; }
  return gmax;
}
