/* Diagnostics that are NOT part of the product ABI (include/ceeus_b200.h):
 *   - libceeus_b200_probe.so (csrc/tc_probe.cu, built next to the product library): self-tests of the tcgen05 / TMEM / mbarrier
 *     building blocks, used by tests/test_gpu_tc_probe.py;
 *   - a -DCEEUS_DEBUG_TOOLS build of the product sources (python cee-us_b200/build.py --debug -> libceeus_b200_dbg.so): the
 *     pipeline timeline probes of the rollout kernel and the epilogue micro-benchmark. */
#ifndef CEEUS_DEBUG_TOOLS_H
#define CEEUS_DEBUG_TOOLS_H
#include "../../include/ceeus_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* CEEUS_PREC_TC profiling aid: enable=1 makes thread 0 of CTA 0 record (tag, clock64) pairs at every pipeline stage of
 * the next rollouts; out_host (nullable) receives the first n_pairs pairs.  enable=0 releases the buffer. */
int ceeus_tc_debug_timeline(CeeusHandle h, int enable, long long* out_host, int n_pairs);
/* Diagnostic: cycles per hidden-layer epilogue (128 columns, LayerNorm + ReLU, TMEM -> TMEM) of one warp when
 * `warps_per_scheduler` (1..3) warps run it concurrently on every SM sub-partition; with_mma = 1 keeps the tensor core busy
 * with back-to-back MMAs from another warp.  cycles_host[0] = cycles per epilogue, [1] = MMAs issued meanwhile. */
int ceeus_selftest_epilogue(int warps_per_scheduler, int with_bias, int with_mma, long long* cycles_host);
/* CEEUS_PREC_TC profiling aid (needs ceeus_tc_debug_timeline enabled): flag 1 = slot 1 of every CTA idles. */
int ceeus_tc_debug_flag(CeeusHandle h, long long flag);
/* Diagnostic: one D[128*cta_group, N] = A[.,K] * Bt[N,K]^T tile (fp16 in, fp32 out) through exactly the tcgen05 / TMEM /
 * mbarrier building blocks of the CEEUS_PREC_TC rollout kernel (cta_group 1 = single CTA, 2 = CTA pair).  Synchronous.
 * status_host: 0 ok, >0 a bounded barrier wait expired, <0 CUDA error.  No reference counterpart (test hook). */
int ceeus_selftest_umma(int cta_group, const void* a_dev, const void* bt_dev, float* d_dev, int K, int N, int variant,
                        int* status_host);

/* Diagnostic: cycle counts of TMEM <-> register transfers (tcgen05.ld / tcgen05.st 32x32b.x32) for 1 / 4 / 8 active warps;
 * out_host receives 32 values (see tmem_bench_kernel in csrc/tc_probe.cu).  No reference counterpart (profiling aid). */
int ceeus_selftest_tmem(long long* out_host);
/* Diagnostic: latency of one tcgen05.ld / tcgen05.st issued while `nmma` tcgen05.mma are in flight (out_host[0..3]). */
int ceeus_selftest_tmem_contention(int nmma, long long* out_host);

#ifdef __cplusplus
}
#endif
#endif
