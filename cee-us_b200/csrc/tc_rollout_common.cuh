// Device helpers shared by the tensor-core rollout kernels (rollout_tc.cu: GNN ensemble, rollout_mlp_tc.cu: dense MLP
// ensemble): bounded mbarrier waits with an error flag, the "A operand ready" signal, activations.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace ceeus {
namespace {
using namespace tc;

constexpr uint32_t FULL = 0xffffffffu;
__device__ __forceinline__ float act_fn(float x, int act) {
  switch (act) {
    case CEEUS_ACT_RELU: return fmaxf(x, 0.f);
    case CEEUS_ACT_SILU: return x / (1.f + __expf(-x));
    case CEEUS_ACT_TANH: return tanhf(x);
    default: return x;
  }
}

struct Ctx {
  volatile int* abort;  // per CTA
  int* err;             // global
  uint32_t a_ready_leader[4];  // cluster addresses of the leader's a_ready barriers (one per slot)
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

// Spinning variant for the MMA-issuing warp (reacts within a few cycles of the last arrival).
__device__ __forceinline__ bool wait_bar_spin(const Ctx& cx, uint64_t* bar, uint32_t parity, int code) {
  for (uint32_t it = 0; it < (1u << 28); ++it) {
    if (mbar_test_wait(bar, parity)) return true;
    if ((it & 4095u) == 4095u && *cx.abort) return false;
  }
  *cx.abort = 1;
  atomicExch(cx.err, code);
  return false;
}

// The A operand of this warpgroup's next MMA stage is complete in TMEM: retire the tcgen05.st, order them (and the
// earlier tcgen05.ld of the accumulator) before the barrier, one arrival per warp on the leader's barrier.
__device__ __forceinline__ void signal_a(const Ctx& cx, int s, int lane) {
  tmem_wait_st();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster_relaxed(cx.a_ready_leader[s]);
}

}  // namespace
}  // namespace ceeus
