// Micro-benchmark (manual diagnostic): issue cost per warp instruction of the conversion / packed-math instructions the
// rollout epilogue is made of, for 1, 2 and 4 warps per scheduler.   nvcc -arch=sm_100a -O3 -o cee-us_b200/build/ubench_pipes scripts/ubench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
template <int OP>
__global__ void k(long long* out, int reps) {
  float x[32];
  uint32_t acc = 0;
  for (int j = 0; j < 32; ++j) x[j] = 0.001f * (threadIdx.x + j);
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      if (OP == 0) {  // cvt.rn.relu.f16x2.f32
        uint32_t d; asm volatile("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(x[j + 1]), "f"(x[j])); acc ^= d;
      } else if (OP == 1) {  // cvt.rn.f16x2.f32
        uint32_t d; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(x[j + 1]), "f"(x[j])); acc ^= d;
      } else if (OP == 2) {  // fma.rn.f32x2
        uint64_t a, d; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(x[j]), "f"(x[j + 1]));
        asm volatile("fma.rn.f32x2 %0, %1, %1, %1;" : "=l"(d) : "l"(a)); acc ^= (uint32_t)d;
      } else if (OP == 3) {  // plain FFMA x2
        float a = x[j], b = x[j + 1]; asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a)); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(b));
        acc ^= __float_as_uint(a) ^ __float_as_uint(b);
      } else if (OP == 4) {  // cvt.rn.bf16x2.f32
        uint32_t d; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(x[j + 1]), "f"(x[j])); acc ^= d;
      } else if (OP == 5) {  // HADD2
        uint32_t a = __float_as_uint(x[j]), b = __float_as_uint(x[j + 1]), d; asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); acc ^= d;
      } else if (OP == 7) {  // tanh.approx.f32 (MUFU)
        float a, b; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(a) : "f"(x[j])); asm volatile("tanh.approx.f32 %0, %1;" : "=f"(b) : "f"(x[j + 1]));
        acc ^= __float_as_uint(a) ^ __float_as_uint(b);
      } else if (OP == 8) {  // SiLU as in the MLP epilogue: h = x/2; h + h*tanh(h); pack
        const float h0 = 0.5f * x[j], h1 = 0.5f * x[j + 1];
        float t0, t1; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0)); asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
        uint32_t d; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(fmaf(h1, t1, h1)), "f"(fmaf(h0, t0, h0))); acc ^= d;
      } else if (OP == 6) {  // shfl
        acc ^= __shfl_down_sync(0xffffffffu, __float_as_uint(x[j]), 1);
      }
    }
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) out[63] = acc;
  if (threadIdx.x == 0) out[0] = (t1 - t0);
}
int main() {
  long long* d; cudaMalloc(&d, 64 * 8);
  const char* names[] = {"cvt.rn.relu.f16x2.f32", "cvt.rn.f16x2.f32", "fma.rn.f32x2", "2x fma.rn.f32", "cvt.rn.bf16x2.f32", "add.f16x2", "shfl.down", "2x tanh.approx.f32", "SiLU pair (2 MUFU)"};
  const int reps = 200;
  for (int op = 0; op < 9; ++op)
    for (int w = 1; w <= 4; w *= 2) {
      const int threads = 128 * w;
      switch (op) { case 0: k<0><<<1, threads>>>(d, reps); break; case 1: k<1><<<1, threads>>>(d, reps); break; case 2: k<2><<<1, threads>>>(d, reps); break;
        case 3: k<3><<<1, threads>>>(d, reps); break; case 4: k<4><<<1, threads>>>(d, reps); break; case 5: k<5><<<1, threads>>>(d, reps); break; case 6: k<6><<<1, threads>>>(d, reps); break; case 7: k<7><<<1, threads>>>(d, reps); break; default: k<8><<<1, threads>>>(d, reps); }
      long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("%-24s warps/scheduler %d: %.2f cycles per warp-instruction (per scheduler: %.2f)\n", names[op], w, (double)h / (reps * 16), (double)h / (reps * 16) / w);
    }
  return 0;
}
