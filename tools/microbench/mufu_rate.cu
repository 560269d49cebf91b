// Micro-benchmark: MUFU.EX2 issue rate per SM for f32, f16x2 and bf16x2 operands (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, int iters) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0x3c003c00u + threadIdx.x + i;  // ~1.0 in f16x2 / small float bits
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        float x = __uint_as_float(a[i]), y;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        a[i] = __float_as_uint(y) & 0x3fffffffu;
      } else if (MODE == 1) {
        uint32_t y;
        asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(a[i]));
        a[i] = y & 0x3bff3bffu;
      } else {
        uint32_t y;
        asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(a[i]));
        a[i] = y & 0x3f7f3f7fu;
      }
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = static_cast<uint32_t>(t1 - t0);
}

int main() {
  uint32_t* d;
  const int blocks = 148, threads = 1024, iters = 2000;
  cudaMalloc(&d, (blocks * threads + 1) * 4);
  const char* names[3] = {"ex2.f32", "ex2.f16x2", "ex2.bf16x2"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<blocks, threads>>>(d, iters);
      if (mode == 1) k<1><<<blocks, threads>>>(d, iters);
      if (mode == 2) k<2><<<blocks, threads>>>(d, iters);
      cudaDeviceSynchronize();
    }
    uint32_t cyc;
    cudaMemcpy(&cyc, d + blocks * threads, 4, cudaMemcpyDeviceToHost);
    const double results = (mode == 0 ? 1.0 : 2.0) * 8.0 * iters * threads;  // exponentials per SM
    printf("%-11s %u cycles for %d iters x 8 ops x 32 warps/SM -> %.2f results/clk/SM (%.2f instr-lanes/clk/SM)\n", names[mode], cyc, iters,
           results / cyc, 8.0 * iters * threads / cyc);
  }
  return 0;
}
