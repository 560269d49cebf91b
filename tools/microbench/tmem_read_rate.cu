// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM on sm_100a, to substantiate the "TMEM read floor" that
// attention_tcgen05.cu quotes (S is read back in fp32: 4 bytes per score).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_read_rate tmem_read_rate.cu ; run on a B200.
//
// Every warp reads its own 32-lane quarter (warp % 4) over and over; variants: instruction width (x16 / x32 / x64 columns per
// instruction), loads in flight before tcgen05.wait::ld, 4 or 8 warps per CTA, 1 or 2 CTAs per SM.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

#define LD16(a, v)                                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"      \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),     \
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                        \
               : "r"(a))
#define LD32(a, v)                                                                                                              \
  asm volatile(                                                                                                                 \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"  \
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"                                                                      \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),    \
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),    \
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                                      \
      : "r"(a))
// 16 lanes x 256 bits per row, x8: also 32 registers per thread (the layout cuBLAS-style epilogues use)
#define LD16x256(a, v)                                                                                                          \
  asm volatile(                                                                                                                 \
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"  \
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"                                                                      \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),    \
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),    \
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                                      \
      : "r"(a))
#define WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory")

// MODE 0: x32, wait after every load      MODE 1: x32, two loads in flight     MODE 2: x16, four in flight
// MODE 3: 16x256b.x8, two in flight        MODE 4: x32 + 32 MUFU.EX2 per load (do the two pipes overlap inside one warp?)
template <int MODE, int COLS>
__global__ void k(uint32_t* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 64;  // warps 4..7: other columns
  uint32_t acc = 0;
  float facc = 0.f;
  uint32_t a[32], b[32];
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      LD32(base, a);
      WAIT_LD();
      acc ^= a[0] ^ a[31];
    } else if (MODE == 1) {
      LD32(base, a);
      LD32(base + 32, b);
      WAIT_LD();
      acc ^= a[0] ^ b[31];
    } else if (MODE == 2) {
      uint32_t c[16], d[16], e[16], f[16];
      LD16(base, c);
      LD16(base + 16, d);
      LD16(base + 32, e);
      LD16(base + 48, f);
      WAIT_LD();
      acc ^= c[0] ^ d[1] ^ e[2] ^ f[3];
    } else if (MODE == 3) {
      LD16x256(base, a);
      LD16x256(base + 32, b);
      WAIT_LD();
      acc ^= a[0] ^ b[31];
    } else {
      LD32(base, a);
      WAIT_LD();
      LD32(base + 32, b);        // in flight while the exponentials of `a` run
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float y;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(a[i] & 0x3fffffffu)));
        facc += y;
      }
      WAIT_LD();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float y;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(b[i] & 0x3fffffffu)));
        facc += y;
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(facc);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = static_cast<uint32_t>(t1 - t0);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(slot), "n"(COLS) : "memory");
}

template <int MODE>
void run(const char* name, uint32_t* d, int ctas_per_sm, int warps, double bytes_per_iter_per_thread) {
  const int iters = 4000, blocks = 148 * ctas_per_sm, threads = warps * 32;
  for (int rep = 0; rep < 2; ++rep) {
    k<MODE, 256><<<blocks, threads>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", name, cudaGetErrorString(e));
      return;
    }
  }
  uint32_t cyc;
  cudaMemcpy(&cyc, d + blocks * threads, 4, cudaMemcpyDeviceToHost);
  const double bytes_per_sm = bytes_per_iter_per_thread * iters * threads * ctas_per_sm;
  printf("%-34s %d CTA/SM x %d warps: %9u cycles -> %7.1f B/clk/SM  (%.1f clk per warp-instruction-equivalent of 4 KB)\n", name, ctas_per_sm,
         warps, cyc, bytes_per_sm / cyc, 4096.0 * cyc / (bytes_per_sm / (ctas_per_sm * warps)) );
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, (148 * 2 * 256 + 1) * 4);
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int warps = 4; warps <= 8; warps += 4) {
      run<0>("32x32b.x32, wait each", d, ctas, warps, 128);
      run<1>("32x32b.x32, 2 in flight", d, ctas, warps, 256);
      run<2>("32x32b.x16, 4 in flight", d, ctas, warps, 256);
      run<3>("16x256b.x8, 2 in flight", d, ctas, warps, 256);
      run<4>("32x32b.x32 + 32 ex2 per load", d, ctas, warps, 256);
    }
  return 0;
}
