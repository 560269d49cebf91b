// Micro-benchmark: how fast can ONE thread issue small tcgen05.mma instructions (sm_100a), and how fast does the tensor pipe
// retire them?  The attention kernel issues M=128, N=64, K=16 MMAs (32 tensor-pipe cycles each at full rate); its timeline
// (tools/attn_trace.py) shows ~130 cycles per issued MMA.  This separates issue cost from execution cost for
//   SS (A and B from shared memory) vs TS (A from tensor memory), N = 64 / 128 / 256, one or two issuing warps per CTA,
//   one or two CTAs per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../tpdm_b200/csrc -o mma_issue_rate mma_issue_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "common.cuh"

using namespace tpdm;

// MODE 0: SS   MODE 1: TS (A = 128 lanes x 8 columns of TMEM)
// `hammer` != 0: warps 4..7 (one per SM sub-partition, like the softmax warps of the attention kernel) issue MUFU.EX2 + FFMA back to
// back while the MMAs are being issued: does a busy sub-partition slow the issuing thread down?
// `commit_every` > 0: a tcgen05.commit (mbarrier arrive on completion) after every commit_every MMAs, as a pipelined kernel issues
// them -- does the commit cost the issuing thread time?
// `converged` != 0: the MMAs are issued by the whole (converged) warp with the lane election inside the asm statement
// (umma_ss_elect / umma_ts_elect in common.cuh) instead of from inside `if (lane == 0)`: no waterfall loop around UTCHMMA.
template <int MODE, int N>
__global__ void __launch_bounds__(256) k(uint32_t* out, int n_mma, int issuers, int hammer, int commit_every = 0, int converged = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint64_t bar2[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&bar2[0], 1);
    mbar_init(&bar2[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc<256>(&slot);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  constexpr uint32_t idesc = make_idesc_bf16(128, N, false);
  const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 16384);
  long long t0 = 0, t1 = 0, t2 = 0;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp >= 4) {
    if (hammer) {
      float a[8], acc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i);
      while (!stop) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float y;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a[i]));
            acc = fmaf(y, 0.5f, acc);
            a[i] = y * 0.25f;
          }
        }
      }
      if (acc == 123.456f) out[32 + threadIdx.x] = 1;
    }
  } else if (warp < issuers) {
    const uint32_t dcol = tmem + (warp * (N <= 64 ? 64 : 0));    // distinct accumulators when they fit (N = 64), else a shared one
    __syncwarp();
    t0 = clock64();
    if (converged) {
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t off = (i & 3) * 32;
        if (MODE == 0)
          umma_ss_elect(dcol, make_smem_desc_sw128(a_base + off, 16, 1024), make_smem_desc_sw128(b_base + off, 16, 1024), idesc, i ? 1u : 0u);
        else
          umma_ts_elect(dcol, tmem + 192 + (i & 3) * 8, make_smem_desc_sw128(b_base + off, 16, 1024), idesc, i ? 1u : 0u);
      }
      t1 = clock64();
      umma_commit_elect(&bar[warp]);
    } else if (lane == 0) {
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t off = (i & 3) * 32;
        if (MODE == 0)
          umma_ss(dcol, make_smem_desc_sw128(a_base + off, 16, 1024), make_smem_desc_sw128(b_base + off, 16, 1024), idesc, i ? 1u : 0u);
        else
          umma_ts(dcol, tmem + 192 + (i & 3) * 8, make_smem_desc_sw128(b_base + off, 16, 1024), idesc, i ? 1u : 0u);
        if (commit_every > 0 && (i + 1) % commit_every == 0) umma_commit(&bar2[warp]);   // nobody waits on it: only the issue cost counts
      }
      t1 = clock64();
      umma_commit(&bar[warp]);
    }
    __syncwarp();
    mbar_wait(&bar[warp], 0);
    t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) {
      out[warp * 2] = static_cast<uint32_t>(t1 - t0);
      out[warp * 2 + 1] = static_cast<uint32_t>(t2 - t0);
    }
    __syncwarp();
    if (lane == 0) atomicAdd(const_cast<int*>(&stop), 1);
  }
  if (warp < 4 && warp >= issuers && lane == 0) atomicAdd(const_cast<int*>(&stop), 0);
  if (hammer && warp >= 4) {
    // nothing: the loop above exits once an issuer has finished (stop != 0); with two issuers the slower one is still timed correctly
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

template <int MODE, int N>
void run(const char* name, uint32_t* d, int ctas_per_sm, int issuers, int hammer = 0, int commit_every = 0, int converged = 0) {
  const int n_mma = 256;
  const size_t smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(k<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  for (int rep = 0; rep < 2; ++rep) {
    k<MODE, N><<<148 * ctas_per_sm, 256, smem>>>(d, n_mma, issuers, hammer, commit_every, converged);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", name, cudaGetErrorString(e));
      return;
    }
  }
  uint32_t h[4];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double floor_clk = 128.0 * N / 256.0;   // tensor-pipe cycles of one M=128, K=16 MMA at full rate
  if (commit_every) printf("[commit after every %d MMAs] ", commit_every);
  if (converged) printf("[converged warp, elected lane] ");
  printf("%s%-4s N=%3d  %d CTA/SM x %d issuer(s): issue %6.1f clk/MMA, issue+drain %6.1f clk/MMA  (full-rate pipe time %4.0f clk/MMA; SM share "
         "%4.0f)\n",
         hammer ? "[4 MUFU-bound warps per CTA] " : "", name, N, ctas_per_sm, issuers, h[0] / double(n_mma), h[1] / double(n_mma), floor_clk, floor_clk * ctas_per_sm * issuers);
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 4096);
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int iss = 1; iss <= 2; ++iss) {
      run<0, 64>("SS", d, ctas, iss);
      run<1, 64>("TS", d, ctas, iss);
      run<0, 128>("SS", d, ctas, iss);
      run<1, 128>("TS", d, ctas, iss);
      if (iss == 1) run<0, 256>("SS", d, ctas, iss);
    }
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int iss = 1; iss <= 2; ++iss) {
      run<0, 64>("SS", d, ctas, iss, 0, 0, 1);
      run<1, 64>("TS", d, ctas, iss, 0, 0, 1);
      run<0, 128>("SS", d, ctas, iss, 0, 0, 1);
      if (iss == 1) run<0, 256>("SS", d, ctas, iss, 0, 0, 1);
    }
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int ce = 1; ce <= 4; ce *= 2) {
      run<0, 128>("SS", d, ctas, 1, 0, ce);
      run<1, 64>("TS", d, ctas, 1, 0, ce);
      run<1, 64>("TS", d, ctas, 2, 0, ce);
    }
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int iss = 1; iss <= 2; ++iss) {
      run<0, 64>("SS", d, ctas, iss, 1);
      run<1, 64>("TS", d, ctas, iss, 1);
      run<0, 128>("SS", d, ctas, iss, 1);
    }
  return 0;
}
