// tpdm_b200 -- VAE decode of the final latent: the step right after the adaptive loop,
// /root/reference/src/models/stable_diffusion_3/modeling_sd3_pnt.py:653-655 (un-scale, AutoencoderKL.decode, postprocess).
// AutoencoderKL is a diffusers class; the decoder topology restated here is documented in DESIGN.md section 7.
//
// Layout: activations NHWC bf16 (pixel-major rows of C channels) so that every 3x3 convolution is the implicit-GEMM mode of
// the tcgen05 GEMM kernel (4-D TMA over [C, W, H, B] with signed coordinates, zero fill = padding) and the 1x1 shortcut /
// attention projections are plain GEMMs.  GroupNorm + SiLU is a statistics pass (fp32 partials, fp64 atomics per 4-channel
// unit) and an apply pass.  The single-head mid-block attention (head width = C = 512, which does not fit the flash kernel's
// TMEM budget) runs as GEMMs over query-row chunks: S = Q K^T in fp32, row softmax -> bf16 P, O = P V.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <new>
#include <vector>

#include "../../include/tpdm_b200.h"
#include "common.cuh"
#include "host.h"
#include "kernels.h"

using namespace tpdm;

namespace {

constexpr int kCinPad = 64;   // latent channels are zero-padded to one 64-wide k-block
constexpr int kOutPad = 8;    // conv_out produces 8 fp32 columns per pixel (3 used)
constexpr int kAttnChunk = 16384;  // query rows per attention chunk (bounds the S / P workspace)

__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// latents NCHW fp32 [Cl][h][w] (one sample) -> z = latents / scaling + shift, NHWC bf16 [h][w][64]
__global__ void vae_prep_latent_kernel(const float* __restrict__ lat, bf16* __restrict__ z, int Cl, int hw, float inv_scale, float shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw * kCinPad) return;
  const int p = i / kCinPad, c = i - p * kCinPad;
  z[i] = __float2bfloat16(c < Cl ? lat[static_cast<long long>(c) * hw + p] * inv_scale + shift : 0.f);
}

// GroupNorm statistics over NHWC bf16 x[P][C]: per 4-channel unit u (= 8 bytes) sum and sum of squares, accumulated
// into stats[u] (double2, zeroed by the caller).  A thread owns one 16-byte column slot (2 units) and strides over pixels.
__global__ void __launch_bounds__(256) vae_gn_stats_kernel(const bf16* __restrict__ x, long long P, int C, double* __restrict__ stats) {
  const int slots = C / 8;                      // 16-byte slots per pixel
  const int slot = threadIdx.x % slots;         // blockDim.x is a multiple of slots (checked by the launcher)
  const int ppb = blockDim.x / slots;           // pixels per block-iteration
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
  const long long pstride = static_cast<long long>(gridDim.x) * ppb;
  for (long long p0 = static_cast<long long>(blockIdx.x) * ppb + threadIdx.x / slots; p0 < P; p0 += 4 * pstride) {
    uint4 vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long p = p0 + u * pstride;
      vv[u] = p < P ? ldg_nc_u4(x + p * C + slot * 8) : make_uint4(0u, 0u, 0u, 0u);   // zeros add nothing to either sum
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint4 v = vv[u];
      const float a0 = __uint_as_float(v.x << 16), a1 = __uint_as_float(v.x & 0xffff0000u), a2 = __uint_as_float(v.y << 16),
                  a3 = __uint_as_float(v.y & 0xffff0000u), b0 = __uint_as_float(v.z << 16), b1 = __uint_as_float(v.z & 0xffff0000u),
                  b2 = __uint_as_float(v.w << 16), b3 = __uint_as_float(v.w & 0xffff0000u);
      s0 += (a0 + a1) + (a2 + a3);
      q0 += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      s1 += (b0 + b1) + (b2 + b3);
      q1 += (b0 * b0 + b1 * b1) + (b2 * b2 + b3 * b3);
    }
  }
  __shared__ float red[256 * 4];
  red[threadIdx.x * 4 + 0] = s0;
  red[threadIdx.x * 4 + 1] = q0;
  red[threadIdx.x * 4 + 2] = s1;
  red[threadIdx.x * 4 + 3] = q1;
  __syncthreads();
  if (threadIdx.x < slots) {
    double a[4] = {0, 0, 0, 0};
    for (int t = threadIdx.x; t < blockDim.x; t += slots)
      for (int k = 0; k < 4; ++k) a[k] += red[t * 4 + k];
    atomicAdd(&stats[(2 * slot) * 2 + 0], a[0]);
    atomicAdd(&stats[(2 * slot) * 2 + 1], a[1]);
    atomicAdd(&stats[(2 * slot + 1) * 2 + 0], a[2]);
    atomicAdd(&stats[(2 * slot + 1) * 2 + 1], a[3]);
  }
}

__device__ __forceinline__ float silu(float v) { return __fdividef(v, 1.f + __expf(-v)); }  // MUFU.EX2 + MUFU.RCP: the IEEE division made the apply pass compute-bound

// y = (x - mean_g) * rstd_g * gamma[c] + beta[c], optional SiLU; x, y NHWC bf16 [P][C]; groups of C / G channels (>= 4)
__global__ void __launch_bounds__(256) vae_gn_apply_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long P, int C, int G,
                                                           const double* __restrict__ stats, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, int act) {
  __shared__ float mean_s[64], rstd_s[64];
  const int cpg = C / G, upg = cpg / 4;
  if (threadIdx.x < G) {
    double s = 0, q = 0;
    for (int u = 0; u < upg; ++u) {
      s += stats[(threadIdx.x * upg + u) * 2];
      q += stats[(threadIdx.x * upg + u) * 2 + 1];
    }
    const double n = static_cast<double>(P) * cpg, m = s / n;
    double var = q / n - m * m;
    var = var < 0 ? 0 : var;
    mean_s[threadIdx.x] = static_cast<float>(m);
    rstd_s[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  const int slots = C / 8;
  const long long total = P * slots;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  // four 16-byte loads in flight per thread (a single load per iteration reached 43 % of the HBM rate)
  for (long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < total) v[u] = ldg_nc_u4(x + i * 8);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i >= total) break;
      const int c0 = static_cast<int>(i % slots) * 8;
      float f[8] = {__uint_as_float(v[u].x << 16), __uint_as_float(v[u].x & 0xffff0000u), __uint_as_float(v[u].y << 16),
                    __uint_as_float(v[u].y & 0xffff0000u), __uint_as_float(v[u].z << 16), __uint_as_float(v[u].z & 0xffff0000u),
                    __uint_as_float(v[u].w << 16), __uint_as_float(v[u].w & 0xffff0000u)};
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + c0), g1 = *reinterpret_cast<const float4*>(gamma + c0 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + c0), b1 = *reinterpret_cast<const float4*>(beta + c0 + 4);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const int ga = c0 / cpg, gb = (c0 + 4) / cpg;  // the two 4-channel units of this slot may sit in different groups
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int g = e < 4 ? ga : gb;
        const float r = (f[e] - mean_s[g]) * rstd_s[g] * gm[e] + bt[e];
        f[e] = act ? silu(r) : r;
      }
      uint4 w;
      w.x = pack_bf16x2(f[0], f[1]);
      w.y = pack_bf16x2(f[2], f[3]);
      w.z = pack_bf16x2(f[4], f[5]);
      w.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(y + i * 8) = w;
    }
  }
}

// nearest-neighbour x2: in [H][W][C] -> out [2H][2W][C] (F.interpolate(scale_factor=2, mode="nearest"))
__global__ void __launch_bounds__(256) vae_upsample2x_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int H, int W, int C) {
  const int slots = C / 8;
  const long long total = static_cast<long long>(H) * W * slots;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int slot = static_cast<int>(i % slots);
    const long long p = i / slots;
    const int xx = static_cast<int>(p % W), yy = static_cast<int>(p / W);
    const uint4 v = *reinterpret_cast<const uint4*>(in + i * 8);
    bf16* o = out + ((static_cast<long long>(2 * yy) * (2 * W) + 2 * xx) * C + slot * 8);
    *reinterpret_cast<uint4*>(o) = v;
    *reinterpret_cast<uint4*>(o + C) = v;
    *reinterpret_cast<uint4*>(o + static_cast<long long>(2 * W) * C) = v;
    *reinterpret_cast<uint4*>(o + static_cast<long long>(2 * W) * C + C) = v;
  }
}

// P[r][:] = softmax(scale * S[r][:]) : fp32 [rows][n] -> bf16, one block per row, row cached in registers (n <= 256 * 64 * 4)
__global__ void __launch_bounds__(256) vae_softmax_rows_kernel(const float* __restrict__ S, bf16* __restrict__ Pm, int n, float scale_log2e) {
  const long long row = blockIdx.x;
  const float* s = S + row * n;
  bf16* o = Pm + row * n;
  __shared__ float red[8];
  __shared__ float bcast;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mx = -INFINITY;
  for (int k = threadIdx.x * 4; k < n; k += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(s + k);
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    bcast = m;
  }
  __syncthreads();
  mx = bcast;
  float sum = 0.f;
  for (int k = threadIdx.x * 4; k < n; k += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(s + k);
    sum += (exp2f((v.x - mx) * scale_log2e) + exp2f((v.y - mx) * scale_log2e)) + (exp2f((v.z - mx) * scale_log2e) + exp2f((v.w - mx) * scale_log2e));
  }
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    bcast = 1.f / t;
  }
  __syncthreads();
  const float inv = bcast;
  for (int k = threadIdx.x * 4; k < n; k += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(s + k);
    uint2 w;
    w.x = pack_bf16x2(exp2f((v.x - mx) * scale_log2e) * inv, exp2f((v.y - mx) * scale_log2e) * inv);
    w.y = pack_bf16x2(exp2f((v.z - mx) * scale_log2e) * inv, exp2f((v.w - mx) * scale_log2e) * inv);
    *reinterpret_cast<uint2*>(o + k) = w;
  }
}

// conv_out result fp32 [P][8] -> image NCHW fp32 [3][P] and / or uint8 HWC [P][3] = round(clamp(v / 2 + 0.5, 0, 1) * 255)
__global__ void vae_image_out_kernel(const float* __restrict__ y, long long P, int oc, float* __restrict__ image, uint8_t* __restrict__ rgb) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float4 v = *reinterpret_cast<const float4*>(y + p * kOutPad);
  const float c[4] = {v.x, v.y, v.z, v.w};
  for (int k = 0; k < oc && k < 4; ++k) {
    if (image) image[static_cast<long long>(k) * P + p] = c[k];
    if (rgb) rgb[p * oc + k] = static_cast<uint8_t>(rintf(fminf(fmaxf(c[k] * 0.5f + 0.5f, 0.f), 1.f) * 255.f));
  }
}

struct Resnet {
  tpdm_vae_resnet w;
  int cin, cout;
};

}  // namespace

struct tpdm_vae {
  tpdm_vae_config cfg{};
  tpdm_vae_weights w{};
  std::vector<Resnet> resnets;   // mid.0, mid.1, then up_blocks in order
  std::vector<const void*> up_w;
  std::vector<const float*> up_b;
  bool has_weights = false;
  int levels = 0;
  int ch[8] = {0};               // reversed block_out_channels
};

namespace {

int blocks_for(long long n, int per_block, int cap) {
  long long b = (n + per_block - 1) / per_block;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

struct Buffers {
  bf16 *a, *b, *c, *d;   // four activation buffers of max size
  bf16 *q, *k, *vt, *o;  // attention projections
  float* S;
  bf16* Pm;
  float* yout;           // conv_out fp32 [P][8]
  double* stats;         // [C/4][2]
};

size_t align_up(size_t v) { return (v + 1023) & ~static_cast<size_t>(1023); }

// sizes for ONE sample of an (h, w) latent
void plan_sizes(const tpdm_vae* v, int h, int w, size_t* act_bytes, size_t* attn_rows) {
  size_t mx = 0;
  int H = h, W = w;
  mx = static_cast<size_t>(H) * W * v->ch[0] * 2;
  for (int i = 0; i < v->levels; ++i) {
    const size_t cur = static_cast<size_t>(H) * W * (i == 0 ? v->ch[0] : (v->ch[i - 1] > v->ch[i] ? v->ch[i - 1] : v->ch[i])) * 2;
    mx = cur > mx ? cur : mx;
    if (i != v->levels - 1) {
      H *= 2;
      W *= 2;
      const size_t up = static_cast<size_t>(H) * W * v->ch[i] * 2;
      mx = up > mx ? up : mx;
    }
  }
  *act_bytes = align_up(mx);
  const size_t n = static_cast<size_t>(h) * w;
  *attn_rows = n < static_cast<size_t>(kAttnChunk) ? n : kAttnChunk;
}

size_t workspace_bytes(const tpdm_vae* v, int h, int w) {
  size_t act, rows;
  plan_sizes(v, h, w, &act, &rows);
  const size_t n = static_cast<size_t>(h) * w, C = v->ch[0];
  int levels_up = v->levels - 1;
  const size_t Pout = n << (2 * levels_up);
  return 4 * act + 4 * align_up(n * C * 2) + align_up(rows * n * 4) + align_up(rows * n * 2) + align_up(Pout * kOutPad * 4) +
         align_up(static_cast<size_t>(512) * 2 * 8) + 4096;
}

int gn(const bf16* x, bf16* y, long long P, int C, int G, const float* gamma, const float* beta, int act, double* stats, cudaStream_t s) {
  TPDM_CHECK(C % 8 == 0 && C % G == 0 && (C / G) % 4 == 0 && G <= 64 && C / 8 <= 256 && 256 % (C / 8) == 0, TPDM_ERR_SHAPE,
             "vae group norm: unsupported channels %d / groups %d", C, G);
  TPDM_CUDA_OK(cudaMemsetAsync(stats, 0, static_cast<size_t>(C / 4) * 2 * sizeof(double), s));
  const int ppb = 256 / (C / 8);
  vae_gn_stats_kernel<<<blocks_for(P, ppb * 4, 8 * num_sms()), 256, 0, s>>>(x, P, C, stats);
  count_launch();
  vae_gn_apply_kernel<<<blocks_for(P * (C / 8), 256 * 4, 16 * num_sms()), 256, 0, s>>>(x, y, P, C, G, stats, gamma, beta, 1e-6f, act);
  count_launch();
  TPDM_CUDA_OK(cudaGetLastError());
  return 0;
}

int conv(const bf16* x, int H, int W, int Cin, const void* wt, const float* bias, int Cout, bf16* out, const bf16* res, cudaStream_t s) {
  GemmOp op;
  TPDM_TRY(gemm_op_init_conv3x3_hw(&op, x, 1, H, W, Cin, wt, Cout, res ? EPI_BIAS_ADD_BF16 : EPI_BIAS_BF16, out, Cout, bias, res));
  return gemm_launch(&op, 1, s);
}

int linear(const bf16* a, long long rows, int K, const void* wt, const float* bias, int N, void* out, int ldo, int epi, const bf16* res,
           cudaStream_t s) {
  GemmOp op;
  TPDM_TRY(gemm_op_init(&op, a, K, 0, static_cast<int>(rows), 1, K, wt, N, epi, out, 0, ldo, bias, nullptr, 0));
  op.res = res;
  return gemm_launch(&op, 1, s);
}

// x [H*W][cin] -> result [H*W][cout]; uses t1 / t2 as scratch; returns the buffer holding the result (x itself, or sc when
// the width changes)
int resnet(const tpdm_vae* v, const Resnet& r, bf16* x, bf16* t1, bf16* t2, bf16* sc, int H, int W, double* stats, bf16** result,
           cudaStream_t s) {
  const long long P = static_cast<long long>(H) * W;
  const int G = v->cfg.norm_num_groups;
  TPDM_TRY(gn(x, t1, P, r.cin, G, r.w.norm1_w, r.w.norm1_b, 1, stats, s));
  TPDM_TRY(conv(t1, H, W, r.cin, r.w.conv1_w, r.w.conv1_b, r.cout, t2, nullptr, s));
  TPDM_TRY(gn(t2, t1, P, r.cout, G, r.w.norm2_w, r.w.norm2_b, 1, stats, s));
  bf16* res = x;
  if (r.cin != r.cout) {
    TPDM_CHECK(r.w.short_w != nullptr, TPDM_ERR_STATE, "vae resnet %d -> %d has no conv_shortcut weights", r.cin, r.cout);
    TPDM_TRY(linear(x, P, r.cin, r.w.short_w, r.w.short_b, r.cout, sc, r.cout, EPI_BIAS_BF16, nullptr, s));
    res = sc;
  }
  TPDM_TRY(conv(t1, H, W, r.cout, r.w.conv2_w, r.w.conv2_b, r.cout, res, res, s));  // in place: out = conv + bias + res
  *result = res;
  return 0;
}

int attention(const tpdm_vae* v, bf16* x, bf16* t1, const Buffers& B, int H, int W, cudaStream_t s) {
  const int C = v->ch[0];
  const long long n = static_cast<long long>(H) * W;
  const tpdm_vae_weights& w = v->w;
  TPDM_CHECK(n % 8 == 0, TPDM_ERR_SHAPE, "vae attention: %lld tokens must be a multiple of 8", n);
  TPDM_TRY(gn(x, t1, n, C, v->cfg.norm_num_groups, w.attn_norm_w, w.attn_norm_b, 0, B.stats, s));
  TPDM_TRY(linear(t1, n, C, w.attn_q_w, w.attn_q_b, C, B.q, C, EPI_BIAS_BF16, nullptr, s));
  TPDM_TRY(linear(t1, n, C, w.attn_k_w, w.attn_k_b, C, B.k, C, EPI_BIAS_BF16, nullptr, s));
  // V^T [C][n] = Wv [C][C] . xn^T ; the value bias is added after P V (softmax rows sum to one)
  TPDM_TRY(linear(reinterpret_cast<const bf16*>(w.attn_v_w), C, C, t1, nullptr, static_cast<int>(n), B.vt, static_cast<int>(n), EPI_BIAS_BF16,
                  nullptr, s));
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(C));
  for (long long r0 = 0; r0 < n; r0 += kAttnChunk) {
    const long long rows = n - r0 < kAttnChunk ? n - r0 : kAttnChunk;
    TPDM_TRY(linear(B.q + r0 * C, rows, C, B.k, nullptr, static_cast<int>(n), B.S, static_cast<int>(n), EPI_BIAS_F32, nullptr, s));
    vae_softmax_rows_kernel<<<static_cast<unsigned>(rows), 256, 0, s>>>(B.S, B.Pm, static_cast<int>(n), scale_log2e);
    count_launch();
    TPDM_TRY(linear(B.Pm, rows, static_cast<int>(n), B.vt, w.attn_v_b, C, B.o + r0 * C, C, EPI_BIAS_BF16, nullptr, s));
  }
  TPDM_TRY(linear(B.o, n, C, w.attn_o_w, w.attn_o_b, C, x, C, EPI_BIAS_ADD_BF16, x, s));  // + residual, in place
  TPDM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" {

int tpdm_vae_create(const tpdm_vae_config* cfg, tpdm_vae** out) {
  TPDM_CHECK(cfg && out, TPDM_ERR_ARG, "tpdm_vae_create: null argument");
  TPDM_CHECK(cfg->num_levels >= 1 && cfg->num_levels <= 8 && cfg->layers_per_block >= 1, TPDM_ERR_ARG, "tpdm_vae_create: bad topology");
  TPDM_CHECK(cfg->latent_channels >= 1 && cfg->latent_channels <= kCinPad && cfg->out_channels >= 1 && cfg->out_channels <= 4, TPDM_ERR_SHAPE,
             "tpdm_vae_create: latent_channels %d (<= 64) / out_channels %d (<= 4) not supported", cfg->latent_channels, cfg->out_channels);
  for (int i = 0; i < cfg->num_levels; ++i)
    TPDM_CHECK(cfg->block_out_channels[i] % 64 == 0 && cfg->block_out_channels[i] % (4 * cfg->norm_num_groups) == 0, TPDM_ERR_SHAPE,
               "tpdm_vae_create: block_out_channels[%d] = %d must be a multiple of 64 and of 4 * norm_num_groups", i,
               cfg->block_out_channels[i]);
  tpdm_vae* v = new (std::nothrow) tpdm_vae();
  TPDM_CHECK(v != nullptr, TPDM_ERR_ARG, "tpdm_vae_create: out of memory");
  v->cfg = *cfg;
  v->levels = cfg->num_levels;
  for (int i = 0; i < v->levels; ++i) v->ch[i] = cfg->block_out_channels[v->levels - 1 - i];
  *out = v;
  return 0;
}

int tpdm_vae_destroy(tpdm_vae* v) {
  delete v;
  return 0;
}

int tpdm_vae_num_resnets(const tpdm_vae* v) { return v ? 2 + v->levels * (v->cfg.layers_per_block + 1) : 0; }

int tpdm_vae_set_weights(tpdm_vae* v, const tpdm_vae_weights* w) {
  TPDM_CHECK(v && w, TPDM_ERR_ARG, "tpdm_vae_set_weights: null argument");
  const int nres = tpdm_vae_num_resnets(v);
  TPDM_CHECK(w->resnets != nullptr && w->n_resnets == nres, TPDM_ERR_ARG, "tpdm_vae_set_weights: expected %d resnets, got %d", nres, w->n_resnets);
  TPDM_CHECK(w->n_upsamplers == v->levels - 1 && (v->levels == 1 || (w->up_conv_w && w->up_conv_b)), TPDM_ERR_ARG,
             "tpdm_vae_set_weights: expected %d upsamplers", v->levels - 1);
  TPDM_CHECK(w->conv_in_w && w->conv_in_b && w->conv_out_w && w->conv_out_b && w->norm_out_w && w->norm_out_b && w->attn_norm_w &&
                 w->attn_norm_b && w->attn_q_w && w->attn_q_b && w->attn_k_w && w->attn_k_b && w->attn_v_w && w->attn_v_b && w->attn_o_w &&
                 w->attn_o_b,
             TPDM_ERR_ARG, "tpdm_vae_set_weights: null weight pointer");
  v->w = *w;
  v->resnets.clear();
  int prev = v->ch[0];
  for (int i = 0; i < nres; ++i) {
    Resnet r;
    r.w = w->resnets[i];
    if (i < 2) {
      r.cin = r.cout = v->ch[0];
    } else {
      const int lvl = (i - 2) / (v->cfg.layers_per_block + 1), j = (i - 2) % (v->cfg.layers_per_block + 1);
      r.cout = v->ch[lvl];
      r.cin = j == 0 ? prev : r.cout;
      if (j == v->cfg.layers_per_block) prev = r.cout;
    }
    TPDM_CHECK(r.w.conv1_w && r.w.conv1_b && r.w.conv2_w && r.w.conv2_b && r.w.norm1_w && r.w.norm1_b && r.w.norm2_w && r.w.norm2_b,
               TPDM_ERR_ARG, "tpdm_vae_set_weights: resnet %d has a null pointer", i);
    TPDM_CHECK(r.cin == r.cout || (r.w.short_w && r.w.short_b), TPDM_ERR_ARG, "tpdm_vae_set_weights: resnet %d (%d -> %d) needs conv_shortcut",
               i, r.cin, r.cout);
    v->resnets.push_back(r);
  }
  v->up_w.assign(w->up_conv_w, w->up_conv_w + w->n_upsamplers);
  v->up_b.assign(w->up_conv_b, w->up_conv_b + w->n_upsamplers);
  v->w.resnets = nullptr;
  v->w.up_conv_w = nullptr;
  v->w.up_conv_b = nullptr;
  v->has_weights = true;
  return 0;
}

size_t tpdm_vae_workspace_bytes(const tpdm_vae* v, int latent_h, int latent_w) {
  if (!v || latent_h <= 0 || latent_w <= 0) return 0;
  return workspace_bytes(v, latent_h, latent_w);
}

int tpdm_vae_decode(tpdm_vae* v, const float* latents, int apply_scaling, int batch, int h, int w, void* workspace,
                    size_t workspace_bytes_given, float* image, unsigned char* rgb, void* stream) {
  TPDM_CHECK(v && latents && workspace && (image || rgb), TPDM_ERR_ARG, "tpdm_vae_decode: null argument");
  TPDM_CHECK(v->has_weights, TPDM_ERR_STATE, "tpdm_vae_decode: weights were not set");
  TPDM_CHECK(batch > 0 && h > 0 && w > 0, TPDM_ERR_SHAPE, "tpdm_vae_decode: empty input");
  TPDM_CHECK((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, TPDM_ERR_ARG, "tpdm_vae_decode: workspace must be 1 KiB aligned");
  TPDM_CHECK(workspace_bytes_given >= workspace_bytes(v, h, w), TPDM_ERR_ARG, "tpdm_vae_decode: workspace too small (%zu < %zu)",
             workspace_bytes_given, workspace_bytes(v, h, w));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  size_t act, rows;
  plan_sizes(v, h, w, &act, &rows);
  const size_t n = static_cast<size_t>(h) * w, C0 = v->ch[0];
  const int ups = v->levels - 1;
  const long long Pout = static_cast<long long>(n) << (2 * ups);
  uint8_t* p = static_cast<uint8_t*>(workspace);
  auto take = [&](size_t bytes) {
    uint8_t* r = p;
    p += align_up(bytes);
    return r;
  };
  Buffers B;
  B.a = reinterpret_cast<bf16*>(take(act));
  B.b = reinterpret_cast<bf16*>(take(act));
  B.c = reinterpret_cast<bf16*>(take(act));
  B.d = reinterpret_cast<bf16*>(take(act));
  B.q = reinterpret_cast<bf16*>(take(n * C0 * 2));
  B.k = reinterpret_cast<bf16*>(take(n * C0 * 2));
  B.vt = reinterpret_cast<bf16*>(take(n * C0 * 2));
  B.o = reinterpret_cast<bf16*>(take(n * C0 * 2));
  B.S = reinterpret_cast<float*>(take(rows * n * 4));
  B.Pm = reinterpret_cast<bf16*>(take(rows * n * 2));
  B.yout = reinterpret_cast<float*>(take(static_cast<size_t>(Pout) * kOutPad * 4));
  B.stats = reinterpret_cast<double*>(take(512 * 2 * 8));
  const int Cl = v->cfg.latent_channels, oc = v->cfg.out_channels, L1 = v->cfg.layers_per_block + 1;

  for (int b = 0; b < batch; ++b) {
    int H = h, W = w;
    // z = latents / scaling + shift (modeling_sd3_pnt.py:653), NHWC bf16, channels padded to 64
    vae_prep_latent_kernel<<<blocks_for(static_cast<long long>(n) * kCinPad, 256, 1 << 30), 256, 0, s>>>(
        latents + static_cast<size_t>(b) * Cl * n, B.d, Cl, static_cast<int>(n), apply_scaling ? 1.0f / v->cfg.scaling_factor : 1.0f,
        apply_scaling ? v->cfg.shift_factor : 0.0f);
    count_launch();
    bf16 *x = B.a, *t1 = B.b, *t2 = B.c, *sc = B.d;
    TPDM_TRY(conv(B.d, H, W, kCinPad, v->w.conv_in_w, v->w.conv_in_b, static_cast<int>(C0), x, nullptr, s));
    bf16* r = nullptr;
    TPDM_TRY(resnet(v, v->resnets[0], x, t1, t2, sc, H, W, B.stats, &r, s));
    TPDM_TRY(attention(v, x, t1, B, H, W, s));
    TPDM_TRY(resnet(v, v->resnets[1], x, t1, t2, sc, H, W, B.stats, &r, s));
    for (int lvl = 0; lvl < v->levels; ++lvl) {
      for (int j = 0; j < L1; ++j) {
        TPDM_TRY(resnet(v, v->resnets[2 + lvl * L1 + j], x, t1, t2, sc, H, W, B.stats, &r, s));
        if (r != x) {  // the width changed: the result lives in the shortcut buffer
          sc = x;
          x = r;
        }
      }
      if (lvl != v->levels - 1) {
        const int C = v->ch[lvl];
        vae_upsample2x_kernel<<<blocks_for(static_cast<long long>(H) * W * (C / 8), 256 * 2, 8 * num_sms()), 256, 0, s>>>(x, t1, H, W, C);
        count_launch();
        H *= 2;
        W *= 2;
        TPDM_TRY(conv(t1, H, W, C, v->up_w[lvl], v->up_b[lvl], C, x, nullptr, s));
      }
    }
    const int Cf = v->ch[v->levels - 1];
    TPDM_TRY(gn(x, t1, static_cast<long long>(H) * W, Cf, v->cfg.norm_num_groups, v->w.norm_out_w, v->w.norm_out_b, 1, B.stats, s));
    {
      GemmOp op;
      TPDM_TRY(gemm_op_init_conv3x3_hw(&op, t1, 1, H, W, Cf, v->w.conv_out_w, kOutPad, EPI_BIAS_F32, B.yout, kOutPad, v->w.conv_out_b, nullptr));
      TPDM_TRY(gemm_launch(&op, 1, s));
    }
    vae_image_out_kernel<<<blocks_for(Pout, 256, 1 << 30), 256, 0, s>>>(B.yout, Pout, oc, image ? image + static_cast<size_t>(b) * oc * Pout : nullptr,
                                                                       rgb ? rgb + static_cast<size_t>(b) * oc * Pout : nullptr);
    count_launch();
  }
  TPDM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
