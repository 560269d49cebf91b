// tpdm_b200 -- training half of the TPDM path: TimePredictor forward with saved
// activations, its backward (weight gradients only: the inputs are constants recorded during the rollout), the PPO-clip
// loss on summed log-probs and a fused clip + AdamW step.
//
// Replaces, for the TimePredictor only (the MMDiT is frozen, modeling_sd3_pnt.py:760-763):
//   only_predict_logprobs               /root/reference/src/models/stable_diffusion_3/modeling_sd3_pnt.py:670-726
//   ratio / PPO-clip loss / backward    /root/reference/src/train/rloo_trainer.py:485-501
//   clip_grad_norm_ + AdamW step        /root/reference/src/train/rloo_trainer.py:505-523
// conv1 (29 GFLOP per sample-step forward, the same again for its weight gradient) runs on tcgen05 tensor cores through
// the GEMM kernel's conv modes; everything else is bandwidth / latency bound.
#include <math.h>
#include <stdlib.h>

#include <new>

#include "common.cuh"
#include "host.h"
#include "kernels.h"

using namespace tpdm;

struct tpdm_tpm_trainer {
  int D, C1, g, max_samples, ns;
  float tpm_eps;
  // parameters (flat fp32, caller owned) and the bf16 copy of conv1 the tensor cores read
  float *params, *grads;
  bf16* conv1_bf16;
  long long off[13];
  // saved activations / scratch
  const bf16* x_nhwc;
  const float* temb;
  bf16 *x_nchw, *dy1t;
  float *y1, *a2, *y2, *emb, *pooled, *u, *ab, *dpooled, *da2, *sums;
  int* amax;
  double* stats;
};

namespace {

enum { P_CONV1_W, P_CONV1_B, P_LIN_W, P_LIN_B, P_GN_W, P_GN_B, P_CONV2_W, P_CONV2_B, P_FC1_W, P_FC1_B, P_FC2_W, P_FC2_B, P_END };

void param_offsets(int D, int C1, long long* off) {
  const long long sizes[P_END] = {static_cast<long long>(C1) * 9 * 2 * D, C1, 2LL * C1 * D, 2LL * C1, C1, C1, 9LL * C1 * C1, C1, 128LL * C1, 128, 256, 2};
  long long o = 0;
  for (int i = 0; i < P_END; ++i) {
    off[i] = o;
    o += (sizes[i] + 3) / 4 * 4;  // keep every tensor 16-byte aligned
  }
  off[P_END] = o;
}

struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<uint8_t*>(b)) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 1023) & ~size_t(1023);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

void carve(tpdm_tpm_trainer* t, Carver& c) {
  const size_t ns = t->max_samples, P = static_cast<size_t>(t->g) * t->g, C1 = t->C1, D = t->D, go = t->g / 2;
  t->x_nchw = c.take<bf16>(ns * 3 * 2 * D * P);  // three x-shifted NCHW copies (conv1 weight-gradient B operand)
  t->dy1t = c.take<bf16>(ns * C1 * P);
  t->y1 = c.take<float>(ns * P * C1);
  t->a2 = c.take<float>(ns * P * C1);
  t->da2 = c.take<float>(ns * P * C1);
  t->y2 = c.take<float>(ns * go * go * C1);
  t->emb = c.take<float>(ns * 2 * C1);
  t->pooled = c.take<float>(ns * C1);
  t->u = c.take<float>(ns * 128);
  t->ab = c.take<float>(ns * 2);
  t->dpooled = c.take<float>(ns * C1);
  t->sums = c.take<float>(ns * 5 * C1);
  t->amax = c.take<int>(ns * C1);
  t->stats = c.take<double>(ns * 2);
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float dsilu_f(float x) {
  const float s = sigmoid_f(x);
  return s * (1.f + x * (1.f - s));
}

// ---- forward tail with saves (modeling_sd3_pnt.py:110-115) ------------------------------------------------------------
__global__ void __launch_bounds__(1024) tpm_tail_train_kernel(const float* __restrict__ y2, int go, int C, const float* __restrict__ fc1_w,
                                                              const float* __restrict__ fc1_b, const float* __restrict__ fc2_w,
                                                              const float* __restrict__ fc2_b, float eps, float* __restrict__ alpha_beta,
                                                              float* __restrict__ pooled_out, int* __restrict__ amax_out,
                                                              float* __restrict__ u_out) {
  __shared__ float part[8][128];
  __shared__ int parti[8][128];
  __shared__ float pooled[128];
  __shared__ float hid[128];
  const int b = blockIdx.x, t = threadIdx.x & 127, grp = threadIdx.x >> 7;
  const float* p = y2 + static_cast<long long>(b) * go * go * C;
  float mx = -INFINITY;
  int arg = 0;
  if (t < C) {
    for (int cell = grp * 32; cell < grp * 32 + 32; ++cell) {
      const int i = cell >> 4, j = cell & 15;
      const int r0 = (i * go) / 16, r1 = ((i + 1) * go + 15) / 16;
      const int c0 = (j * go) / 16, c1 = ((j + 1) * go + 15) / 16;
      float s = 0.f;
      for (int r = r0; r < r1; ++r)
        for (int c = c0; c < c1; ++c) s += p[(static_cast<long long>(r) * go + c) * C + t];
      s /= static_cast<float>((r1 - r0) * (c1 - c0));
      if (s > mx) {  // first maximum wins, as torch's max-pool backward does
        mx = s;
        arg = cell;
      }
    }
  }
  part[grp][t] = mx;
  parti[grp][t] = arg;
  __syncthreads();
  if (threadIdx.x < 128) {
    float m = part[0][t];
    int a = parti[0][t];
    for (int g2 = 1; g2 < 8; ++g2)
      if (part[g2][t] > m) {
        m = part[g2][t];
        a = parti[g2][t];
      }
    pooled[t] = m;
    if (t < C) {
      pooled_out[b * C + t] = m;
      amax_out[b * C + t] = a;
    }
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    float acc = fc1_b[t];
    for (int c = 0; c < C; ++c) acc = fmaf(fc1_w[t * C + c], pooled[c], acc);
    u_out[b * 128 + t] = acc;
    hid[t] = silu_f(acc);
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float acc = fc2_b[t];
    for (int j = 0; j < 128; ++j) acc = fmaf(fc2_w[t * 128 + j], hid[j], acc);
    alpha_beta[b * 2 + t] = expf(acc) + eps;
  }
}

// ---- PPO-clip loss on summed log-probs (rloo_trainer.py:485-495) + d loss / d fc2-output ------------------------------
__device__ double digamma_d(double x) {
  double r = 0.0;
  while (x < 6.0) {
    r -= 1.0 / x;
    x += 1.0;
  }
  const double f = 1.0 / (x * x);
  return r + log(x) - 0.5 / x - f * (1.0 / 12.0 - f * (1.0 / 120.0 - f * (1.0 / 252.0 - f * (1.0 / 240.0 - f / 132.0))));
}

// Beta(alpha, beta).log_prob(r) and its derivatives with respect to the two fc2 outputs z (TimePredictor.forward ends in
// p = exp(z) + eps, modeling_sd3_pnt.py:115, so dp/dz = p - eps).  prediction_type 0: (alpha, beta) = (p1, p2);
// 1 ("mode_concentration", :559-563): alpha = p1 (p2 - 2) + 1, beta = (1 - p1)(p2 - 2) + 1, chain-ruled here so that the
// rollout (schedule_kernel), the replay and the PPO gradient all use the same distribution.
__device__ void beta_logprob_terms(double p1, double p2, double r, int prediction_type, double tpm_eps, double* lp, double* dz0, double* dz1) {
  double A = p1, B = p2;
  if (prediction_type != 0) {
    A = p1 * (p2 - 2.0) + 1.0;
    B = (1.0 - p1) * (p2 - 2.0) + 1.0;
  }
  *lp = (A - 1.0) * log(r) + (B - 1.0) * log1p(-r) + lgamma(A + B) - lgamma(A) - lgamma(B);
  const double psi_ab = digamma_d(A + B);
  const double dA = log(r) + psi_ab - digamma_d(A), dB = log1p(-r) + psi_ab - digamma_d(B);
  if (prediction_type == 0) {
    *dz0 = dA * (p1 - tpm_eps);
    *dz1 = dB * (p2 - tpm_eps);
  } else {
    *dz0 = (dA - dB) * (p2 - 2.0) * (p1 - tpm_eps);
    *dz1 = (dA * p1 + dB * (1.0 - p1)) * (p2 - tpm_eps);
  }
}

// the ratio the replay scores (modeling_sd3_pnt.py:703-712): sigma_next / sigma, or sigma - sigma_next when not relative
__device__ __forceinline__ float replay_ratio(float sigma, float sigma_next, int relative, float eps) {
  const float r = relative ? sigma_next / sigma : sigma - sigma_next;
  return fminf(fmaxf(r, eps), 1.f - eps);
}

struct PpoArgs {
  const float *alpha_beta, *sigmas, *old_logprobs, *advantages;
  float *new_logprobs, *dz, *stats, *reduce_tail;
  int mb, T, relative, prediction_type;
  float min_sigma, eps, cliprange, tpm_eps;
};

__global__ void ppo_clip_kernel(const PpoArgs a) {
  __shared__ float red[4][32];
  const int b = threadIdx.x;
  float loss = 0.f, clipped = 0.f, kl = 0.f, ratio_out = 0.f;
  if (b < a.mb) {
    double sum_new = 0.0, sum_old = 0.0;
    float sigma = 1.0f;
    for (int t = 0; t < a.T; ++t) {
      const int o = b * a.T + t;
      const float sigma_next = a.sigmas[o];
      float lp = 1.0f;  // INVALID_LOGPROB for finished samples (modeling_sd3_pnt.py:721-724)
      if (!(sigma < a.min_sigma)) {
        double l, d0, d1;
        beta_logprob_terms(a.alpha_beta[2 * o], a.alpha_beta[2 * o + 1], replay_ratio(sigma, sigma_next, a.relative, a.eps), a.prediction_type,
                           a.tpm_eps, &l, &d0, &d1);
        lp = static_cast<float>(l);
        a.dz[2 * o] = static_cast<float>(d0);       // scaled by the loss coefficient below
        a.dz[2 * o + 1] = static_cast<float>(d1);
      } else {
        a.dz[2 * o] = 0.f;
        a.dz[2 * o + 1] = 0.f;
      }
      a.new_logprobs[o] = lp;
      sum_new += lp;
      sum_old += a.old_logprobs[o];
      sigma = sigma_next;
    }
    const float diff = static_cast<float>(sum_new - sum_old);
    const float ratio = expf(diff), adv = a.advantages[b];
    const float l1 = -adv * ratio, l2 = -adv * fminf(fmaxf(ratio, 1.f - a.cliprange), 1.f + a.cliprange);
    loss = fmaxf(l1, l2) / a.mb;
    clipped = (l2 > l1 ? 1.f : 0.f) / a.mb;
    kl = 0.5f * diff * diff / a.mb;
    ratio_out = ratio / a.mb;
    // d loss / d sum_new: through -A*ratio when that branch is the max, else through the clamp (zero outside the range)
    const bool inside = ratio >= 1.f - a.cliprange && ratio <= 1.f + a.cliprange;
    const float coef = (l1 >= l2 || inside) ? (-adv * ratio / a.mb) : 0.f;
    for (int t = 0; t < a.T; ++t) {
      const int o = b * a.T + t;
      a.dz[2 * o] *= coef;
      a.dz[2 * o + 1] *= coef;
    }
  }
  float vals[4] = {loss, clipped, kl, ratio_out};
  for (int k = 0; k < 4; ++k) {
    const float s = warp_sum(vals[k]);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) s += red[threadIdx.x][w];
    a.stats[threadIdx.x] = s;
    if (threadIdx.x == 0 && a.reduce_tail) {
      // rides behind the gradients in the flat all-reduce buffer (rloo_trainer.py:497-500: gather(loss), NaN / Inf guard)
      a.reduce_tail[0] = isfinite(s) ? s : 0.f;
      a.reduce_tail[1] = isfinite(s) ? 0.f : 1.f;
    }
  }
}

// only_predict_logprobs (modeling_sd3_pnt.py:670-726) after the TimePredictor: log-prob of the recorded ratio per (sample,
// step), 1.0 at finished steps, and d logprob / d z for the autograd hook of the drop-in module
struct LogprobArgs {
  const float *alpha_beta, *sigmas;
  float *logprobs, *dlp_dz;
  int mb, T, relative, prediction_type;
  float min_sigma, eps, tpm_eps;
};

__global__ void beta_logprob_kernel(const LogprobArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.mb) return;
  float sigma = 1.0f;
  for (int t = 0; t < a.T; ++t) {
    const int o = b * a.T + t;
    const float sigma_next = a.sigmas[o];
    float lp = 1.0f, d0 = 0.f, d1 = 0.f;
    if (!(sigma < a.min_sigma)) {
      double l, g0, g1;
      beta_logprob_terms(a.alpha_beta[2 * o], a.alpha_beta[2 * o + 1], replay_ratio(sigma, sigma_next, a.relative, a.eps), a.prediction_type,
                         a.tpm_eps, &l, &g0, &g1);
      lp = static_cast<float>(l);
      d0 = static_cast<float>(g0);
      d1 = static_cast<float>(g1);
    }
    a.logprobs[o] = lp;
    if (a.dlp_dz) {
      a.dlp_dz[2 * o] = d0;
      a.dlp_dz[2 * o + 1] = d1;
    }
    sigma = sigma_next;
  }
}

// ---- backward of fc2 / SiLU / fc1 / (avg-pool, max) ------------------------------------------------------------------
__global__ void __launch_bounds__(128) tail_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ u, const float* __restrict__ pooled,
                                                       const float* __restrict__ fc1_w, const float* __restrict__ fc2_w, int C,
                                                       float* __restrict__ g_fc1_w, float* __restrict__ g_fc1_b, float* __restrict__ g_fc2_w,
                                                       float* __restrict__ g_fc2_b, float* __restrict__ dpooled) {
  __shared__ float du[128];
  __shared__ float pl[128];
  const int b = blockIdx.x, j = threadIdx.x;
  const float dz0 = dz[2 * b], dz1 = dz[2 * b + 1];
  const float uj = u[b * 128 + j];
  const float hj = silu_f(uj);
  atomicAdd(&g_fc2_w[j], dz0 * hj);
  atomicAdd(&g_fc2_w[128 + j], dz1 * hj);
  if (j < 2) atomicAdd(&g_fc2_b[j], j == 0 ? dz0 : dz1);
  const float duj = (fc2_w[j] * dz0 + fc2_w[128 + j] * dz1) * dsilu_f(uj);
  du[j] = duj;
  pl[j] = j < C ? pooled[b * C + j] : 0.f;
  atomicAdd(&g_fc1_b[j], duj);
  __syncthreads();
  for (int c = 0; c < C; ++c) atomicAdd(&g_fc1_w[j * C + c], duj * pl[c]);
  if (j < C) {
    float acc = 0.f;
    for (int jj = 0; jj < 128; ++jj) acc = fmaf(fc1_w[jj * C + j], du[jj], acc);
    dpooled[b * C + j] = acc;
  }
}

// ---- backward of conv2 (3x3, stride 2): the upstream gradient is non-zero only on each channel's arg-max pooling window --
__global__ void __launch_bounds__(128) conv2_bwd_kernel(const float* __restrict__ a2, const float* __restrict__ w2, const float* __restrict__ dpooled,
                                                        const int* __restrict__ amax, int g, int C, float* __restrict__ g_w2,
                                                        float* __restrict__ g_b2, float* __restrict__ da2) {
  const int b = blockIdx.x, oc = blockIdx.y, cin = threadIdx.x;
  const int go = g / 2;
  const int cell = amax[b * C + oc];
  const int i = cell >> 4, j = cell & 15;
  const int r0 = (i * go) / 16, r1 = ((i + 1) * go + 15) / 16;
  const int c0 = (j * go) / 16, c1 = ((j + 1) * go + 15) / 16;
  const float dy = dpooled[b * C + oc] / static_cast<float>((r1 - r0) * (c1 - c0));
  if (cin == 0) atomicAdd(&g_b2[oc], dy * static_cast<float>((r1 - r0) * (c1 - c0)));
  if (cin >= C) return;
  float gw[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) gw[t] = 0.f;
  const float* ab = a2 + static_cast<long long>(b) * g * g * C;
  float* dab = da2 + static_cast<long long>(b) * g * g * C;
  for (int r = r0; r < r1; ++r)
    for (int c = c0; c < c1; ++c)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int iy = 2 * r + t / 3 - 1, ix = 2 * c + t % 3 - 1;
        if (iy >= 0 && iy < g && ix >= 0 && ix < g) {
          const long long o = (static_cast<long long>(iy) * g + ix) * C + cin;
          gw[t] = fmaf(ab[o], dy, gw[t]);
          atomicAdd(&dab[o], w2[(static_cast<long long>(t) * C + cin) * C + oc] * dy);
        }
      }
#pragma unroll
  for (int t = 0; t < 9; ++t) atomicAdd(&g_w2[(static_cast<long long>(t) * C + cin) * C + oc], gw[t]);
}

// ---- backward of SiLU / adaLN modulation / GroupNorm(1 group) -----------------------------------------------------------
// per (sample, channel) sums over pixels: [0] dv (d shift), [1] dv*n (d scale), [2] dn*xhat (d gamma), [3] dn (d beta), [4] xhat
__global__ void __launch_bounds__(128) gn_bwd_sums_kernel(const float* __restrict__ y1, const float* __restrict__ da2, const double* __restrict__ stats,
                                                          const float* __restrict__ gn_w, const float* __restrict__ gn_b, const float* __restrict__ emb,
                                                          int npix, int C, int pix_per_block, float* __restrict__ sums) {
  const int b = blockIdx.x, c = threadIdx.x;
  if (c >= C) return;
  const double cnt = static_cast<double>(npix) * C;
  const double mean = stats[2 * b] / cnt;
  const double var = stats[2 * b + 1] / cnt - mean * mean;
  const float rstd = rsqrtf(static_cast<float>(var > 0 ? var : 0) + 1e-6f), mu = static_cast<float>(mean);
  const float shift = emb[b * 2 * C + c], scale = emb[b * 2 * C + C + c], gam = gn_w[c], bet = gn_b[c];
  const int p0 = blockIdx.y * pix_per_block, p1 = min(npix, p0 + pix_per_block);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
  for (int p = p0; p < p1; ++p) {
    const long long o = (static_cast<long long>(b) * npix + p) * C + c;
    const float xhat = (y1[o] - mu) * rstd;
    const float n = xhat * gam + bet;
    const float v = n * (1.f + scale) + shift;
    const float dv = da2[o] * dsilu_f(v);
    const float dn = dv * (1.f + scale);
    s0 += dv;
    s1 += dv * n;
    s2 += dn * xhat;
    s3 += dn;
    s4 += xhat;
  }
  float* s = sums + static_cast<long long>(b) * 5 * C;
  atomicAdd(&s[c], s0);
  atomicAdd(&s[C + c], s1);
  atomicAdd(&s[2 * C + c], s2);
  atomicAdd(&s[3 * C + c], s3);
  atomicAdd(&s[4 * C + c], s4);
}

// dy1 = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)); written transposed as bf16 [sample][c][pix] (the K-major A
// operand of the conv1 weight-gradient GEMM); conv1 bias gradient accumulated on the way.
__global__ void __launch_bounds__(128) gn_bwd_apply_kernel(const float* __restrict__ y1, const float* __restrict__ da2, const double* __restrict__ stats,
                                                           const float* __restrict__ gn_w, const float* __restrict__ gn_b, const float* __restrict__ emb,
                                                           const float* __restrict__ sums, int npix, int C, bf16* __restrict__ dy1t,
                                                           float* __restrict__ g_b1) {
  __shared__ float tile[32][129];
  __shared__ float m12[2];
  const int b = blockIdx.x, c = threadIdx.x, p0 = blockIdx.y * 32;
  const double cnt = static_cast<double>(npix) * C;
  const double mean = stats[2 * b] / cnt;
  const double var = stats[2 * b + 1] / cnt - mean * mean;
  const float rstd = rsqrtf(static_cast<float>(var > 0 ? var : 0) + 1e-6f), mu = static_cast<float>(mean);
  const float* s = sums + static_cast<long long>(b) * 5 * C;
  if (threadIdx.x < 32) {  // m1 = mean(dxhat) = sum_c gamma_c * S3_c / cnt ; m2 = mean(dxhat * xhat) = sum_c gamma_c * S2_c / cnt
    float a1 = 0.f, a2s = 0.f;
    for (int cc = threadIdx.x; cc < C; cc += 32) {
      a1 += gn_w[cc] * s[3 * C + cc];
      a2s += gn_w[cc] * s[2 * C + cc];
    }
    a1 = warp_sum(a1);
    a2s = warp_sum(a2s);
    if (threadIdx.x == 0) {
      m12[0] = a1 / static_cast<float>(cnt);
      m12[1] = a2s / static_cast<float>(cnt);
    }
  }
  __syncthreads();
  float bsum = 0.f;
  if (c < C) {
    const float shift = emb[b * 2 * C + c], scale = emb[b * 2 * C + C + c], gam = gn_w[c], bet = gn_b[c];
    for (int pp = 0; pp < 32; ++pp) {
      const long long o = (static_cast<long long>(b) * npix + p0 + pp) * C + c;
      const float xhat = (y1[o] - mu) * rstd;
      const float v = (xhat * gam + bet) * (1.f + scale) + shift;
      const float dxhat = da2[o] * dsilu_f(v) * (1.f + scale) * gam;
      const float d = rstd * (dxhat - m12[0] - xhat * m12[1]);
      tile[pp][c] = d;
      bsum += d;
    }
    atomicAdd(&g_b1[c], bsum);
  }
  __syncthreads();
  // transposed store: 4 threads per channel row segment of 32 pixels (64 bytes)
  for (int idx = threadIdx.x; idx < C * 4; idx += 128) {
    const int cc = idx >> 2, q4 = idx & 3;
    uint4 w;
    w.x = pack_bf16x2(tile[q4 * 8 + 0][cc], tile[q4 * 8 + 1][cc]);
    w.y = pack_bf16x2(tile[q4 * 8 + 2][cc], tile[q4 * 8 + 3][cc]);
    w.z = pack_bf16x2(tile[q4 * 8 + 4][cc], tile[q4 * 8 + 5][cc]);
    w.w = pack_bf16x2(tile[q4 * 8 + 6][cc], tile[q4 * 8 + 7][cc]);
    *reinterpret_cast<uint4*>(dy1t + (static_cast<long long>(b) * C + cc) * npix + p0 + q4 * 8) = w;
  }
}

// gradients of norm1.linear (D -> 2C), norm1.norm affine: reductions over the samples
__global__ void __launch_bounds__(256) lin_grad_kernel(const float* __restrict__ sums, const float* __restrict__ temb, int ns, int C, int D,
                                                       float* __restrict__ g_lin_w, float* __restrict__ g_lin_b, float* __restrict__ g_gn_w,
                                                       float* __restrict__ g_gn_b) {
  const int j = blockIdx.x;  // row of norm1.linear: [0,C) shift, [C,2C) scale
  const int which = j < C ? 0 : 1, c = j < C ? j : j - C;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < ns; ++s) acc = fmaf(sums[(static_cast<long long>(s) * 5 + which) * C + c], silu_f(temb[static_cast<long long>(s) * D + k]), acc);
    g_lin_w[static_cast<long long>(j) * D + k] = acc;
  }
  if (threadIdx.x == 0) {
    float acc = 0.f, gw = 0.f, gb = 0.f;
    for (int s = 0; s < ns; ++s) {
      acc += sums[(static_cast<long long>(s) * 5 + which) * C + c];
      gw += sums[(static_cast<long long>(s) * 5 + 2) * C + c];
      gb += sums[(static_cast<long long>(s) * 5 + 3) * C + c];
    }
    g_lin_b[j] = acc;
    if (which == 0) {
      g_gn_w[c] = gw;
      g_gn_b[c] = gb;
    }
  }
}

// NHWC -> three x-shifted NCHW copies: out[b][k][c][y][x] = x[b][y][x + k - 1][c] (0 outside), k = 0..2
__global__ void nhwc_to_nchw_shifted_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int g, int C) {
  __shared__ bf16 tile[34][34];  // [pixel x (with one halo column each side)][channel]
  const int P = g * g;
  const int b = blockIdx.z, c0 = blockIdx.y * 32;
  const int y = (blockIdx.x * 32) / g, x0 = (blockIdx.x * 32) % g;  // 32 consecutive pixels of one image row (g >= 32) ...
  const int npx = g < 32 ? g : 32;                                  // ... or a whole row when g < 32
  const int yy = g < 32 ? blockIdx.x : y, xx0 = g < 32 ? 0 : x0;
  for (int i = threadIdx.y; i < npx + 2; i += blockDim.y) {
    const int xs = xx0 + i - 1;
    tile[i][threadIdx.x] = (xs >= 0 && xs < g) ? x[(static_cast<long long>(b) * P + yy * g + xs) * C + c0 + threadIdx.x] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int k = 0; k < 3; ++k)
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
      if (threadIdx.x < npx)
        out[((static_cast<long long>(b) * 3 + k) * C + c0 + i) * P + yy * g + xx0 + threadIdx.x] = tile[threadIdx.x + k][i];
}

// ---- clip_grad_norm_ + AdamW (rloo_trainer.py:505-523; torch.optim.AdamW semantics) -----------------------------------
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float scale, double* __restrict__ out) {
  float s = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = g[i] * scale;
    s += v * v;
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0;
    for (int i = 0; i < 8; ++i) a += red[i];
    atomicAdd(out, a);
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                    long long n, float lr, float b1, float b2, float eps, float wd, float max_norm, float grad_scale,
                                                    float bc1, float bc2, const double* __restrict__ sumsq, bf16* __restrict__ bf16_copy,
                                                    long long bf16_n, const float* __restrict__ skip_flag) {
  const float norm = sqrtf(static_cast<float>(*sumsq));
  // NaN / Inf gradient norm, or a non-finite loss on ANY rank (flag summed by the gradient all-reduce): skip the update
  // (rloo_trainer.py:497-500, 518-520)
  const bool finite = isfinite(norm) && (skip_flag == nullptr || *skip_flag == 0.f);
  const float clip = (max_norm > 0.f && norm > max_norm) ? max_norm / (norm + 1e-6f) : 1.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float w = p[i];
    if (finite) {
      const float gi = g[i] * grad_scale * clip;
      const float mi = b1 * m[i] + (1.f - b1) * gi;
      const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      w = w * (1.f - lr * wd) - lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
      p[i] = w;
    }
    if (i < bf16_n) bf16_copy[i] = __float2bfloat16(w);
  }
}

// TPDM_DEBUG_SYNC=1: synchronise after every launch of the backward so a fault is attributed to the right kernel
#define TPDM_DBG_SYNC(stream, what)                                                                              \
  do {                                                                                                           \
    static const bool dbg = getenv("TPDM_DEBUG_SYNC") != nullptr;                                                \
    if (dbg) {                                                                                                   \
      cudaError_t _e = cudaStreamSynchronize(stream);                                                            \
      if (_e != cudaSuccess) return ::tpdm::fail(TPDM_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(_e));   \
    }                                                                                                            \
  } while (0)

int check_trainer(const tpdm_tpm_trainer* t) {
  TPDM_CHECK(t != nullptr, TPDM_ERR_ARG, "null trainer");
  TPDM_CHECK(t->params && t->grads && t->conv1_bf16, TPDM_ERR_STATE, "tpdm_tpm_trainer_bind has not been called");
  return 0;
}

}  // namespace

namespace {

// Reward shaping of one RLOO rollout, one thread per rollout (batch <= 1024), replacing three host-side Python loops:
//   kl[b][t]   = KL(Beta(alpha, beta) || reference Beta) with the reference schedule's Beta(get_ref_beta(sigma_in)) when
//                `relative`, else Beta(1.4, 11.2); 0 at masked steps        (modeling_sd3_pnt.py:875-901,
//                reference_distributions.py:9-19, torch.distributions.kl._kl_beta_beta)
//   score[b]   = mean_{j<=last} last_reward * gamma^(last-j), last = last unmasked step      (modeling_sd3_pnt.py:828-841)
//   rlhf[b]    = score - kl_coef * (sum | mean)_t kl                                           (rloo_trainer.py:447-451)
//   adv[b]     = rlhf - (sum over the rloo_k repeats of the same prompt - rlhf) / (rloo_k - 1)  (rloo_trainer.py:458-461)
// sigma_in[t] = 1 for t = 0, sigmas[t-1] after (F.pad(sigmas[..., :-1], (1, 0), value=1.0)).
struct ShapeArgs {
  const float *alphas, *betas, *sigmas, *last_rewards;
  const int* masks;
  float *kl, *scores, *rlhf, *adv;
  int B, T, relative, ref_steps, mean_kl, rloo_k;
  float gamma, kl_coef;
};

__device__ double log_beta_fn(double a, double b) { return lgamma(a) + lgamma(b) - lgamma(a + b); }

__global__ void rollout_shaping_kernel(ShapeArgs a) {
  __shared__ float rl[1024];
  const int b = threadIdx.x;
  const float ex = 2.718281828459045f;  // reference_distributions.py:7, ex = math.exp(1); python scalar * fp32 tensor -> fp32
  if (b < a.B) {
    double kl_sum = 0.0;
    int last = -1;
    for (int t = 0; t < a.T; ++t) {
      const int i = b * a.T + t;
      double kl = 0.0;
      if (a.masks[i] == 0) {
        last = t;
        const double al = a.alphas[i], be = a.betas[i];
        double ra = 1.4, rb = 11.2;
        if (a.relative) {
          // float arithmetic on purpose: the reference evaluates get_ref_beta on fp32 tensors
          const float s1 = t == 0 ? 1.0f : a.sigmas[i - 1];
          const float t1 = s1 / (ex + (1.0f - ex) * s1);
          const float t2 = fmaxf(t1 - static_cast<float>(1.0 / a.ref_steps), 1e-3f);
          const float s2 = ex / (ex + 1.0f / t2 - 1.0f);
          const float mode = s2 / s1;
          ra = mode * 18.0f + 1.0f;
          rb = (1.0f - mode) * 18.0f + 1.0f;
        }
        const double psi_ab = digamma_d(al + be);
        kl = log_beta_fn(ra, rb) - log_beta_fn(al, be) + (al - ra) * digamma_d(al) + (be - rb) * digamma_d(be) +
             (ra - al + rb - be) * psi_ab;
      }
      if (a.kl) a.kl[i] = static_cast<float>(kl);
      kl_sum += kl;
    }
    float score = 0.f;
    if (a.last_rewards != nullptr && last >= 0) {
      double acc = 0.0, w = 1.0;
      for (int j = 0; j <= last; ++j) {
        acc += w;
        w *= a.gamma;
      }
      score = static_cast<float>(a.last_rewards[b] * acc / (last + 1));
    }
    const float non_score = -a.kl_coef * static_cast<float>(a.mean_kl ? kl_sum / a.T : kl_sum);
    const float r = score + non_score;
    if (a.scores) a.scores[b] = score;
    if (a.rlhf) a.rlhf[b] = r;
    rl[b] = r;
  }
  __syncthreads();
  if (b < a.B && a.adv != nullptr) {
    const int prompts = a.B / a.rloo_k, p = b % prompts;  // layout: rloo_k repeats x prompts
    float sum = 0.f;
    for (int k = 0; k < a.rloo_k; ++k) sum += rl[k * prompts + p];
    a.adv[b] = rl[b] - (sum - rl[b]) / static_cast<float>(a.rloo_k - 1);
  }
}

}  // namespace

extern "C" {

int tpdm_tpm_param_offsets(int D, int C1, long long* out13) {
  TPDM_CHECK(out13 && D > 0 && C1 > 0, TPDM_ERR_ARG, "tpdm_tpm_param_offsets: bad argument");
  param_offsets(D, C1, out13);
  return 0;
}

size_t tpdm_tpm_trainer_workspace_bytes(int D, int C1, int g, int max_samples) {
  tpdm_tpm_trainer t{};
  t.D = D;
  t.C1 = C1;
  t.g = g;
  t.max_samples = max_samples;
  Carver c(nullptr);
  carve(&t, c);
  return c.off + 1024;
}

int tpdm_tpm_trainer_create(int D, int C1, int g, int max_samples, float tpm_epsilon, void* workspace, size_t bytes, tpdm_tpm_trainer** out) {
  TPDM_CHECK(workspace && out, TPDM_ERR_ARG, "tpdm_tpm_trainer_create: null argument");
  TPDM_CHECK(C1 == 128, TPDM_ERR_SHAPE, "trainer: conv_out_channels must be 128 (got %d)", C1);
  TPDM_CHECK(D > 0 && (2 * D) % 256 == 0, TPDM_ERR_SHAPE, "trainer: 2*D=%d must be a multiple of 256", 2 * D);
  TPDM_CHECK(g >= 8 && g <= 128 && (g & (g - 1)) == 0, TPDM_ERR_SHAPE, "trainer: grid side %d must be a power of two in [8,128]", g);
  TPDM_CHECK(max_samples > 0, TPDM_ERR_SHAPE, "trainer: max_samples must be positive");
  TPDM_CHECK((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, TPDM_ERR_ARG, "workspace must be 1024-byte aligned");
  TPDM_CHECK(bytes >= tpdm_tpm_trainer_workspace_bytes(D, C1, g, max_samples), TPDM_ERR_NOMEM, "trainer workspace too small");
  tpdm_tpm_trainer* t = new (std::nothrow) tpdm_tpm_trainer();
  TPDM_CHECK(t, TPDM_ERR_NOMEM, "out of host memory");
  *t = tpdm_tpm_trainer{};
  t->D = D;
  t->C1 = C1;
  t->g = g;
  t->max_samples = max_samples;
  t->tpm_eps = tpm_epsilon;
  param_offsets(D, C1, t->off);
  Carver c(workspace);
  carve(t, c);
  *out = t;
  return 0;
}

int tpdm_tpm_trainer_destroy(tpdm_tpm_trainer* t) {
  delete t;
  return 0;
}

int tpdm_tpm_trainer_bind(tpdm_tpm_trainer* t, float* params, float* grads, void* conv1_w_bf16) {
  TPDM_CHECK(t && params && grads && conv1_w_bf16, TPDM_ERR_ARG, "tpdm_tpm_trainer_bind: null argument");
  t->params = params;
  t->grads = grads;
  t->conv1_bf16 = reinterpret_cast<bf16*>(conv1_w_bf16);
  return 0;
}

int tpdm_tpm_train_forward(tpdm_tpm_trainer* t, const void* x_nhwc, const float* temb, int ns, float* alpha_beta, void* stream) {
  TPDM_TRY(check_trainer(t));
  TPDM_CHECK(x_nhwc && temb && alpha_beta, TPDM_ERR_ARG, "tpdm_tpm_train_forward: null argument");
  TPDM_CHECK(ns > 0 && ns <= t->max_samples, TPDM_ERR_SHAPE, "tpdm_tpm_train_forward: %d samples outside [1,%d]", ns, t->max_samples);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C1 = t->C1, g = t->g, D = t->D;
  const float* P = t->params;
  t->ns = ns;
  t->x_nhwc = reinterpret_cast<const bf16*>(x_nhwc);
  t->temb = temb;
  GemmOp conv;
  TPDM_TRY(gemm_op_init_conv3x3(&conv, x_nhwc, ns, g, 2 * D, t->conv1_bf16, C1, EPI_BIAS_F32, t->y1, C1, P + t->off[P_CONV1_B]));
  TPDM_TRY(gemm_launch(&conv, 1, s));
  TPDM_CUDA_OK(cudaMemsetAsync(t->stats, 0, sizeof(double) * 2 * ns, s));
  TPDM_TRY(k_gn_stats(t->y1, t->stats, ns, static_cast<long long>(g) * g * C1, s));
  TPDM_TRY(k_gemv_f32(P + t->off[P_LIN_W], P + t->off[P_LIN_B], temb, D, nullptr, t->emb, 2 * C1, ns, 2 * C1, D, 1, s));
  TPDM_TRY(k_gn_mod_silu(t->y1, t->stats, P + t->off[P_GN_W], P + t->off[P_GN_B], t->emb, t->a2, ns, g * g, C1, s));
  TPDM_TRY(k_conv3x3_s2(t->a2, P + t->off[P_CONV2_W], P + t->off[P_CONV2_B], t->y2, ns, g, C1, s));
  tpm_tail_train_kernel<<<ns, 1024, 0, s>>>(t->y2, g / 2, C1, P + t->off[P_FC1_W], P + t->off[P_FC1_B], P + t->off[P_FC2_W], P + t->off[P_FC2_B],
                                           t->tpm_eps, t->ab, t->pooled, t->amax, t->u);
  count_launch();
  TPDM_CUDA_OK(cudaGetLastError());
  TPDM_CUDA_OK(cudaMemcpyAsync(alpha_beta, t->ab, sizeof(float) * 2 * ns, cudaMemcpyDeviceToDevice, s));
  return 0;
}

int tpdm_tpm_train_backward(tpdm_tpm_trainer* t, const float* dz, void* stream) {
  TPDM_TRY(check_trainer(t));
  TPDM_CHECK(dz, TPDM_ERR_ARG, "tpdm_tpm_train_backward: null argument");
  TPDM_CHECK(t->ns > 0, TPDM_ERR_STATE, "tpdm_tpm_train_backward: call tpdm_tpm_train_forward first");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C1 = t->C1, g = t->g, D = t->D, ns = t->ns, P = g * g;
  const float* W = t->params;
  float* G = t->grads;
  TPDM_CUDA_OK(cudaMemsetAsync(G, 0, sizeof(float) * t->off[P_END], s));
  TPDM_CUDA_OK(cudaMemsetAsync(t->da2, 0, sizeof(float) * static_cast<size_t>(ns) * P * C1, s));
  TPDM_CUDA_OK(cudaMemsetAsync(t->sums, 0, sizeof(float) * static_cast<size_t>(ns) * 5 * C1, s));
  tail_bwd_kernel<<<ns, 128, 0, s>>>(dz, t->u, t->pooled, W + t->off[P_FC1_W], W + t->off[P_FC2_W], C1, G + t->off[P_FC1_W], G + t->off[P_FC1_B],
                                     G + t->off[P_FC2_W], G + t->off[P_FC2_B], t->dpooled);
  count_launch();
  TPDM_DBG_SYNC(s, "tail_bwd_kernel");
  conv2_bwd_kernel<<<dim3(ns, C1), 128, 0, s>>>(t->a2, W + t->off[P_CONV2_W], t->dpooled, t->amax, g, C1, G + t->off[P_CONV2_W], G + t->off[P_CONV2_B],
                                                t->da2);
  count_launch();
  TPDM_DBG_SYNC(s, "conv2_bwd_kernel");
  const int ppb = 64;
  gn_bwd_sums_kernel<<<dim3(ns, (P + ppb - 1) / ppb), 128, 0, s>>>(t->y1, t->da2, t->stats, W + t->off[P_GN_W], W + t->off[P_GN_B], t->emb, P, C1, ppb,
                                                                  t->sums);
  count_launch();
  TPDM_DBG_SYNC(s, "gn_bwd_sums_kernel");
  gn_bwd_apply_kernel<<<dim3(ns, P / 32), 128, 0, s>>>(t->y1, t->da2, t->stats, W + t->off[P_GN_W], W + t->off[P_GN_B], t->emb, t->sums, P, C1, t->dy1t,
                                                      G + t->off[P_CONV1_B]);
  count_launch();
  TPDM_DBG_SYNC(s, "gn_bwd_apply_kernel");
  lin_grad_kernel<<<2 * C1, 256, 0, s>>>(t->sums, t->temb, ns, C1, D, G + t->off[P_LIN_W], G + t->off[P_LIN_B], G + t->off[P_GN_W], G + t->off[P_GN_B]);
  count_launch();
  TPDM_DBG_SYNC(s, "lin_grad_kernel");
  nhwc_to_nchw_shifted_kernel<<<dim3(g < 32 ? g : P / 32, 2 * D / 32, ns), dim3(32, 8), 0, s>>>(t->x_nhwc, t->x_nchw, g, 2 * D);
  count_launch();
  TPDM_DBG_SYNC(s, "nhwc_to_nchw_shifted_kernel");
  TPDM_CUDA_OK(cudaGetLastError());
  GemmOp wg;
  TPDM_TRY(gemm_op_init_conv3x3_wgrad(&wg, t->dy1t, t->x_nchw, ns, g, 2 * D, C1, G + t->off[P_CONV1_W]));
  TPDM_TRY(gemm_launch(&wg, 1, s));
  TPDM_DBG_SYNC(s, "conv1 wgrad gemm");
  return 0;
}

int tpdm_beta_logprob(const float* alpha_beta, const float* sigmas, int mb, int T, float min_sigma, float epsilon, int relative,
                      int prediction_type, float tpm_epsilon, float* logprobs, float* dlp_dz, void* stream) {
  TPDM_CHECK(alpha_beta && sigmas && logprobs, TPDM_ERR_ARG, "tpdm_beta_logprob: null argument");
  TPDM_CHECK(mb > 0 && T > 0, TPDM_ERR_SHAPE, "tpdm_beta_logprob: empty input");
  TPDM_CHECK(prediction_type == 0 || prediction_type == 1, TPDM_ERR_ARG, "tpdm_beta_logprob: prediction_type %d unknown", prediction_type);
  LogprobArgs a{alpha_beta, sigmas, logprobs, dlp_dz, mb, T, relative, prediction_type, min_sigma, epsilon, tpm_epsilon};
  beta_logprob_kernel<<<(mb + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  TPDM_CUDA_OK(cudaGetLastError());
  return 0;
}

int tpdm_ppo_clip_loss(const float* alpha_beta, const float* sigmas, const float* old_logprobs, const float* advantages, int mb, int T,
                       float min_sigma, float epsilon, int relative, int prediction_type, float cliprange, float tpm_epsilon,
                       float* new_logprobs, float* dz, float* stats4, float* reduce_tail, void* stream) {
  TPDM_CHECK(alpha_beta && sigmas && old_logprobs && advantages && new_logprobs && dz && stats4, TPDM_ERR_ARG, "tpdm_ppo_clip_loss: null argument");
  TPDM_CHECK(mb > 0 && mb <= 1024 && T > 0, TPDM_ERR_SHAPE, "tpdm_ppo_clip_loss: micro-batch %d outside [1,1024]", mb);
  TPDM_CHECK(prediction_type == 0 || prediction_type == 1, TPDM_ERR_ARG, "tpdm_ppo_clip_loss: prediction_type %d unknown", prediction_type);
  PpoArgs a{alpha_beta, sigmas, old_logprobs, advantages, new_logprobs, dz, stats4, reduce_tail, mb, T, relative, prediction_type,
            min_sigma, epsilon, cliprange, tpm_epsilon};
  ppo_clip_kernel<<<1, ((mb + 31) / 32) * 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  TPDM_CUDA_OK(cudaGetLastError());
  return 0;
}

int tpdm_rollout_shaping(const float* alphas, const float* betas, const float* sigmas, const int* masks, const float* last_rewards,
                         int batch, int steps, int relative, int ref_steps, float gamma, float kl_coef, int mean_kl, int rloo_k, float* kl,
                         float* scores, float* rlhf_reward, float* advantages, void* stream) {
  TPDM_CHECK(alphas && betas && sigmas && masks, TPDM_ERR_ARG, "tpdm_rollout_shaping: null argument");
  TPDM_CHECK(batch > 0 && batch <= 1024 && steps > 0, TPDM_ERR_SHAPE, "tpdm_rollout_shaping: batch %d outside [1,1024]", batch);
  TPDM_CHECK(ref_steps > 0, TPDM_ERR_ARG, "tpdm_rollout_shaping: ref_steps must be positive");
  TPDM_CHECK(advantages == nullptr || (rloo_k >= 2 && batch % rloo_k == 0), TPDM_ERR_ARG,
             "tpdm_rollout_shaping: batch %d is not a multiple of rloo_k %d (>= 2)", batch, rloo_k);
  ShapeArgs a{alphas, betas, sigmas, last_rewards, masks, kl, scores, rlhf_reward, advantages, batch, steps, relative, ref_steps,
                    mean_kl, rloo_k, gamma, kl_coef};
  rollout_shaping_kernel<<<1, ((batch + 31) / 32) * 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  TPDM_CUDA_OK(cudaGetLastError());
  return 0;
}

int tpdm_adamw_step(float* params, const float* grads, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                    float weight_decay, float max_grad_norm, int step, float grad_scale, double* scratch_sumsq, void* bf16_copy,
                    long long bf16_n, const float* skip_flag, void* stream) {
  TPDM_CHECK(params && grads && m && v && scratch_sumsq, TPDM_ERR_ARG, "tpdm_adamw_step: null argument");
  TPDM_CHECK(n > 0 && step >= 1, TPDM_ERR_ARG, "tpdm_adamw_step: n and step must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TPDM_CUDA_OK(cudaMemsetAsync(scratch_sumsq, 0, sizeof(double), s));
  sumsq_kernel<<<296, 256, 0, s>>>(grads, n, grad_scale, scratch_sumsq);
  count_launch();
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step)), bc2 = 1.f - powf(beta2, static_cast<float>(step));
  adamw_kernel<<<592, 256, 0, s>>>(params, grads, m, v, n, lr, beta1, beta2, eps, weight_decay, max_grad_norm, grad_scale, bc1, bc2, scratch_sumsq,
                                  reinterpret_cast<bf16*>(bf16_copy), bf16_copy ? bf16_n : 0, skip_flag);
  count_launch();
  TPDM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
