// tpdm_b200 -- joint text+image attention, softmax(Q K^T / sqrt(d)) V, on tcgen05 tensor cores (sm_100a).
//
// Replaces F.scaled_dot_product_attention inside diffusers' JointAttnProcessor2_0 as called from
// JointTransformerBlock (/root/reference/src/models/stable_diffusion_3/transformer_sd3.py:361-365): image tokens first,
// text tokens second, no mask, no dropout.  Q/K/V are read straight out of the token-major fused-QKV GEMM output
// [Bt][S][3*H*dp] through 4-D TMA tensor maps (no head permute, no concat copy); O is written token-major [Bt][S][H*dp]
// so the output projection consumes it as-is.
//
// One CTA = one 128-row query tile of one (batch, head); 256 threads; two CTAs are co-resident per SM (dp=64) so that
// one CTA's softmax overlaps the other's MMAs:
//   warp 0      TMA producer  (Q once, K/V tiles of 128 keys through a 2-stage ring)
//   warp 1      MMA issuer    S = Q K^T (SS, 128x128xdp) -> TMEM;  O += P V (A = P from TMEM, B = V MN-major from smem)
//   warp 2      TMEM allocator
//   warps 4-7   softmax: one thread per query row, fp32, exp2 with the log2(e)/sqrt(d) scale folded in, online max with
//               lazy rescale (O in TMEM is only touched when the running max grows by more than 2^8), P written back to
//               TMEM as packed bf16; final 1/l normalisation and bf16 store.
#include <cuda_bf16.h>

#include "common.cuh"
#include "host.h"

namespace tpdm {

namespace {

constexpr int kAttnThreads = 256;
constexpr int kQT = 128;   // query rows per CTA
constexpr int kKT = 128;   // keys per KV tile
constexpr int kKVStages = 2;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

template <int DP>
struct AttnSmem {
  static constexpr int kTile = kQT * DP * 2;  // bytes of one Q / K / V tile
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kTile;
  static constexpr int kVOff = kKOff + kKVStages * kTile;
  static constexpr int kBarOff = kVOff + kKVStages * kTile;
  static constexpr int kTotal = kBarOff + 256 + 1024;
  static constexpr uint32_t kTmemCols = DP == 64 ? 256 : 512;
  static constexpr uint32_t kSCol = 0, kPCol = 128, kOCol = 192;
};

template <int DP>
__global__ void __launch_bounds__(kAttnThreads, DP == 64 ? 2 : 1) joint_attention_tcgen05_kernel(const __grid_constant__ AttnOp A) {
  using L = AttnSmem<DP>;
  if (A.skip != nullptr && *A.skip != 0) return;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;                 // [kKVStages]
  uint64_t* v_full = k_full + kKVStages;       // [kKVStages]
  uint64_t* k_empty = v_full + kKVStages;      // [kKVStages]
  uint64_t* v_empty = k_empty + kKVStages;     // [kKVStages]
  uint64_t* s_full = v_empty + kKVStages;
  uint64_t* s_empty = s_full + 1;
  uint64_t* p_full = s_empty + 1;
  uint64_t* pv_done = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * kQT;
  const int n_kv = (A.S + kKT - 1) / kKT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&A.tmQ);
    tma_prefetch_desc(&A.tmK);
    tma_prefetch_desc(&A.tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kKVStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<L::kTmemCols>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, L::kTile);
#pragma unroll
      for (int hh = 0; hh < DP / 64; ++hh) tma_load_4d(smem + L::kQOff + hh * (kQT * 128), &A.tmQ, q_full, hh * 64, h, q0, b);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&k_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&k_full[stage], L::kTile);
#pragma unroll
        for (int hh = 0; hh < DP / 64; ++hh)
          tma_load_4d(smem + L::kKOff + stage * L::kTile + hh * (kKT * 128), &A.tmK, &k_full[stage], hh * 64, h, j * kKT, b);
        mbar_wait(&v_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&v_full[stage], L::kTile);
#pragma unroll
        for (int hh = 0; hh < DP / 64; ++hh)
          tma_load_4d(smem + L::kVOff + stage * L::kTile + hh * (kKT * 128), &A.tmV, &v_full[stage], hh * 64, h, j * kKT, b);
        if (++stage == kKVStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_qk = make_idesc_bf16(kQT, kKT, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(kQT, DP, true);
    const uint32_t s_tmem = tmem_base + L::kSCol, p_tmem = tmem_base + L::kPCol, o_tmem = tmem_base + L::kOCol;
    const uint32_t q_base = smem_u32(smem + L::kQOff);

    auto issue_qk = [&](int stage) {
      const uint32_t k_base = smem_u32(smem + L::kKOff + stage * L::kTile);
#pragma unroll
      for (int ks = 0; ks < DP / 16; ++ks) {
        const uint32_t off = (ks / 4) * (kQT * 128) + (ks % 4) * 32;
        umma_ss(s_tmem, make_smem_desc_sw128(q_base + off, 16, 1024), make_smem_desc_sw128(k_base + off, 16, 1024), idesc_qk,
                ks != 0 ? 1u : 0u);
      }
    };

    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    if (lane == 0) {
      issue_qk(0);
      umma_commit(&k_empty[0]);
      umma_commit(s_full);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < n_kv; ++j) {
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == kKVStages) {
        nstage = 0;
        nphase ^= 1;
      }
      if (j + 1 < n_kv) {
        mbar_wait(&k_full[nstage], nphase);
        mbar_wait(s_empty, j & 1);  // softmax has read S_j
        tc_fence_after();
        if (lane == 0) {
          issue_qk(nstage);
          umma_commit(&k_empty[nstage]);
          umma_commit(s_full);
        }
        __syncwarp();
      }
      mbar_wait(p_full, j & 1);  // P_j in TMEM, O corrected
      mbar_wait(&v_full[stage], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t v_base = smem_u32(smem + L::kVOff + stage * L::kTile);
#pragma unroll
        for (int ks = 0; ks < kKT / 16; ++ks) {
          // B = V tile, MN-major: 16 keys per MMA = 16 rows of 128 B; 64-wide d-groups are kKT*128 B apart (LBO)
          umma_ts(o_tmem, p_tmem + ks * 8, make_smem_desc_sw128(v_base + ks * 2048, kKT * 128, 1024), idesc_pv,
                  (j | ks) != 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[stage]);
        umma_commit(pv_done);
      }
      __syncwarp();
      stage = nstage;
      phase = nphase;
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax / correction / epilogue
    const int q = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_off + L::kSCol;
    const uint32_t p_tmem = tmem_base + lane_off + L::kPCol;
    const uint32_t o_tmem = tmem_base + lane_off + L::kOCol;
    const float scale = A.scale_log2;
    float m_used = -INFINITY;  // running (possibly stale) max, log2 domain
    float l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int valid = A.S - j * kKT;  // keys valid in this tile (>=1)
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row max
      float m_tile = -INFINITY;
#pragma unroll
      for (int c = 0; c < kKT / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(s_tmem + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(v[i]);
          if (c * 32 + i >= valid) s = -INFINITY;
          m_tile = fmaxf(m_tile, s);
        }
      }
      m_tile *= scale;
      float alpha = 1.f;
      bool need = false;
      if (j == 0) {
        m_used = m_tile;
      } else if (m_tile > m_used + kRescaleThreshold) {
        need = true;
        alpha = exp2_approx(m_used - m_tile);
        m_used = m_tile;
      }
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // PV_{j-1} finished: P and O may be touched
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
#pragma unroll
          for (int c = 0; c < DP / 32; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(o_tmem + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(o_tmem + c * 32, o);
          }
          l *= alpha;
        }
      }
      // pass 2: p = exp2(s*scale - m), row sum, bf16 P -> TMEM
#pragma unroll
      for (int c = 0; c < kKT / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(s_tmem + c * 32, v);
        tmem_wait_ld();
        if (c == kKT / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty);
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = exp2_approx(fmaf(__uint_as_float(v[i]), scale, -m_used));
          float p1 = exp2_approx(fmaf(__uint_as_float(v[i + 1]), scale, -m_used));
          if (c * 32 + i >= valid) p0 = 0.f;
          if (c * 32 + i + 1 >= valid) p1 = 0.f;
          l += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x16(p_tmem + c * 16, pk);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: O / l -> bf16, token-major
    mbar_wait(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    const int row = q0 + q * 32 + lane;
    const float inv_l = 1.0f / l;
    __nv_bfloat16* out = A.out + (static_cast<long long>(b) * A.S + row) * (static_cast<long long>(A.H) * DP) + h * DP;
#pragma unroll
    for (int c = 0; c < DP / 32; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(o_tmem + c * 32, o);
      tmem_wait_ld();
      if (row < A.S) {
        uint4* dst = reinterpret_cast<uint4*>(out + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
          w.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
          w.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
          w.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
          dst[i] = w;
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<L::kTmemCols>(tmem_base);
}

template <int DP>
int attn_launch_impl(const AttnOp& op, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    TPDM_CUDA_OK(cudaFuncSetAttribute(joint_attention_tcgen05_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      AttnSmem<DP>::kTotal));
    attr_set = true;
  }
  dim3 grid(op.q_tiles, op.H, op.Bt);
  const double q_rows = op.q_tiles * kQT < op.S ? op.q_tiles * kQT : op.S;
  prof_begin(1, 4.0 * op.Bt * op.H * q_rows * op.S * op.head_dim, stream);
  joint_attention_tcgen05_kernel<DP><<<grid, kAttnThreads, AttnSmem<DP>::kTotal, stream>>>(op);
  prof_end(stream);
  count_launch();
  TPDM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

int attn_op_init(AttnOp* op, const void* qkv, int Bt, int S, int H, int dp, int head_dim, void* out) {
  *op = AttnOp{};
  TPDM_CHECK(dp == 64 || dp == 128, TPDM_ERR_SHAPE, "attention: padded head dim %d must be 64 or 128", dp);
  TPDM_CHECK(head_dim > 0 && head_dim <= dp, TPDM_ERR_SHAPE, "attention: head_dim %d exceeds padded %d", head_dim, dp);
  TPDM_CHECK(S > 0 && H > 0 && Bt > 0, TPDM_ERR_SHAPE, "attention: empty problem");
  op->S = S;
  op->H = H;
  op->Bt = Bt;
  op->dp = dp;
  op->head_dim = head_dim;
  op->q_tiles = (S + kQT - 1) / kQT;
  op->scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(head_dim));
  op->out = reinterpret_cast<__nv_bfloat16*>(out);
  const uint64_t row = static_cast<uint64_t>(3) * H * dp;  // elements per token in the fused qkv buffer
  uint64_t dims[4] = {static_cast<uint64_t>(dp), static_cast<uint64_t>(H), static_cast<uint64_t>(S), static_cast<uint64_t>(Bt)};
  uint64_t strides[3] = {static_cast<uint64_t>(dp) * 2, row * 2, row * S * 2};
  uint32_t box[4] = {64, 1, kQT, 1};
  const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(qkv);
  TPDM_TRY(encode_tmap_bf16(&op->tmQ, base, 4, dims, strides, box));
  TPDM_TRY(encode_tmap_bf16(&op->tmK, base + static_cast<size_t>(H) * dp, 4, dims, strides, box));
  TPDM_TRY(encode_tmap_bf16(&op->tmV, base + static_cast<size_t>(2) * H * dp, 4, dims, strides, box));
  return 0;
}

int attn_launch(const AttnOp* op_in, cudaStream_t stream) {
  AttnOp op_copy = *op_in;
  op_copy.skip = skip_flag();
  const AttnOp* op = &op_copy;
  return op->dp == 64 ? attn_launch_impl<64>(*op, stream) : attn_launch_impl<128>(*op, stream);
}

}  // namespace tpdm
