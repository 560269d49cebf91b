// tpdm_b200 -- joint text+image attention, softmax(Q K^T / sqrt(d)) V, on tcgen05 tensor cores (sm_100a).
//
// Replaces F.scaled_dot_product_attention inside diffusers' JointAttnProcessor2_0 as called from
// JointTransformerBlock (/root/reference/src/models/stable_diffusion_3/transformer_sd3.py:361-365): image tokens first,
// text tokens second, no mask, no dropout.  Q/K/V are read straight out of the token-major fused-QKV GEMM output
// [Bt][S][3*H*dp] through 4-D TMA tensor maps (no head permute, no concat copy); O is written token-major [Bt][S][H*dp]
// so the output projection consumes it as-is.
//
// One CTA = one 128-row query tile of one (batch, head); 256 threads; two CTAs co-resident per SM (dp = 64).
//   warp 0      TMA producer  (Q once, K/V tiles of 128 keys through a 2-stage smem ring)
//   warp 1      MMA issuer for S.  Keys are consumed in HALF-tiles of 64:  S[b] = Q K_half^T (SS, 128x64xdp), b = parity
//   warp 3      MMA issuer for O += P[b] V_half (A = P from TMEM, B = V MN-major from smem)
//   warp 2      TMEM allocator
//   warps 4-7   softmax, one thread per query row.
// S and P are DOUBLE-BUFFERED in TMEM (S[2] 64 fp32 columns each, P[2] 32 packed-bf16 columns each, O dp columns), so the
// softmax warps never wait for the tensor core in steady state: while they work on half-tile i the MMA warps have already
// produced S(i+1) and are accumulating P(i-1) V.
// Softmax: fp32, exp2 with log2(e)/sqrt(d) folded into one FFMA2, ONE pass per half-tile against a reference max that is
// fixed by the first half-tile and not tracked afterwards: fp32 sums and bf16 P carry 8 exponent bits, so the result stays
// exact (the final 1/l cancels the reference) until a row grows past 2^128 relative to it.  That is detected from the row
// sum before P is published; such a half-tile, the first one and the masked tail take an exact two-pass route that moves
// the reference and rescales O and l once the outstanding P V has drained.
//
// Where the time goes (ncu source-level samples, profiles/r01_attention_ncu.txt): per 64-key half-tile and CTA the kernel
// needs 512 cycles of the MUFU pipe (ex2, 16 / clk / SM) AND 512 cycles of the TMEM read port (S in fp32, 64 B / clk / SM)
// -- both floors are 214 us for S = 4429, H = 24, Bt = 2 -- against 256 cycles of tensor pipe; it runs at 61 % of either.
// Measured and rejected: tracking the half-tile's own max (32 FMNMX3 per row and half-tile; removing it changed nothing,
// so it was dropped), 8 softmax warps with the columns of a half-tile split between two warps per lane quarter (360 us
// stand-alone, 3 % slower in the trajectory), prefetching the next half-tile's first 32 S columns during the second
// exponential block (398 us: 128 registers and spills), a second MMA-issuing warp for P V (kept, +2 %).  With the
// exponentials compiled out the kernel still takes 260 us, so MUFU and the TMEM-read side are about equally loaded.
// Not tried: fp16 accumulators for S (half the TMEM read) -- the error it adds grows with the logit magnitude and cannot
// be checked against real checkpoints offline (random-init weights give near-uniform attention).
#include <cuda_bf16.h>

#include "common.cuh"
#include "host.h"

namespace tpdm {

namespace {

constexpr int kAttnThreads = 256;
constexpr int kQT = 128;   // query rows per CTA
constexpr int kKT = 128;   // keys per K/V smem tile
constexpr int kHT = 64;    // keys per half-tile (one S / P buffer)
constexpr int kKVStages = 2;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
// Every kPolyEvery-th pair of exponentials on the fast path is computed on the FMA/ALU pipes (Cody-Waite range
// reduction + cubic minimax polynomial, max relative error 7.5e-5 -- P is rounded to bf16 anyway) instead of MUFU.EX2,
// meant to relieve the 16/clk/SM MUFU rate.  MEASURED on B200 (S=4429, H=24, Bt=2): 0 -> 359 us, 1/8 -> 376 us,
// 1/4 -> 400 us, 1/2 -> 418 us: the softmax warps are issue/latency bound, so the extra instructions cost more than the
// MUFU slots they free.  Left in for head dims / occupancies where MUFU does bind; 0 disables the emulation.
#ifndef TPDM_POLY_EVERY
#define TPDM_POLY_EVERY 0
#endif
constexpr int kPolyEvery = TPDM_POLY_EVERY;

// -DTPDM_ATTN_TRACE: clock64() time stamps of one CTA's softmax warp 4 (role 0, 8 slots per half-tile) and of the two MMA-issuing
// warps (roles 1 and 2, 4 slots per half-tile), read back with tpdm_attn_trace_read (tools/attn_trace.py).  Diagnostic builds only.
#ifdef TPDM_ATTN_TRACE
__device__ long long g_attn_trace[3][2048];
#define ATRACE(role, idx)                                                                       \
  do {                                                                                          \
    if (trace_on && lane == 0 && (idx) < 2048) g_attn_trace[role][idx] = clock64();             \
  } while (0)
#else
#define ATRACE(role, idx)
#endif

template <int DP>
struct AttnSmem {
  static constexpr int kTile = kQT * DP * 2;  // bytes of one Q / K / V tile
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kTile;
  static constexpr int kVOff = kKOff + kKVStages * kTile;
  static constexpr int kBarOff = kVOff + kKVStages * kTile;
  static constexpr int kTotal = kBarOff + 256 + 1024;
  static constexpr uint32_t kTmemCols = DP == 64 ? 256 : 512;
  static constexpr uint32_t kSCol = 0;     // S[b] at kSCol + 64 b
  static constexpr uint32_t kPCol = 128;   // P[b] at kPCol + 32 b
  static constexpr uint32_t kOCol = 192;   // O: DP columns
};

template <int DP>
__global__ void __launch_bounds__(kAttnThreads, DP == 64 ? 2 : 1) joint_attention_tcgen05_kernel(const __grid_constant__ AttnOp A) {
  using L = AttnSmem<DP>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;                 // [kKVStages]
  uint64_t* v_full = k_full + kKVStages;       // [kKVStages]
  uint64_t* k_empty = v_full + kKVStages;      // [kKVStages]
  uint64_t* v_empty = k_empty + kKVStages;     // [kKVStages]
  uint64_t* s_full = v_empty + kKVStages;      // [2]  MMA -> softmax : S[b] written
  uint64_t* s_free = s_full + 2;               // [2]  softmax -> MMA : S[b] read into registers
  uint64_t* p_full = s_free + 2;               // [2]  softmax -> MMA : P[b] written (and O rescaled if needed)
  uint64_t* pv_done = p_full + 2;              // [2]  MMA -> softmax : P[b] V accumulated (P[b] may be overwritten)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * kQT;
  const int n_kv = (A.S + kKT - 1) / kKT;
  const int n_half = (A.S + kHT - 1) / kHT;
#ifdef TPDM_ATTN_TRACE
  const bool trace_on = blockIdx.x == 5 && blockIdx.y == 3 && blockIdx.z == 0;
#endif

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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  pdl_wait();
  if (A.skip != nullptr && *A.skip != 0) return;
  if (A.bmask != nullptr && A.bmask[blockIdx.z % A.bslots] == 0) return;  // emptied queue slot
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
  } else if (warp == 1 || warp == 3) {
    // ---------------------------------------------------------------- MMA issuers (warp 1: S = Q K^T, warp 3: O += P V)
    constexpr uint32_t idesc_qk = make_idesc_bf16(kQT, kHT, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(kQT, DP, true);
    const uint32_t q_base = smem_u32(smem + L::kQOff);

    // half-tile i lives in K/V smem tile t = i/2 (ring stage t % 2, ring phase (t/2) & 1), rows [64 (i&1), +64)
    auto issue_qk = [&](int i) {
      const int t = i >> 1, hf = i & 1, stage = t % kKVStages;
      ATRACE(1, 4 * i);
      if (hf == 0) mbar_wait(&k_full[stage], (t / kKVStages) & 1);
      if (i >= 2) mbar_wait(&s_free[i & 1], ((i - 2) >> 1) & 1);  // softmax(i-2) has read S[i&1]
      tc_fence_after();
      ATRACE(1, 4 * i + 1);
      if (lane == 0) {
        const uint32_t k_base = smem_u32(smem + L::kKOff + stage * L::kTile) + hf * (kHT * 128);
        const uint32_t s_tmem = tmem_base + L::kSCol + (i & 1) * kHT;
#pragma unroll
        for (int ks = 0; ks < DP / 16; ++ks) {
          const uint32_t off = (ks / 4) * (kQT * 128) + (ks % 4) * 32;
          umma_ss(s_tmem, make_smem_desc_sw128(q_base + off, 16, 1024), make_smem_desc_sw128(k_base + off, 16, 1024), idesc_qk,
                  ks != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[i & 1]);
        if (hf == 1 || i == n_half - 1) umma_commit(&k_empty[stage]);
      }
      ATRACE(1, 4 * i + 2);
      __syncwarp();
    };
    auto issue_pv = [&](int i) {
      const int t = i >> 1, hf = i & 1, stage = t % kKVStages;
      ATRACE(2, 4 * i);
      mbar_wait(&p_full[i & 1], (i >> 1) & 1);
      if (hf == 0) mbar_wait(&v_full[stage], (t / kKVStages) & 1);
      tc_fence_after();
      ATRACE(2, 4 * i + 1);
      if (lane == 0) {
        const uint32_t v_base = smem_u32(smem + L::kVOff + stage * L::kTile) + hf * (kHT * 128);
        const uint32_t p_tmem = tmem_base + L::kPCol + (i & 1) * (kHT / 2);
        const uint32_t o_tmem = tmem_base + L::kOCol;
#pragma unroll
        for (int ks = 0; ks < kHT / 16; ++ks) {
          // B = V half-tile, MN-major: 16 keys per MMA = 16 rows of 128 B; 64-wide d-groups are kKT*128 B apart (LBO)
          umma_ts(o_tmem, p_tmem + ks * 8, make_smem_desc_sw128(v_base + ks * 2048, kKT * 128, 1024), idesc_pv,
                  (i | ks) != 0 ? 1u : 0u);
        }
        umma_commit(&pv_done[i & 1]);
        if (hf == 1 || i == n_half - 1) umma_commit(&v_empty[stage]);
      }
      ATRACE(2, 4 * i + 2);
      __syncwarp();
    };

    // QK and PV are issued by TWO warps.  Every tcgen05.mma costs the issuing thread ~150 cycles before the next one can
    // go out, however small it is, and with one issuer the 16 MMAs per 128 keys (8 QK + 8 PV) -- not the softmax -- paced
    // this kernel (every measured variant fits that model, see the list at the top).  All hand-offs between the two
    // streams go through mbarriers (s_free / p_full / pv_done), none relies on a common issue order.
    if (warp == 1) {
      mbar_wait(q_full, 0);
      for (int i = 0; i < n_half; ++i) issue_qk(i);
    } else {
      for (int i = 0; i < n_half; ++i) issue_pv(i);
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax / correction / epilogue
    const int q = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t o_tmem = tmem_base + lane_off + L::kOCol;
    const float scale = A.scale_log2;
    const uint64_t scale2 = pack_f32x2(scale, scale);
    float m_used = -INFINITY;            // reference max of the running sums (log2 domain); may lag the true max
    uint64_t l2 = pack_f32x2(0.f, 0.f);  // row sum as two partial sums
    float alpha_pending = 1.f;           // rescale discovered in half-tile i-1, applied before P(i) is published
    bool need_pending = false;

    auto rescale_o = [&](float alpha) {
#pragma unroll
      for (int c = 0; c < DP / 32; ++c) {
        uint32_t o[32];
        tmem_ld_32x32(o_tmem + c * 32, o);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_32x32(o_tmem + c * 32, o);
      }
      l2 = fmul2(l2, pack_f32x2(alpha, alpha));
    };
    // p = exp2(s * scale - m) for one 32-column chunk, packed to bf16 into P; accumulates the row sum
    auto exp_chunk = [&](const uint32_t (&v)[32], uint32_t p_dst, uint64_t negm2, const bool emulate) {
      uint32_t pk[16];
      const uint64_t magic2 = pack_f32x2(12582912.f, 12582912.f), nmagic2 = pack_f32x2(-12582912.f, -12582912.f);
      const uint64_t mone2 = pack_f32x2(-1.f, -1.f);
      const uint64_t c0 = pack_f32x2(0.9999280572f, 0.9999280572f), c1 = pack_f32x2(0.6932609677f, 0.6932609677f),
                     c2 = pack_f32x2(0.2426111251f, 0.2426111251f), c3 = pack_f32x2(0.0551716499f, 0.0551716499f);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float x0, x1, p0, p1;
        unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), scale2, negm2), x0, x1);
        if (kPolyEvery > 0 && emulate && (i % (kPolyEvery > 0 ? kPolyEvery : 1)) == (kPolyEvery - 1)) {
          // 2^x = 2^n * 2^f, n = rint(x) read from the mantissa of x + 1.5*2^23, f = x - n in [-0.5, 0.5]
          const uint64_t xp = pack_f32x2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
          const uint64_t t = fadd2(xp, magic2);
          const uint64_t fr = ffma2(fadd2(t, nmagic2), mone2, xp);
          uint64_t pp = ffma2(c3, fr, c2);
          pp = ffma2(pp, fr, c1);
          pp = ffma2(pp, fr, c0);
          float t0, t1;
          unpack_f32x2(t, t0, t1);
          unpack_f32x2(pp, p0, p1);
          p0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
          p1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
        } else {
          p0 = exp2_approx(x0);
          p1 = exp2_approx(x1);
        }
        l2 = fadd2(l2, pack_f32x2(p0, p1));
        pk[i] = pack_bf16x2(p0, p1);
      }
      tmem_st_32x16(p_dst, pk);
    };
    auto chunk_max = [&](const uint32_t (&v)[32], float m) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) m = fmax3(m, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
      return m;
    };

    for (int i = 0; i < n_half; ++i) {
      const int bb = i & 1;
      const uint32_t s_tmem = tmem_base + lane_off + L::kSCol + bb * kHT;
      const uint32_t p_tmem = tmem_base + lane_off + L::kPCol + bb * (kHT / 2);
      const int valid = A.S - i * kHT;  // keys valid in this half-tile (>= 1)
#ifdef TPDM_ATTN_TRACE
      const bool tr4 = trace_on && warp == 4;
#define STRACE(e) do { if (tr4 && lane == 0 && 8 * i + (e) < 2048) g_attn_trace[0][8 * i + (e)] = clock64(); } while (0)
#else
#define STRACE(e)
#endif
      STRACE(0);
      mbar_wait(&s_full[bb], (i >> 1) & 1);
      if (i >= 2) mbar_wait(&pv_done[bb], ((i - 2) >> 1) & 1);  // P(i-2) V done: P[bb] may be overwritten
      tc_fence_after();
      STRACE(1);
      if (__any_sync(0xffffffffu, need_pending)) {
        // every P V issued so far must have drained before O is touched (P(i-1) V is the newest; the MMAs retire in order)
        mbar_wait(&pv_done[bb ^ 1], ((i - 1) >> 1) & 1);
        tc_fence_after();
        rescale_o(alpha_pending);
      }
      need_pending = false;
      alpha_pending = 1.f;
      bool fast_ok = false;
      if (i > 0 && valid >= kHT) {
        // ---- fast path: one pass against the reference max m_used, which is NOT updated here.  fp32 sums and bf16 P carry an
        // 8-bit exponent, so p = 2^(s - m_used) stays exact in relative terms however far the true row max has moved past
        // m_used -- until 2^128.  Overflow is detected after the fact (an infinite row sum) BEFORE P is published; S[bb] is
        // released only after that check, so the exact route below can redo the half-tile and move m_used.  This drops the
        // 32 FMNMX3 and their dependent chain per half-tile that tracking the tile's own max cost.
        const uint64_t l_save = l2;
        const uint64_t negm2 = pack_f32x2(-m_used, -m_used);
        uint32_t va[32], vb[32];
        tmem_ld_32x32(s_tmem, va);
        tmem_wait_ld();
        STRACE(2);
        tmem_ld_32x32(s_tmem + 32, vb);
        exp_chunk(va, p_tmem, negm2, true);
        STRACE(3);
        tmem_wait_ld();
        STRACE(4);
        exp_chunk(vb, p_tmem + 16, negm2, true);
        float l_lo, l_hi;
        unpack_f32x2(l2, l_lo, l_hi);
        const bool overflow = !(l_lo + l_hi < 1e37f);  // inf, NaN, or close enough to the bf16 / fp32 limit to round to inf
        if (!__any_sync(0xffffffffu, overflow)) {
          fast_ok = true;
          STRACE(5);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[bb]);
          STRACE(6);
        } else {
          l2 = l_save;  // S[bb] is still intact (s_free not signalled): redo on the exact route
        }
      }
      if (!fast_ok) {
        // ---- exact two-pass route: first half-tile, masked tail, or a jump of the row max above 2^100
        float m_tile = -INFINITY;
#pragma unroll
        for (int c = 0; c < kHT / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(s_tmem + c * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            float sv = __uint_as_float(v[e]);
            if (c * 32 + e >= valid) sv = -INFINITY;
            m_tile = fmaxf(m_tile, sv);
          }
        }
        m_tile *= scale;
        if (i == 0) {
          m_used = m_tile;
        } else {
          const bool need = m_tile > m_used + kRescaleThreshold;
          const float alpha = need ? exp2_approx(m_used - m_tile) : 1.f;
          if (need) m_used = m_tile;
          if (__any_sync(0xffffffffu, need)) {
            mbar_wait(&pv_done[bb ^ 1], ((i - 1) >> 1) & 1);
            tc_fence_after();
            rescale_o(alpha);
          }
        }
        const uint64_t negm2 = pack_f32x2(-m_used, -m_used);
#pragma unroll
        for (int c = 0; c < kHT / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(s_tmem + c * 32, v);
          tmem_wait_ld();
          if (c == kHT / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[bb]);
          }
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c * 32 + e >= valid) v[e] = 0xff800000u;  // -inf -> p = 0
          exp_chunk(v, p_tmem + c * 16, negm2, false);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[bb]);
      STRACE(7);
    }
    float l_lo, l_hi;
    unpack_f32x2(l2, l_lo, l_hi);
    const float l = l_lo + l_hi;
    // epilogue: O / l -> bf16, token-major
    mbar_wait(&pv_done[(n_half - 1) & 1], ((n_half - 1) >> 1) & 1);
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
        for (int e = 0; e < 4; ++e) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[8 * e + 0]) * inv_l, __uint_as_float(o[8 * e + 1]) * inv_l);
          w.y = pack_bf16x2(__uint_as_float(o[8 * e + 2]) * inv_l, __uint_as_float(o[8 * e + 3]) * inv_l);
          w.z = pack_bf16x2(__uint_as_float(o[8 * e + 4]) * inv_l, __uint_as_float(o[8 * e + 5]) * inv_l);
          w.w = pack_bf16x2(__uint_as_float(o[8 * e + 6]) * inv_l, __uint_as_float(o[8 * e + 7]) * inv_l);
          dst[e] = w;
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
  TPDM_CUDA_OK(launch_pdl(joint_attention_tcgen05_kernel<DP>, grid, dim3(kAttnThreads), AttnSmem<DP>::kTotal, stream, op));
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

#ifdef TPDM_ATTN_TRACE
extern "C" int tpdm_attn_trace_read(long long* host, int n) {
  return cudaMemcpyFromSymbol(host, g_attn_trace, sizeof(long long) * (n < 3 * 2048 ? n : 3 * 2048)) == cudaSuccess ? 0 : -1;
}
#endif

int attn_launch(const AttnOp* op_in, cudaStream_t stream) {
  AttnOp op_copy = *op_in;
  op_copy.skip = skip_flag();
  op_copy.bmask = batch_mask();
  op_copy.bslots = batch_mask_slots();
  const AttnOp* op = &op_copy;
  return op->dp == 64 ? attn_launch_impl<64>(*op, stream) : attn_launch_impl<128>(*op, stream);
}

}  // namespace tpdm
