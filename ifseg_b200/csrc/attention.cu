// Fused multi-head attention for head_dim 64 on sm_100a.
//
//   O[b,i,h,:] = head_scale[h] * softmax_j( Q[b,i,h,:].K[b,j,h,:] + bias[h,i,j] + mask ) V[b,j,h,:]
//
// One CTA per (128-query tile, head, batch); 7 warps, two CTAs per SM:
//   warps 0..3 : softmax, thread r owns query row r (TMEM lane r); the 64 keys of a tile are processed as two 32-key
//                sub-tiles of the online softmax (32 scores live at a time; the TMEM load of the second half runs
//                under the arithmetic of the first)
//   warp 4     : score issuer  -- one lane: S = Q K^T and S += bias * I tcgen05.mma, up to two tiles ahead of the softmax
//   warp 5     : output issuer -- one lane: V TMA loads and the O += P V tcgen05.mma
//   warp 6     : loader        -- one lane: Q, then the K ring and the bias ring (TMA), three tiles ahead of the score issuer
// What the r02 timeline and ablation measurements say (profiles/r02_attention_analysis.txt, r02_attention_pair.txt): without
// a bias the tile loop is bound by the softmax threads' own chain (TMEM load round trip, max, exp2, pack, hand-off: a warp
// needs ~1000 clocks per 32-key sub-tile of which the MUFU pipe is busy 256); with a bias the tensor pipe's operand fetch
// (~90 B/clk/SM for M128 N64 K16 MMAs, which re-read their 4 KB A operand every instruction) comes on top.  Hence: issuers
// split so S runs two tiles ahead of the softmax instead of behind P V, P double-buffered so softmax(j+1) never waits for
// P V(j), 3-stage K and bias rings, the additive bias added by the tensor core with the bias tile as the A operand (read
// once), and P handed over through TENSOR MEMORY (no shared-memory stores, no generic->async proxy fence).
// Per 64-key tile j (all asynchronous, mbarrier hand-offs, nothing waits on the tensor core in line):
//   S(j) = Q K(j)^T      tcgen05.mma M128 N64 K64 into one of two TMEM score buffers (score issuer)
//   S(j) += bias(j) I    the TMA-staged fp16 bias tile [128 queries x 64 keys] is the K-major A operand, the fp16 identity
//                        [64 x 64] kept in shared memory the B operand (b * 1.0 is exact, fp32 accumulate): 4 MMAs
//   softmax(j)           tcgen05.ld the row, masks, running max with LAZY rescaling (the accumulator is only touched when
//                        the row max grows by more than 2^8), exp2, row sum; P(j) -> tensor memory as packed bf16 pairs
//                        (tcgen05.st, 16 columns per 32 keys), the layout tcgen05.mma reads as an A operand
//   O += P(j) V(j)       tcgen05.mma M128 N64 K64 accumulating in TMEM, A = P(j) from tensor memory, V tile = MN-major B
// 112 KB of shared memory (Q 16, K 3 x 8, V 2 x 8, identity 8, bias 3 x 16) and 256 TMEM columns (2 x 64 scores, 64 output,
// 2 x 32 probabilities) per CTA keep two CTAs per SM.
#include <stdlib.h>
#include <string.h>

#include <cuda_fp16.h>

#include "common.cuh"

namespace sgf {

static constexpr int kQTile = 128;
static constexpr int kKTile = 64;
static constexpr int kHeadDim = 64;
static constexpr int kAttnThreads = 224;
static constexpr int kSoftmaxWarps = 4;
static constexpr int kKvStages = 2;  // V ring (refilled the moment its reader retires, two tiles ahead)
static constexpr int kKStages = 3;   // K ring and bias ring: refilled two tiles ahead of the score issuer, which itself runs up to
static constexpr int kBStages = 3;   // two tiles ahead of the softmax (a 2-deep ring leaves one tile, ~1 us, for the TMA round trip)
static constexpr float kLog2e = 1.4426950408889634f;
static constexpr float kRescaleThreshold = 8.0f;  // log2 domain: probabilities stay below 2^8

struct AttnParams {
  void* out;
  int64_t o_row_stride, o_batch_stride;
  const void* bias;
  const float* head_scale;
  const uint8_t* kpm;
  int B, H, Tq, Tk, causal;
  float* lse;
  unsigned long long* trace;  // debug: per-role clock64 timeline of a few CTAs (sgf_debug_set_attention_trace)
};

struct AttnSmem {
  static constexpr int kQ = kQTile * kHeadDim * 2;     // 16 KB
  static constexpr int kKV = kKTile * kHeadDim * 2;    // 8 KB
  static constexpr int kIdent = kKTile * kKTile * 2;   // 8 KB fp16 identity [64 x 64]: the B operand of the bias MMAs
  static constexpr int kBias = kQTile * kKTile * 2;    // 16 KB fp16 bias tile: one 128B-swizzled [128 x 64] box
  static constexpr int offQ = 0;
  static constexpr int offK = offQ + kQ;
  static constexpr int offV = offK + kKStages * kKV;
  static constexpr int offIdent = offV + kKvStages * kKV;
  static constexpr int offBias = offIdent + kIdent;  // bias ring
  static constexpr int offBar = offBias + kBStages * kBias;
  static constexpr int kTotal = offBar + 256;
};

struct AttnBars {
  uint64_t q_full, k_full[kKStages], k_empty[kKStages], v_full[kKvStages], v_empty[kKvStages];
  uint64_t s_full[2], s_empty[2], p_full[2], b_full[kBStages], b_empty[kBStages], o_done[2];
  uint32_t tmem_slot;
};
static_assert(sizeof(AttnBars) <= 256, "barrier block");

// debug timeline: slot = (cta, role, iteration, event); 4 CTAs x 4 roles x 32 iterations x 4 events
SGF_DEVICE void attn_trace(const AttnParams& p, int cta_slot, int role, int j, int ev) {
  if (p.trace && cta_slot >= 0 && j < 32)
    p.trace[((cta_slot * 4 + role) * 32 + j) * 4 + ev] = static_cast<unsigned long long>(clock64());
}

__global__ void __launch_bounds__(kAttnThreads, 2) attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                            const __grid_constant__ CUtensorMap tmK,
                                                                            const __grid_constant__ CUtensorMap tmV,
                                                                            const __grid_constant__ CUtensorMap tmB,
                                                                            const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + AttnSmem::offBar);

  pdl_trigger();
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int b = blockIdx.x;  // batch fastest: the CTAs streaming the same (batch-invariant) bias tiles run together
  // causal: the query tiles with the most key tiles are scheduled first (longest processing time first)
  const int q0 = static_cast<int>(p.causal ? gridDim.y - 1 - blockIdx.y : blockIdx.y) * kQTile;
  const int h = blockIdx.z;

  const int lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  const int tslot = !p.trace ? -1 : (lin == 0 ? 0 : (lin == 1 ? 1 : (lin == 300 ? 2 : (lin == 600 ? 3 : -1))));
  int n_kt = (p.Tk + kKTile - 1) / kKTile;
  if (p.causal) n_kt = min(n_kt, (min(q0 + kQTile, p.Tq) + kKTile - 1) / kKTile);  // keys j <= max row of the tile

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();  // the 128B swizzle needs a 1024-byte aligned base
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < kKStages; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], 1);
    }
    for (int i = 0; i < kKvStages; ++i) {
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], 1);
    }
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(&bars->b_full[i], 1);   // bias(j) tile landed (TMA)
      mbar_init(&bars->b_empty[i], 1);  // tcgen05.commit: the bias MMAs of S(j) have read it
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->s_full[i], 1);  // tcgen05.commit of S(j) (scores AND bias MMAs)
      mbar_init(&bars->s_empty[i], kSoftmaxWarps);  // one arrival per softmax warp (lane 0 after __syncwarp)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->p_full[i], kSoftmaxWarps);
      mbar_init(&bars->o_done[i], 1);  // P V(t) retired -> o_done[t & 1]
    }
    fence_mbar_init();
  }
  if (warp == kSoftmaxWarps) {
    if ((tid & 31) == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      if (p.bias) tma_prefetch_desc(&tmB);
    }
    tmem_alloc<256>(&bars->tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  const uint32_t tmem_s = tmem_base;        // score ping-pong: columns [0,64) and [64,128)
  const uint32_t tmem_o = tmem_base + 128;  // output accumulator: columns [128,192)
  const uint32_t tmem_p = tmem_base + 192;  // bf16 probabilities, two buffers of 32 columns (64 keys): columns [192,256)
  if (p.bias) {
    // The additive bias tile is added to the scores BY THE TENSOR CORE: S(j) += bias(j) * I, with the TMA-staged fp16 bias
    // tile [128 queries x 64 keys] as the K-major A operand and the fp16 identity [64 x 64] kept in shared memory as the
    // B operand -- exact (b * 1.0, fp32 accumulate), four K = 16 steps.  (r02a: the softmax threads added it: 4 x LDS.128 + 32
    // conversions + 16 packed adds per thread and tile, 10 us of an 85 us launch.  r02b: identity [128 x 128] in TMEM as A
    // and the bias tile as the MN-major B: eight K steps that re-read the identity for every key tile, 48 KB of operand
    // fetch per tile -- with Q K^T and P V the tensor pipe's operand fetch, ~90 B/clk, paced the kernel,
    // profiles/r02_attention_pair.txt.)  Swizzled K-major rows of 128 B: row n holds 1.0 at element n.
    for (int ci = tid; ci < AttnSmem::kIdent / 16; ci += kAttnThreads) {
      const int n = ci >> 3;                        // row
      const int logical = (ci & 7) ^ (n & 7);       // 16-byte chunk of the row stored at position ci & 7
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      if (logical == (n >> 3)) w[(n & 7) >> 1] = (n & 1) ? 0x3C000000u : 0x00003C00u;
      *reinterpret_cast<uint4*>(smem + AttnSmem::offIdent + ci * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
    __syncthreads();
  }
  pdl_wait();

  if (warp == kSoftmaxWarps) {
    // ============ score issuer: Q / K / bias loads and S(j) = Q K(j)^T, up to two tiles ahead of the softmax ============
    // (r01 issued S and P V from ONE thread in program order S(j), P V(j-1): S(j) then waited for softmax(j-2) to FINISH
    //  although it only needs softmax(j-2) to have READ its score buffer, and the ~1700-cycle round trip
    //  p_full -> wake -> issue -> retire -> s_full sat on the critical path of every tile: ncu showed the softmax warps
    //  parked on s_full while the issuing thread was parked on p_full.)
    if ((tid & 31) == 0 && n_kt > 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(kQTile, kKTile, 0, 0);
      constexpr uint32_t idesc_bias = make_idesc_f16(kQTile, kKTile, 0, 0);  // A = bias tile, B = identity, both K-major
      const uint64_t di = make_smem_desc_sw128(smem_u32(smem + AttnSmem::offIdent));
      const uint64_t dq = make_smem_desc_sw128(smem_u32(smem + AttnSmem::offQ));
#pragma unroll 1
      for (int j = 0; j < n_kt; ++j) {
        // ---- S(j) into score buffer j&1 (free once softmax(j-2) has read it) ----
        const int st = j % kKStages;
        attn_trace(p, tslot, 0, j, 0);
        if (j == 0) mbar_wait(&bars->q_full, 0);
        mbar_wait(&bars->k_full[st], (j / kKStages) & 1);
        attn_trace(p, tslot, 0, j, 1);
        mbar_wait(&bars->s_empty[j & 1], ((j >> 1) & 1) ^ 1);
        attn_trace(p, tslot, 0, j, 2);
        tc_fence_after();
        const uint64_t dk = make_smem_desc_sw128(smem_u32(smem + AttnSmem::offK + st * AttnSmem::kKV));
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          umma_f16(tmem_s + (j & 1) * 64, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        // with a bias the score buffer is completed by S(j) += bias(j) * I (four more MMAs over the 64 keys)
        if (p.bias) {
          const int sb = j % kBStages;
          mbar_wait(&bars->b_full[sb], (j / kBStages) & 1);
          tc_fence_after();
          const uint64_t db = make_smem_desc_sw128(smem_u32(smem + AttnSmem::offBias + sb * AttnSmem::kBias));
#pragma unroll
          for (int k = 0; k < kKTile / 16; ++k) umma_f16(tmem_s + (j & 1) * 64, db + 2 * k, di + 2 * k, idesc_bias, 1u);
          umma_commit(&bars->b_empty[sb]);
        }
        umma_commit(&bars->s_full[j & 1]);
        umma_commit(&bars->k_empty[st]);
        attn_trace(p, tslot, 0, j, 3);
      }
    }
  } else if (warp == kSoftmaxWarps + 1) {
    // ============ output issuer: V loads and O += P(t) V(t) as soon as softmax(t) has published P ============
    if ((tid & 31) == 0 && n_kt > 0) {
      constexpr uint32_t idesc_pv = make_idesc_bf16(kQTile, kHeadDim, 0, 1);  // B (= V) is MN-major
      auto load_v = [&](int t) {
        const int st = t % kKvStages;
        mbar_expect_tx(&bars->v_full[st], AttnSmem::kKV);
        tma_load_4d(smem + AttnSmem::offV + st * AttnSmem::kKV, &tmV, &bars->v_full[st], 0, h, t * kKTile, b);
      };
      for (int t = 0; t < kKvStages && t < n_kt; ++t) load_v(t);
#pragma unroll 1
      for (int t = 0; t < n_kt; ++t) {
        const int st = t % kKvStages;
        attn_trace(p, tslot, 1, t, 0);
        mbar_wait(&bars->v_full[st], (t / kKvStages) & 1);
        attn_trace(p, tslot, 1, t, 1);
        mbar_wait(&bars->p_full[t & 1], (t >> 1) & 1);
        attn_trace(p, tslot, 1, t, 2);
        tc_fence_after();
        const uint32_t tp = tmem_p + (t & 1) * 32;  // P(t): bf16 pairs, 8 columns per 16-key K step
        const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + AttnSmem::offV + st * AttnSmem::kKV));
#pragma unroll
        for (int k = 0; k < kKTile / 16; ++k)  // V: 16 key rows = 2048 B per K step -> +128 in the address field
          umma_f16_ts(tmem_o, tp + 8 * k, dv + 128 * k, idesc_pv, (t | k) != 0 ? 1u : 0u);
        umma_commit(&bars->v_empty[st]);
        umma_commit(&bars->o_done[t & 1]);
        attn_trace(p, tslot, 1, t, 3);
        if (t + kKvStages < n_kt) {  // V(t+2) goes into the stage P V(t) is reading
          mbar_wait(&bars->v_empty[st], (t / kKvStages) & 1);
          load_v(t + kKvStages);
        }
      }
    }
  } else if (warp == kSoftmaxWarps + 2) {
    // ============ loader: Q, then the K and bias rings, each stage refilled the moment its reader S(j-3) retires ============
    // (the score issuer used to do this itself: five mbarrier round trips, two TMA issues, eight MMAs and three commits per
    //  tile from ONE thread took ~2200 clocks and starved the softmax warps by 600-800 clocks per tile)
    if ((tid & 31) == 0 && n_kt > 0) {
      auto load_k = [&](int t) {
        const int st = t % kKStages;
        mbar_expect_tx(&bars->k_full[st], AttnSmem::kKV);
        tma_load_4d(smem + AttnSmem::offK + st * AttnSmem::kKV, &tmK, &bars->k_full[st], 0, h, t * kKTile, b);
      };
      auto load_bias = [&](int t) {
        const int st = t % kBStages;
        mbar_expect_tx(&bars->b_full[st], AttnSmem::kBias);
        tma_load_3d(smem + AttnSmem::offBias + st * AttnSmem::kBias, &tmB, &bars->b_full[st], t * kKTile, q0, h);
      };
      static_assert(kKStages == kBStages, "one loop refills both rings");
      mbar_expect_tx(&bars->q_full, AttnSmem::kQ);
      tma_load_4d(smem + AttnSmem::offQ, &tmQ, &bars->q_full, 0, h, q0, b);
#pragma unroll 1
      for (int t = 0; t < n_kt; ++t) {
        if (t >= kKStages) mbar_wait(&bars->k_empty[t % kKStages], ((t / kKStages) - 1) & 1);
        load_k(t);
        if (p.bias) {
          if (t >= kBStages) mbar_wait(&bars->b_empty[t % kBStages], ((t / kBStages) - 1) & 1);
          load_bias(t);
        }
      }
    }
  } else {
    // =================================== softmax warps ===================================
    // thread = query row (TMEM lane); the 64 keys of a tile are taken as two 32-key sub-tiles of the online softmax so
    // that only 32 scores are live at a time (~100 registers instead of the 164 of a 64-wide row) and the TMEM load of
    // the second half overlaps the arithmetic of the first
    const int lane = tid & 31;
    const int rowl = warp * 32 + lane;
    const int row = q0 + rowl;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const uint8_t* kpm_row = p.kpm ? p.kpm + static_cast<int64_t>(b) * p.Tk : nullptr;
    float m_used = 0.f;  // log2-domain reference max the probabilities are expressed against
    float l_run = 0.f;
    const int trole = (lane == 0 && (warp == 0 || warp == 3)) ? (warp == 0 ? 2 : 3) : -1;

#pragma unroll 1
    for (int j = 0; j < n_kt; ++j) {
      const uint32_t tp = tmem_p + lane_addr + (j & 1) * 32;  // this row of P(j): 16 columns per 32-key half
      if (trole >= 0) attn_trace(p, tslot, trole, j, 0);
      mbar_wait(&bars->s_full[j & 1], (j >> 1) & 1);  // S(j) = Q K(j)^T + bias(j) retired
      tc_fence_after();
      if (trole >= 0) attn_trace(p, tslot, trole, j, 1);
      // keys of this tile a row may attend to: [0, lim) (tail of the sequence, causal diagonal)
      int lim = p.Tk - j * kKTile;
      if (p.causal) lim = min(lim, row - j * kKTile + 1);
      const bool need_mask = lim < kKTile || kpm_row != nullptr;  // (rows differ under the causal mask: per thread)
      uint32_t acc[32];
      tmem_ld_32x32(tmem_s + (j & 1) * 64 + lane_addr, acc);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float s[32];
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(acc[i]);
        if (hf == 0) {
          tmem_ld_32x32(tmem_s + (j & 1) * 64 + 32 + lane_addr, acc);  // second half: in flight under the math below
        } else {  // both halves of S(j) are in registers: hand the score buffer back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->s_empty[j & 1]);
        }
        if (need_mask) {
          const int l2 = lim - hf * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= l2) s[i] = -INFINITY;
          if (kpm_row) {
            const int c0 = j * kKTile + hf * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i < p.Tk && kpm_row[c0 + i] != 0) s[i] = -INFINITY;
          }
        }
        // 4 independent partial maxima (a single 32-deep fmax chain would serialise on FP latency)
        float mxp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) mxp[i] = s[i];
#pragma unroll
        for (int i = 4; i < 32; ++i) mxp[i & 3] = fmaxf(mxp[i & 3], s[i]);
        const float m_new = fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3])) * kLog2e;
        if (j == 0 && hf == 0) {
          m_used = (m_new == -INFINITY) ? 0.f : m_new;
        } else {
          // lazy rescale: only when some row of the warp outgrew its reference max by more than 2^8
          const bool grow = m_new > m_used + kRescaleThreshold;
          if (__any_sync(0xffffffffu, grow)) {
            const float f = grow ? fast_exp2(m_used - m_new) : 1.0f;
            if (j >= 1) {
              mbar_wait(&bars->o_done[(j - 1) & 1], ((j - 1) >> 1) & 1);  // every P V issued so far has retired
              tc_fence_after();
#pragma unroll
              for (int hb = 0; hb < 4; ++hb) {
                uint32_t r[16];
                tmem_ld_32x16(tmem_o + lane_addr + hb * 16, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
                tmem_st_32x16(tmem_o + lane_addr + hb * 16, r);
              }
              tmem_st_wait();
              tc_fence_before();
            }
            if (hf == 1) {  // the first half of P(j) is already in tensor memory against the old reference
              uint32_t w[16];
              tmem_ld_32x16(tp, w);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const float2 v = unpack_bf16x2(w[q]);
                w[q] = pack_bf16x2(v.x * f, v.y * f);
              }
              tmem_st_32x16(tp, w);
            }
            if (grow) {
              l_run *= f;
              m_used = m_new;
            }
          }
        }
        if (trole >= 0 && hf == 0) attn_trace(p, tslot, trole, j, 2);
        float2 ps[4] = {splat2(0.f), splat2(0.f), splat2(0.f), splat2(0.f)};
        const float2 l2e = splat2(kLog2e), nm = splat2(-m_used);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 e = fma2(make_float2(s[i], s[i + 1]), l2e, nm);
          s[i] = fast_exp2(e.x);
          s[i + 1] = fast_exp2(e.y);
          ps[(i >> 1) & 3] = add2(ps[(i >> 1) & 3], make_float2(s[i], s[i + 1]));
        }
        const float2 pt = add2(add2(ps[0], ps[1]), add2(ps[2], ps[3]));
        l_run += pt.x + pt.y;
        if (trole >= 0 && hf == 0) attn_trace(p, tslot, trole, j, 3);
        // P (bf16) -> buffer j&1 once its previous reader P V(j-2) has retired (issued two tiles ago)
        if (hf == 0 && j >= 2) mbar_wait(&bars->o_done[j & 1], ((j - 2) >> 1) & 1);
        {
          uint32_t w[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) w[q] = pack_bf16x2(s[2 * q], s[2 * q + 1]);
          tmem_st_32x16(tp + hf * 16, w);
        }
      }
      tmem_st_wait();
      tc_fence_before();  // tensor-memory writes of P(j) -> ordered before the arrive the issuing thread waits on
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[j & 1]);
    }

    if (n_kt > 0) {
      mbar_wait(&bars->o_done[(n_kt - 1) & 1], ((n_kt - 1) >> 1) & 1);
      tc_fence_after();
    }
    {
      // the TMEM loads are warp-collective: every lane executes them, rows >= Tq only skip the stores
      const float inv = (1.0f / l_run) * (p.head_scale ? p.head_scale[h] : 1.0f);
      if (p.lse && row < p.Tq)  // log2-domain log-sum-exp of the (biased, masked) score row, for the backward
        p.lse[(static_cast<int64_t>(b) * p.H + h) * p.Tq + row] = m_used + __log2f(l_run);
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<int64_t>(b) * p.o_batch_stride +
                           static_cast<int64_t>(row) * p.o_row_stride + h * kHeadDim;
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_o + lane_addr + hb * 32, r);
        tmem_ld_wait();
        if (row < p.Tq) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(r[8 * c + 0]) * inv, __uint_as_float(r[8 * c + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(r[8 * c + 2]) * inv, __uint_as_float(r[8 * c + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(r[8 * c + 4]) * inv, __uint_as_float(r[8 * c + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(r[8 * c + 6]) * inv, __uint_as_float(r[8 * c + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + hb * 32 + 8 * c) = u;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kSoftmaxWarps) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

static int make_qkv_map(CUtensorMap* m, const void* base, int64_t row_stride, int64_t batch_stride, int H, int T, int B,
                        int box_rows) {
  uint64_t dims[4] = {kHeadDim, static_cast<uint64_t>(H), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
  uint64_t strides[3] = {kHeadDim * 2, static_cast<uint64_t>(row_stride) * 2, static_cast<uint64_t>(batch_stride) * 2};
  uint32_t box[4] = {kHeadDim, 1, static_cast<uint32_t>(box_rows), 1};
  return encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace sgf

using namespace sgf;

static unsigned long long* g_attention_trace = nullptr;
extern "C" void sgf_debug_set_attention_trace(void* buf) { g_attention_trace = reinterpret_cast<unsigned long long*>(buf); }

extern "C" int sgf_attention_bf16(const sgf_attention_args* a, void* stream) {
  SGF_REQUIRE(a != nullptr && a->q && a->k && a->v && a->out, "attention: null pointer");
  SGF_REQUIRE(a->B > 0 && a->H > 0 && a->Tq > 0 && a->Tk > 0, "attention: bad shape");
  SGF_REQUIRE(a->q_row_stride % 8 == 0 && a->k_row_stride % 8 == 0 && a->v_row_stride % 8 == 0 &&
                  a->q_batch_stride % 8 == 0 && a->k_batch_stride % 8 == 0 && a->v_batch_stride % 8 == 0 &&
                  a->o_row_stride % 8 == 0 && a->o_batch_stride % 8 == 0,
              "attention: strides must be multiples of 8 elements");
  SGF_REQUIRE((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) |
               reinterpret_cast<uintptr_t>(a->v) | reinterpret_cast<uintptr_t>(a->out)) % 16 == 0,
              "attention: q/k/v/out must be 16-byte aligned");
  if (a->bias)
    SGF_REQUIRE(a->bias_row_stride % 8 == 0 && a->bias_head_stride % 8 == 0 &&
                    reinterpret_cast<uintptr_t>(a->bias) % 16 == 0 && a->bias_row_stride >= a->Tk,
                "attention: the fp16 bias must be 16-byte aligned with row/head strides multiples of 8 elements");
  CUtensorMap tmQ, tmK, tmV;
  if (int rc = make_qkv_map(&tmQ, a->q, a->q_row_stride, a->q_batch_stride, a->H, a->Tq, a->B, kQTile)) return rc;
  if (int rc = make_qkv_map(&tmK, a->k, a->k_row_stride, a->k_batch_stride, a->H, a->Tk, a->B, kKTile)) return rc;
  if (int rc = make_qkv_map(&tmV, a->v, a->v_row_stride, a->v_batch_stride, a->H, a->Tk, a->B, kKTile)) return rc;
  CUtensorMap tmB;
  memset(&tmB, 0, sizeof(tmB));
  if (a->bias) {
    uint64_t dims[3] = {static_cast<uint64_t>(a->bias_row_stride), static_cast<uint64_t>(a->Tq), static_cast<uint64_t>(a->H)};
    uint64_t strides[2] = {static_cast<uint64_t>(a->bias_row_stride) * 2, static_cast<uint64_t>(a->bias_head_stride) * 2};
    uint32_t box[3] = {kKTile, kQTile, 1};
    if (int rc = encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, a->bias, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  AttnParams p{a->out, a->o_row_stride, a->o_batch_stride, a->bias, a->head_scale, a->key_padding_mask,
               a->B, a->H, a->Tq, a->Tk, a->causal, a->lse, g_attention_trace};
  constexpr int smem = AttnSmem::kTotal;
  static bool configured = false;
  if (!configured) {
    SGF_CHECK_CUDA(
        cudaFuncSetAttribute(attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(a->B, (a->Tq + kQTile - 1) / kQTile, a->H);
  SGF_CHECK_CUDA(launch_pdl(attention_tcgen05_kernel, grid, dim3(kAttnThreads), smem,
                            reinterpret_cast<cudaStream_t>(stream), tmQ, tmK, tmV, tmB, p));
  count_launch();
  return SGF_OK;
}
