// Fused multi-head attention for head_dim 64 on sm_100a.
//
//   O[b,i,h,:] = head_scale[h] * softmax_j( Q[b,i,h,:].K[b,j,h,:] + bias[h,i,j] + mask ) V[b,j,h,:]
//
// One CTA per (128-query tile, head, batch), 128 threads; thread r owns query row r.
// Per 64-key tile:
//   S = Q K^T       tcgen05.mma  M128 N64 K64   (Q, K tiles K-major, 128B swizzle, by TMA)  -> TMEM[0,64)
//   softmax         thread r: tcgen05.ld its row, add bias, mask, online max/sum in fp32 (exp2)
//   P -> smem       bf16, written directly in the 128B-swizzled K-major layout UMMA expects
//   T = P V         tcgen05.mma  M128 N64 K64   (V tile is the MN-major B operand)          -> TMEM[64,128)
//   O = O*alpha + T thread r: tcgen05.ld its row of T, rescale-accumulate in registers
// K/V tiles are double-buffered by TMA; 64 KB of shared memory and 128 TMEM columns per CTA let
// three CTAs share an SM so one CTA's MMAs overlap another's softmax.
#include <string.h>

#include "common.cuh"

namespace sgf {

static constexpr int kQTile = 128;
static constexpr int kKTile = 64;
static constexpr int kHeadDim = 64;
static constexpr int kAttnThreads = 128;
static constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
  void* out;
  int64_t o_row_stride, o_batch_stride;
  const float* bias;
  int64_t bias_head_stride, bias_row_stride;
  const float* head_scale;
  const uint8_t* kpm;
  int B, H, Tq, Tk, causal;
};

struct AttnSmem {
  static constexpr int kQ = kQTile * kHeadDim * 2;   // 16 KB
  static constexpr int kKV = kKTile * kHeadDim * 2;  // 8 KB
  static constexpr int kP = kQTile * kKTile * 2;     // 16 KB
  static constexpr int offQ = 0;
  static constexpr int offK = offQ + kQ;       // 2 stages
  static constexpr int offV = offK + 2 * kKV;  // 2 stages
  static constexpr int offP = offV + 2 * kKV;
  static constexpr int kBias = kQTile * kKTile * 4;  // 32 KB fp32 bias tile: two 128B-swizzled [128 x 32] boxes
  static constexpr int offBias = offP + kP;
  static constexpr int offBar = offBias + kBias;
  static constexpr int kTotal = offBar + 64;
};

__global__ void __launch_bounds__(kAttnThreads) attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                         const __grid_constant__ CUtensorMap tmK,
                                                                         const __grid_constant__ CUtensorMap tmV,
                                                                         const __grid_constant__ CUtensorMap tmB,
                                                                         const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + AttnSmem::offBar);
  uint64_t* bar_kv = bar_q + 1;  // [2]
  uint64_t* bar_s = bar_q + 3;  // [2]  S ping-pong
  uint64_t* bar_o = bar_q + 5;
  uint64_t* bar_b = bar_q + 6;  // bias tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 7);

  pdl_trigger();
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int q0 = blockIdx.x * kQTile;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int row = q0 + tid;

  int n_kt = (p.Tk + kKTile - 1) / kKTile;
  if (p.causal) {
    const int lim = (min(q0 + kQTile, p.Tq) + kKTile - 1) / kKTile;  // keys j <= max row
    n_kt = min(n_kt, lim);
  }

  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_q, 1);
    mbar_init(&bar_kv[0], 1);
    mbar_init(&bar_kv[1], 1);
    mbar_init(&bar_s[0], 1);
    mbar_init(&bar_s[1], 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_b, 1);
    fence_mbar_init();
    if (p.bias) tma_prefetch_desc(&tmB);
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;        // S ping-pong: columns [0,64) and [64,128)
  const uint32_t tmem_t = tmem_base + 128;  // T = P V: columns [128,192)
  pdl_wait();

  if (tid == 0) {
    mbar_expect_tx(bar_q, AttnSmem::kQ);
    tma_load_4d(smem + AttnSmem::offQ, &tmQ, bar_q, 0, h, q0, b);
    for (int st = 0; st < 2 && st < n_kt; ++st) {
      mbar_expect_tx(&bar_kv[st], 2 * AttnSmem::kKV);
      tma_load_4d(smem + AttnSmem::offK + st * AttnSmem::kKV, &tmK, &bar_kv[st], 0, h, st * kKTile, b);
      tma_load_4d(smem + AttnSmem::offV + st * AttnSmem::kKV, &tmV, &bar_kv[st], 0, h, st * kKTile, b);
    }
    if (p.bias && n_kt > 0) {
      mbar_expect_tx(bar_b, AttnSmem::kBias);
      tma_load_3d(smem + AttnSmem::offBias, &tmB, bar_b, 0, q0, h);
      tma_load_3d(smem + AttnSmem::offBias + AttnSmem::kBias / 2, &tmB, bar_b, 32, q0, h);
    }
  }

  constexpr uint32_t idesc_qk = make_idesc_bf16(kQTile, kKTile, 0, 0);
  constexpr uint32_t idesc_pv = make_idesc_bf16(kQTile, kHeadDim, 0, 1);  // B (=V) is MN-major
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;

  float o[kHeadDim];
#pragma unroll
  for (int j = 0; j < kHeadDim; ++j) o[j] = 0.f;
  float m_run = -INFINITY;  // running max (log2 domain)
  float l_run = 0.f;

  const uint8_t* bias_row = smem + AttnSmem::offBias + tid * 128;  // this thread's row inside each bias box
  const uint8_t* kpm_row = p.kpm ? p.kpm + static_cast<int64_t>(b) * p.Tk : nullptr;
  uint8_t* p_row = smem + AttnSmem::offP + tid * 128;

  for (int kt = 0; kt < n_kt; ++kt) {
    const int st = kt & 1;
    const uint32_t kv_phase = (kt >> 1) & 1;
    const int k0 = kt * kKTile;

    // S(kt) was issued one iteration ahead (S(0) below); issue S(kt+1) now so that the tensor core
    // works on the next score tile while this tile's softmax runs on the CUDA cores.
    if (tid == 0) {
      const uint64_t dq = make_smem_desc_sw128(smem_u32(smem + AttnSmem::offQ));
      if (kt == 0) {
        mbar_wait(bar_q, 0);
        mbar_wait(&bar_kv[0], 0);
        tc_fence_after();
        const uint64_t dk = make_smem_desc_sw128(smem_u32(smem + AttnSmem::offK));
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) umma_f16(tmem_s, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(&bar_s[0]);
      }
      if (kt + 1 < n_kt) {
        const int sn = (kt + 1) & 1;
        mbar_wait(&bar_kv[sn], ((kt + 1) >> 1) & 1);
        tc_fence_after();
        const uint64_t dk = make_smem_desc_sw128(smem_u32(smem + AttnSmem::offK + sn * AttnSmem::kKV));
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          umma_f16(tmem_s + sn * 64, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(&bar_s[sn]);
      }
    }

    // ---- softmax on this thread's row ----
    mbar_wait(&bar_s[st], kv_phase);
    tc_fence_after();
    float s[kKTile];
    {
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(tmem_s + st * 64 + lane_addr, r0);
      tmem_ld_32x32(tmem_s + st * 64 + lane_addr + 32, r1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        s[j] = __uint_as_float(r0[j]);
        s[32 + j] = __uint_as_float(r1[j]);
      }
    }
    if (p.bias) {  // bias tile staged by TMA (128B swizzle: 16-byte chunk c of row r sits at chunk c ^ (r & 7))
      mbar_wait(bar_b, kt & 1);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_row + half * (AttnSmem::kBias / 2) + ((c ^ (tid & 7)) << 4));
          const int j = half * 32 + c * 4;
          s[j] += b4.x; s[j + 1] += b4.y; s[j + 2] += b4.z; s[j + 3] += b4.w;
        }
      }
    }
    const bool need_mask = (k0 + kKTile > p.Tk) || (p.causal && (k0 + kKTile - 1 > q0)) || (kpm_row != nullptr);
    if (need_mask) {
#pragma unroll
      for (int j = 0; j < kKTile; ++j) {
        const int col = k0 + j;
        bool dead = col >= p.Tk || (p.causal && col > row);
        if (!dead && kpm_row) dead = kpm_row[col] != 0;
        if (dead) s[j] = -INFINITY;
      }
    }
    float mx = s[0];
#pragma unroll
    for (int j = 1; j < kKTile; ++j) mx = fmaxf(mx, s[j]);
    const float m_new = fmaxf(m_run, mx * kLog2e);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    const float alpha = fast_exp2(m_run - m_use);  // m_run == -inf -> 0
    float psum = 0.f;
#pragma unroll
    for (int j = 0; j < kKTile; ++j) {
      s[j] = fast_exp2(fmaf(s[j], kLog2e, -m_use));
      psum += s[j];
    }
    l_run = l_run * alpha + psum;
    m_run = m_new;
    // P (bf16) -> smem, K-major 128B-swizzled: chunk c of row r lands at chunk (c ^ (r & 7))
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint4 u;
      u.x = pack_bf16x2(s[8 * c + 0], s[8 * c + 1]);
      u.y = pack_bf16x2(s[8 * c + 2], s[8 * c + 3]);
      u.z = pack_bf16x2(s[8 * c + 4], s[8 * c + 5]);
      u.w = pack_bf16x2(s[8 * c + 6], s[8 * c + 7]);
      *reinterpret_cast<uint4*>(p_row + ((c ^ (tid & 7)) << 4)) = u;
    }
    fence_proxy_async();  // make generic-proxy smem writes visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();

    if (tid == 0) {
      if (p.bias && kt + 1 < n_kt) {  // every thread has consumed this tile's bias (barrier above)
        mbar_expect_tx(bar_b, AttnSmem::kBias);
        tma_load_3d(smem + AttnSmem::offBias, &tmB, bar_b, (kt + 1) * kKTile, q0, h);
        tma_load_3d(smem + AttnSmem::offBias + AttnSmem::kBias / 2, &tmB, bar_b, (kt + 1) * kKTile + 32, q0, h);
      }
      tc_fence_after();
      const uint64_t dp = make_smem_desc_sw128(smem_u32(smem + AttnSmem::offP));
      const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + AttnSmem::offV + st * AttnSmem::kKV));
#pragma unroll
      for (int k = 0; k < kKTile / 16; ++k)  // V: 16 key rows = 2048 B per K step -> +128 in the address field
        umma_f16(tmem_t, dp + 2 * k, dv + 128 * k, idesc_pv, k != 0 ? 1u : 0u);
      umma_commit(bar_o);
    }

    mbar_wait(bar_o, kt & 1);
    tc_fence_after();
    {
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(tmem_t + lane_addr, r0);
      tmem_ld_32x32(tmem_t + lane_addr + 32, r1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        o[j] = fmaf(o[j], alpha, __uint_as_float(r0[j]));
        o[32 + j] = fmaf(o[32 + j], alpha, __uint_as_float(r1[j]));
      }
    }
    tc_fence_before();
    if (tid == 0 && kt + 2 < n_kt) {  // both MMAs that read stage `st` have completed (bar_o)
      mbar_expect_tx(&bar_kv[st], 2 * AttnSmem::kKV);
      tma_load_4d(smem + AttnSmem::offK + st * AttnSmem::kKV, &tmK, &bar_kv[st], 0, h, (kt + 2) * kKTile, b);
      tma_load_4d(smem + AttnSmem::offV + st * AttnSmem::kKV, &tmV, &bar_kv[st], 0, h, (kt + 2) * kKTile, b);
    }
  }

  if (row < p.Tq) {
    const float inv = (1.0f / l_run) * (p.head_scale ? p.head_scale[h] : 1.0f);
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<int64_t>(b) * p.o_batch_stride +
                         static_cast<int64_t>(row) * p.o_row_stride + h * kHeadDim;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint4 u;
      u.x = pack_bf16x2(o[8 * c + 0] * inv, o[8 * c + 1] * inv);
      u.y = pack_bf16x2(o[8 * c + 2] * inv, o[8 * c + 3] * inv);
      u.z = pack_bf16x2(o[8 * c + 4] * inv, o[8 * c + 5] * inv);
      u.w = pack_bf16x2(o[8 * c + 6] * inv, o[8 * c + 7] * inv);
      *reinterpret_cast<uint4*>(dst + 8 * c) = u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
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
    SGF_REQUIRE(a->bias_row_stride % 4 == 0 && a->bias_head_stride % 4 == 0 &&
                    reinterpret_cast<uintptr_t>(a->bias) % 16 == 0 && a->bias_row_stride >= a->Tk,
                "attention: bias must be 16-byte aligned with row/head strides multiples of 4 floats");
  CUtensorMap tmQ, tmK, tmV;
  if (int rc = make_qkv_map(&tmQ, a->q, a->q_row_stride, a->q_batch_stride, a->H, a->Tq, a->B, kQTile)) return rc;
  if (int rc = make_qkv_map(&tmK, a->k, a->k_row_stride, a->k_batch_stride, a->H, a->Tk, a->B, kKTile)) return rc;
  if (int rc = make_qkv_map(&tmV, a->v, a->v_row_stride, a->v_batch_stride, a->H, a->Tk, a->B, kKTile)) return rc;
  CUtensorMap tmB;
  memset(&tmB, 0, sizeof(tmB));
  if (a->bias) {
    uint64_t dims[3] = {static_cast<uint64_t>(a->bias_row_stride), static_cast<uint64_t>(a->Tq), static_cast<uint64_t>(a->H)};
    uint64_t strides[2] = {static_cast<uint64_t>(a->bias_row_stride) * 4, static_cast<uint64_t>(a->bias_head_stride) * 4};
    uint32_t box[3] = {32, kQTile, 1};
    if (int rc = encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, a->bias, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  AttnParams p{a->out, a->o_row_stride, a->o_batch_stride, a->bias, a->bias_head_stride, a->bias_row_stride,
               a->head_scale, a->key_padding_mask, a->B, a->H, a->Tq, a->Tk, a->causal};
  constexpr int smem = AttnSmem::kTotal + 1024;
  static bool configured = false;
  if (!configured) {
    SGF_CHECK_CUDA(
        cudaFuncSetAttribute(attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((a->Tq + kQTile - 1) / kQTile, a->H, a->B);
  SGF_CHECK_CUDA(launch_pdl(attention_tcgen05_kernel, grid, dim3(kAttnThreads), smem,
                            reinterpret_cast<cudaStream_t>(stream), tmQ, tmK, tmV, tmB, p));
  count_launch();
  return SGF_OK;
}
