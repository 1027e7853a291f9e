// Backward of the fused attention (head_dim 64) on sm_100a -- adjoint of attention.cu.
//
//   P = softmax(S), S = Q K^T + bias + mask;  O = head_scale * P V
//   delta_i = sum_d dO_id O_id                                   (delta kernel, HBM-bound)
//   dP = head_scale * dO V^T ;  dS = P o (dP - delta)
//   dQ = dS K (* dq_scale) ;  dK = dS^T Q ;  dV = head_scale * P^T dO
//
// Two tcgen05 kernels with the forward kernel's structure (5 warps: 4 row-owning "softmax" warps +
// one control warp whose single lane issues every TMA load and MMA; mbarrier hand-offs):
//   dQ   kernel, CTA = (128 queries, head, batch), loop over 64-key tiles:
//          S = Q K^T and dP = dO V^T into TMEM, P recomputed from the saved log-sum-exp,
//          dS -> smem (bf16, K-major A operand), dQ += dS K accumulating in TMEM (K tile = MN-major B)
//   dKdV kernel, CTA = (128 keys, head, batch), loop over 64-query tiles:
//          S^T = K Q^T and dP^T = V dO^T into TMEM (thread = key row), P^T / dS^T -> smem,
//          dV += P^T dO and dK += dS^T Q accumulating in TMEM (dO / Q tiles = MN-major B)
// S and dP are recomputed in both kernels (7 tile GEMMs instead of the minimal 5) which keeps both
// free of atomics and deterministic.
#include <string.h>

#include <cuda_fp16.h>

#include "common.cuh"

namespace sgf {

static constexpr int kBT = 128;   // rows owned by a CTA (queries in dQ, keys in dKdV)
static constexpr int kBS = 64;    // streamed tile (keys in dQ, queries in dKdV)
static constexpr int kHd = 64;
static constexpr int kBwdThreads = 160;
static constexpr float kLog2eB = 1.4426950408889634f;

struct AttnBwdParams {
  const __half* bias; int64_t bias_head_stride, bias_row_stride;  // fp16, the tensor the forward kernel streamed
  const float* head_scale;
  const uint8_t* kpm;
  const float* lse;
  const float* delta;
  void* dq; int64_t dq_row_stride, dq_batch_stride; float dq_scale;
  void* dk; int64_t dk_row_stride, dk_batch_stride;
  void* dv; int64_t dv_row_stride, dv_batch_stride;
  int B, H, Tq, Tk, causal;
  float* dbias;  // optional fp32 [H,Tq,bias_row_stride]: += dS summed over the batch (vector atomics)
  const __half* bias_t; int64_t bias_t_head_stride, bias_t_row_stride;  // optional transposed copy [H,Tk,>=roundup(Tq,64)]
};

// ----------------------------------------------------------------------------------------
// delta[b,h,i] = sum_d dout * out ;  d_head_scale[h] += sum delta / head_scale[h]
// one warp per (b,i) token row covering all heads (8 lanes per head when H*64/8 chunks)
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ out, int64_t o_rs, int64_t o_bs,
                                                         const __nv_bfloat16* __restrict__ dout, int64_t do_rs,
                                                         int64_t do_bs, float* __restrict__ delta,
                                                         const float* __restrict__ head_scale,
                                                         float* __restrict__ d_head_scale, int B, int H, int Tq) {
  __shared__ float hacc[64];
  if (threadIdx.x < 64) hacc[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row < B * Tq) {
    const int b = row / Tq, i = row - b * Tq;
    const __nv_bfloat16* o = out + b * o_bs + i * o_rs;
    const __nv_bfloat16* d = dout + b * do_bs + i * do_rs;
    for (int c0 = 0; c0 < H * 8; c0 += 32) {  // chunk c = 8 elements of head c/8
      const int c = c0 + lane;
      const bool ok = c < H * 8;
      const uint4 uo = ok ? *reinterpret_cast<const uint4*>(o + c * 8) : make_uint4(0, 0, 0, 0);
      const uint4 ud = ok ? *reinterpret_cast<const uint4*>(d + c * 8) : make_uint4(0, 0, 0, 0);
      const float2 o0 = unpack_bf16x2(uo.x), o1 = unpack_bf16x2(uo.y), o2 = unpack_bf16x2(uo.z), o3 = unpack_bf16x2(uo.w);
      const float2 d0 = unpack_bf16x2(ud.x), d1 = unpack_bf16x2(ud.y), d2 = unpack_bf16x2(ud.z), d3 = unpack_bf16x2(ud.w);
      float s = o0.x * d0.x + o0.y * d0.y + o1.x * d1.x + o1.y * d1.y + o2.x * d2.x + o2.y * d2.y + o3.x * d3.x + o3.y * d3.y;
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (ok && (lane & 7) == 0) {
        const int h = c >> 3;
        delta[(static_cast<int64_t>(b) * H + h) * Tq + i] = s;
        if (d_head_scale) atomicAdd(&hacc[h], s);
      }
    }
  }
  if (d_head_scale) {
    __syncthreads();
    if (threadIdx.x < H && hacc[threadIdx.x] != 0.f)
      atomicAdd(d_head_scale + threadIdx.x, hacc[threadIdx.x] / head_scale[threadIdx.x]);
  }
}

// ----------------------------------------------------------------------------------------
// dQ kernel
// ----------------------------------------------------------------------------------------
struct DqSmem {
  static constexpr int kQ = kBT * kHd * 2;       // 16 KB (Q, dO)
  static constexpr int kKV = kBS * kHd * 2;      // 8 KB
  static constexpr int kDS = kBT * kBS * 2;      // 16 KB
  static constexpr int kBias = kBT * kBS * 2;    // 16 KB fp16 tile (one 128B-swizzled [128 x 64] box), two buffers
  static constexpr int offQ = 0;
  static constexpr int offDO = offQ + kQ;
  static constexpr int offK = offDO + kQ;
  static constexpr int offV = offK + 2 * kKV;
  static constexpr int offDS = offV + 2 * kKV;
  static constexpr int offBias = offDS + kDS;
  static constexpr int offBar = offBias + 2 * kBias;
  static constexpr int kTotal = offBar + 256;
};
struct DqBars {
  uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2];
  uint64_t sdp_full, sdp_empty, b_full[2], b_empty[2], ds_full, ds_empty, dq_done;
  uint32_t tmem_slot;
};
static_assert(sizeof(DqBars) <= 256, "barrier block");

__global__ void __launch_bounds__(kBwdThreads, 2) attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                     const __grid_constant__ CUtensorMap tmDO,
                                                                     const __grid_constant__ CUtensorMap tmK,
                                                                     const __grid_constant__ CUtensorMap tmV,
                                                                     const __grid_constant__ CUtensorMap tmB,
                                                                     const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  DqBars* bars = reinterpret_cast<DqBars*>(smem + DqSmem::offBar);
  pdl_trigger();
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  // batch fastest: the CTAs that share a bias tile and add into the same dbias lines run together (L2-resident)
  const int b = blockIdx.x;
  const int q0 = blockIdx.y * kBT;
  const int h = blockIdx.z;
  int n_kt = (p.Tk + kBS - 1) / kBS;
  if (p.causal) n_kt = min(n_kt, (min(q0 + kBT, p.Tq) + kBS - 1) / kBS);

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], 1);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], 1);
    }
    mbar_init(&bars->sdp_full, 1);
    mbar_init(&bars->sdp_empty, 128);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->b_full[i], 1);
      mbar_init(&bars->b_empty[i], 128);
    }
    mbar_init(&bars->ds_full, 128);
    mbar_init(&bars->ds_empty, 1);
    mbar_init(&bars->dq_done, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    if ((tid & 31) == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmDO);
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
  const uint32_t tmem_s = tmem_base;         // [0,64)
  const uint32_t tmem_dp = tmem_base + 64;   // [64,128)
  const uint32_t tmem_dq = tmem_base + 128;  // [128,192)
  pdl_wait();

  if (warp == 4) {
    if ((tid & 31) == 0 && n_kt > 0) {
      constexpr uint32_t idesc_nt = make_idesc_bf16(kBT, kBS, 0, 0);  // A, B both K-major
      constexpr uint32_t idesc_nn = make_idesc_bf16(kBT, kHd, 0, 1);  // B MN-major
      auto load_k = [&](int t) {
        const int st = t & 1;
        mbar_expect_tx(&bars->k_full[st], DqSmem::kKV);
        tma_load_4d(smem + DqSmem::offK + st * DqSmem::kKV, &tmK, &bars->k_full[st], 0, h, t * kBS, b);
      };
      auto load_v = [&](int t) {
        const int st = t & 1;
        mbar_expect_tx(&bars->v_full[st], DqSmem::kKV);
        tma_load_4d(smem + DqSmem::offV + st * DqSmem::kKV, &tmV, &bars->v_full[st], 0, h, t * kBS, b);
      };
      auto load_bias = [&](int t) {  // two tiles ahead of the softmax warps
        const int st = t & 1;
        mbar_expect_tx(&bars->b_full[st], DqSmem::kBias);
        tma_load_3d(smem + DqSmem::offBias + st * DqSmem::kBias, &tmB, &bars->b_full[st], t * kBS, q0, h);
      };
      mbar_expect_tx(&bars->q_full, 2 * DqSmem::kQ);
      tma_load_4d(smem + DqSmem::offQ, &tmQ, &bars->q_full, 0, h, q0, b);
      tma_load_4d(smem + DqSmem::offDO, &tmDO, &bars->q_full, 0, h, q0, b);
      for (int t = 0; t < 2 && t < n_kt; ++t) {
        load_k(t);
        load_v(t);
      }
      if (p.bias) {
        load_bias(0);
        if (n_kt > 1) load_bias(1);
      }
      const uint64_t dq_ = make_smem_desc_sw128(smem_u32(smem + DqSmem::offQ));
      const uint64_t ddo = make_smem_desc_sw128(smem_u32(smem + DqSmem::offDO));
      const uint64_t dds = make_smem_desc_sw128(smem_u32(smem + DqSmem::offDS));

#pragma unroll 1
      for (int j = 0; j <= n_kt; ++j) {
        if (j < n_kt) {
          const int st = j & 1;
          if (j == 0) mbar_wait(&bars->q_full, 0);
          mbar_wait(&bars->k_full[st], (j >> 1) & 1);
          mbar_wait(&bars->v_full[st], (j >> 1) & 1);
          if (j >= 1) mbar_wait(&bars->sdp_empty, (j - 1) & 1);  // softmax(j-1) has drained S / dP
          tc_fence_after();
          const uint64_t dk = make_smem_desc_sw128(smem_u32(smem + DqSmem::offK + st * DqSmem::kKV));
          const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + DqSmem::offV + st * DqSmem::kKV));
#pragma unroll
          for (int k = 0; k < kHd / 16; ++k) umma_f16(tmem_s, dq_ + 2 * k, dk + 2 * k, idesc_nt, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < kHd / 16; ++k) umma_f16(tmem_dp, ddo + 2 * k, dv + 2 * k, idesc_nt, k != 0 ? 1u : 0u);
          umma_commit(&bars->sdp_full);
          umma_commit(&bars->v_empty[st]);
          // V(j+1) goes into the stage dP(j-1) used (retired: its commit preceded sdp_full(j-1))
          if (j >= 1 && j + 1 < n_kt) {
            mbar_wait(&bars->v_empty[(j - 1) & 1], ((j - 1) >> 1) & 1);
            load_v(j + 1);
          }
        }
        if (j >= 1) {
          // ---- dQ += dS(j-1) K(j-1) ----
          const int t = j - 1;
          const int st = t & 1;
          mbar_wait(&bars->ds_full, t & 1);
          tc_fence_after();
          const uint64_t dk = make_smem_desc_sw128(smem_u32(smem + DqSmem::offK + st * DqSmem::kKV));
#pragma unroll
          for (int k = 0; k < kBS / 16; ++k) umma_f16(tmem_dq, dds + 2 * k, dk + 128 * k, idesc_nn, (t | k) != 0 ? 1u : 0u);
          umma_commit(&bars->k_empty[st]);
          umma_commit(&bars->ds_empty);
          if (t == n_kt - 1) umma_commit(&bars->dq_done);
          if (j + 1 < n_kt) {  // K(j+1) reuses K(j-1)'s stage once dQ(j-1) has retired
            mbar_wait(&bars->k_empty[st], (t >> 1) & 1);
            load_k(j + 1);
          }
        }
        if (p.bias && j + 2 < n_kt) {  // bias(j+2) reuses the buffer softmax(j) is reading
          mbar_wait(&bars->b_empty[j & 1], (j >> 1) & 1);
          load_bias(j + 2);
        }
      }
    }
  } else {
    const int row = q0 + tid;
    const bool row_ok = row < p.Tq;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const uint8_t* bias_row0 = smem + DqSmem::offBias + tid * 128;  // this thread's 64-half row inside a bias buffer
    const uint8_t* kpm_row = p.kpm ? p.kpm + static_cast<int64_t>(b) * p.Tk : nullptr;
    uint8_t* ds_row = smem + DqSmem::offDS + tid * 128;
    const int64_t stat_idx = (static_cast<int64_t>(b) * p.H + h) * p.Tq + row;
    const float L2 = row_ok ? p.lse[stat_idx] : 0.f;
    const float dl = row_ok ? p.delta[stat_idx] : 0.f;
    const float hs = p.head_scale ? p.head_scale[h] : 1.0f;

#pragma unroll 1
    for (int j = 0; j < n_kt; ++j) {
      const int k0 = j * kBS;
      mbar_wait(&bars->sdp_full, j & 1);
      if (p.bias) mbar_wait(&bars->b_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint8_t* bias_row = bias_row0 + (j & 1) * DqSmem::kBias;
      const bool need_mask = (k0 + kBS > p.Tk) || (p.causal && (k0 + kBS - 1 > q0)) || (kpm_row != nullptr);
      uint32_t packed[32];  // dS row as bf16 pairs
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t rs[32], rp[32];
        tmem_ld_32x32(tmem_s + lane_addr + half * 32, rs);
        tmem_ld_32x32(tmem_dp + lane_addr + half * 32, rp);
        uint4 bu[4];  // fp16 bias of columns half*32 .. half*32+31 (chunks 4*half .. 4*half+3 of the swizzled row)
        if (p.bias) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            bu[c] = *reinterpret_cast<const uint4*>(bias_row + (((half * 4 + c) ^ (tid & 7)) << 4));
        }
        tmem_ld_wait();
        const float2 l2e = splat2(kLog2eB), nL2 = splat2(-L2), ndl = splat2(-dl), hs2 = splat2(hs);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t w[4] = {bu[c].x, bu[c].y, bu[c].z, bu[c].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int i = 8 * c + 2 * q;
            float2 sv = make_float2(__uint_as_float(rs[i]), __uint_as_float(rs[i + 1]));
            if (p.bias) sv = add2(sv, __half22float2(*reinterpret_cast<const __half2*>(&w[q])));
            const float2 e = fma2(sv, l2e, nL2);
            float2 pr = make_float2(fast_exp2(e.x), fast_exp2(e.y));
            if (need_mask) {  // CTA-uniform: boundary / causal-diagonal / padded tiles only
              const int col = k0 + half * 32 + i;
              bool dead0 = col >= p.Tk || (p.causal && col > row);
              bool dead1 = col + 1 >= p.Tk || (p.causal && col + 1 > row);
              if (kpm_row) {
                if (!dead0) dead0 = kpm_row[col] != 0;
                if (!dead1) dead1 = kpm_row[col + 1] != 0;
              }
              if (dead0) pr.x = 0.f;
              if (dead1) pr.y = 0.f;
            }
            const float2 ds = mul2(pr, fma2(hs2, make_float2(__uint_as_float(rp[i]), __uint_as_float(rp[i + 1])), ndl));
            rs[i] = __float_as_uint(ds.x);
            rs[i + 1] = __float_as_uint(ds.y);
          }
        }
        if (p.dbias && row_ok) {  // d(bias)[h,i,j] += dS: the bias is shared by the batch (and, for abs, by the layers)
          float* db = p.dbias + static_cast<int64_t>(h) * p.bias_head_stride + static_cast<int64_t>(row) * p.bias_row_stride +
                      k0 + half * 32;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            red_add_f32x4(db + 4 * c, __uint_as_float(rs[4 * c]), __uint_as_float(rs[4 * c + 1]),
                          __uint_as_float(rs[4 * c + 2]), __uint_as_float(rs[4 * c + 3]));
        }
#pragma unroll
        for (int i = 0; i < 16; ++i)
          packed[half * 16 + i] = pack_bf16x2(__uint_as_float(rs[2 * i]), __uint_as_float(rs[2 * i + 1]));
      }
      tc_fence_before();
      mbar_arrive(&bars->sdp_empty);
      if (p.bias) mbar_arrive(&bars->b_empty[j & 1]);
      if (j >= 1) mbar_wait(&bars->ds_empty, (j - 1) & 1);  // dQ(j-1) has consumed the dS buffer
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 u = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
        *reinterpret_cast<uint4*>(ds_row + ((c ^ (tid & 7)) << 4)) = u;
      }
      fence_proxy_async();
      mbar_arrive(&bars->ds_full);
    }
    if (n_kt > 0) {
      mbar_wait(&bars->dq_done, 0);
      tc_fence_after();
    }
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.dq) + static_cast<int64_t>(b) * p.dq_batch_stride +
                         static_cast<int64_t>(row) * p.dq_row_stride + h * kHd;
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      uint32_t r[32];
      if (n_kt > 0) {
        tmem_ld_32x32(tmem_dq + lane_addr + hb * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
      if (row_ok) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(r[8 * c + 0]) * p.dq_scale, __uint_as_float(r[8 * c + 1]) * p.dq_scale);
          u.y = pack_bf16x2(__uint_as_float(r[8 * c + 2]) * p.dq_scale, __uint_as_float(r[8 * c + 3]) * p.dq_scale);
          u.z = pack_bf16x2(__uint_as_float(r[8 * c + 4]) * p.dq_scale, __uint_as_float(r[8 * c + 5]) * p.dq_scale);
          u.w = pack_bf16x2(__uint_as_float(r[8 * c + 6]) * p.dq_scale, __uint_as_float(r[8 * c + 7]) * p.dq_scale);
          *reinterpret_cast<uint4*>(dst + hb * 32 + 8 * c) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------------
// dK / dV kernel
// ----------------------------------------------------------------------------------------
struct DkvSmem {
  static constexpr int kKV = kBT * kHd * 2;     // 16 KB (K, V: 128 keys)
  static constexpr int kQ = kBS * kHd * 2;      // 8 KB (Q, dO tiles of 64 queries)
  static constexpr int kP = kBT * kBS * 2;      // 16 KB (P^T, dS^T)
  static constexpr int offK = 0;
  static constexpr int offV = offK + kKV;
  static constexpr int offQ = offV + kKV;
  static constexpr int offDO = offQ + 2 * kQ;
  static constexpr int offP = offDO + 2 * kQ;
  static constexpr int offDS = offP + kP;
  static constexpr int offStat = offDS + kP;       // [2 stages][lse 64 | delta 64] fp32
  static constexpr int offBar = offStat + 2 * 128 * 4;
  static constexpr int kTotal = offBar + 256;
};
struct DkvBars {
  uint64_t kv_full, q_full[2], q_empty[2];
  uint64_t sdp_full, sdp_empty, pds_full, pds_empty, done;
  uint32_t tmem_slot;
};
static_assert(sizeof(DkvBars) <= 256, "barrier block");

__global__ void __launch_bounds__(kBwdThreads, 2) attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                      const __grid_constant__ CUtensorMap tmDO,
                                                                      const __grid_constant__ CUtensorMap tmK,
                                                                      const __grid_constant__ CUtensorMap tmV,
                                                                      const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  DkvBars* bars = reinterpret_cast<DkvBars*>(smem + DkvSmem::offBar);
  pdl_trigger();
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int k0 = blockIdx.x * kBT;  // (key-tile fastest: batch-fastest ordering measured slower for this kernel)
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_qt_all = (p.Tq + kBS - 1) / kBS;
  const int t_first = p.causal ? min(k0 / kBS, n_qt_all) : 0;  // queries i >= key index only
  const int n_t = n_qt_all - t_first;

  if (tid == 0) {
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    mbar_init(&bars->kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->q_full[i], 1);
      mbar_init(&bars->q_empty[i], 1);
    }
    mbar_init(&bars->sdp_full, 1);
    mbar_init(&bars->sdp_empty, 128);
    mbar_init(&bars->pds_full, 128);
    mbar_init(&bars->pds_empty, 1);
    mbar_init(&bars->done, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    if ((tid & 31) == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmDO);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
    }
    tmem_alloc<256>(&bars->tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  const uint32_t tmem_s = tmem_base;          // S^T   [0,64)
  const uint32_t tmem_dp = tmem_base + 64;    // dP^T  [64,128)
  const uint32_t tmem_dv = tmem_base + 128;   // dV    [128,192)
  const uint32_t tmem_dk = tmem_base + 192;   // dK    [192,256)
  pdl_wait();

  if (warp == 4) {
    if ((tid & 31) == 0 && n_t > 0) {
      constexpr uint32_t idesc_nt = make_idesc_bf16(kBT, kBS, 0, 0);
      constexpr uint32_t idesc_nn = make_idesc_bf16(kBT, kHd, 0, 1);
      auto load_q = [&](int t) {  // Q and dO tiles of query tile (t_first + t)
        const int st = t & 1;
        mbar_expect_tx(&bars->q_full[st], 2 * DkvSmem::kQ);
        tma_load_4d(smem + DkvSmem::offQ + st * DkvSmem::kQ, &tmQ, &bars->q_full[st], 0, h, (t_first + t) * kBS, b);
        tma_load_4d(smem + DkvSmem::offDO + st * DkvSmem::kQ, &tmDO, &bars->q_full[st], 0, h, (t_first + t) * kBS, b);
      };
      mbar_expect_tx(&bars->kv_full, 2 * DkvSmem::kKV);
      tma_load_4d(smem + DkvSmem::offK, &tmK, &bars->kv_full, 0, h, k0, b);
      tma_load_4d(smem + DkvSmem::offV, &tmV, &bars->kv_full, 0, h, k0, b);
      for (int t = 0; t < 2 && t < n_t; ++t) load_q(t);
      const uint64_t dk = make_smem_desc_sw128(smem_u32(smem + DkvSmem::offK));
      const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + DkvSmem::offV));
      const uint64_t dp = make_smem_desc_sw128(smem_u32(smem + DkvSmem::offP));
      const uint64_t dds = make_smem_desc_sw128(smem_u32(smem + DkvSmem::offDS));

#pragma unroll 1
      for (int j = 0; j <= n_t; ++j) {
        if (j < n_t) {
          const int st = j & 1;
          if (j == 0) mbar_wait(&bars->kv_full, 0);
          mbar_wait(&bars->q_full[st], (j >> 1) & 1);
          if (j >= 1) mbar_wait(&bars->sdp_empty, (j - 1) & 1);
          tc_fence_after();
          const uint64_t dqt = make_smem_desc_sw128(smem_u32(smem + DkvSmem::offQ + st * DkvSmem::kQ));
          const uint64_t ddo = make_smem_desc_sw128(smem_u32(smem + DkvSmem::offDO + st * DkvSmem::kQ));
#pragma unroll
          for (int k = 0; k < kHd / 16; ++k) umma_f16(tmem_s, dk + 2 * k, dqt + 2 * k, idesc_nt, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < kHd / 16; ++k) umma_f16(tmem_dp, dv + 2 * k, ddo + 2 * k, idesc_nt, k != 0 ? 1u : 0u);
          umma_commit(&bars->sdp_full);
        }
        if (j >= 1) {
          const int t = j - 1;
          const int st = t & 1;
          mbar_wait(&bars->pds_full, t & 1);
          tc_fence_after();
          const uint64_t dqt = make_smem_desc_sw128(smem_u32(smem + DkvSmem::offQ + st * DkvSmem::kQ));
          const uint64_t ddo = make_smem_desc_sw128(smem_u32(smem + DkvSmem::offDO + st * DkvSmem::kQ));
#pragma unroll
          for (int k = 0; k < kBS / 16; ++k) umma_f16(tmem_dv, dp + 2 * k, ddo + 128 * k, idesc_nn, (t | k) != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < kBS / 16; ++k) umma_f16(tmem_dk, dds + 2 * k, dqt + 128 * k, idesc_nn, (t | k) != 0 ? 1u : 0u);
          umma_commit(&bars->q_empty[st]);
          umma_commit(&bars->pds_empty);
          if (t == n_t - 1) umma_commit(&bars->done);
          if (j + 1 < n_t) {  // Q/dO(j+1) reuse the stage of tile j-1
            mbar_wait(&bars->q_empty[st], (t >> 1) & 1);
            load_q(j + 1);
          }
        }
      }
    }
  } else {
    const int key = k0 + tid;
    const bool key_ok = key < p.Tk && !(p.kpm && p.kpm[static_cast<int64_t>(b) * p.Tk + key] != 0);
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    uint8_t* p_row = smem + DkvSmem::offP + tid * 128;
    uint8_t* ds_row = smem + DkvSmem::offDS + tid * 128;
    const float hs = p.head_scale ? p.head_scale[h] : 1.0f;
    const float* lse_bh = p.lse + (static_cast<int64_t>(b) * p.H + h) * p.Tq;
    const float* dl_bh = p.delta + (static_cast<int64_t>(b) * p.H + h) * p.Tq;
    // key-major copy of the bias: this thread's 64 query values of a tile are 128 contiguous bytes
    const __half* bias_tk = p.bias_t ? p.bias_t + static_cast<int64_t>(h) * p.bias_t_head_stride +
                                           static_cast<int64_t>(min(key, p.Tk - 1)) * p.bias_t_row_stride
                                     : nullptr;

    float* stat = reinterpret_cast<float*>(smem + DkvSmem::offStat);
#pragma unroll 1
    for (int j = 0; j < n_t; ++j) {
      const int qb = (t_first + j) * kBS;
      // this key's 64 bias values of the tile: the first 32 are requested before the wait for S / dP, the other 32
      // while the first half is being processed
      uint4 cbias[2][4];
      if (bias_tk) {
#pragma unroll
        for (int c = 0; c < 4; ++c) cbias[0][c] = __ldg(reinterpret_cast<const uint4*>(bias_tk + qb + 8 * c));
      }
      // per-query statistics of this tile -> shared memory (one element per thread), read back as broadcast float4
      {
        const int q = qb + (tid & 63);
        const float* src = tid < 64 ? lse_bh : dl_bh;
        stat[(j & 1) * 128 + tid] = q < p.Tq ? -__ldg(src + q) : (tid < 64 ? -INFINITY : 0.f);  // negated: FMA addends
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      const float4* st_l = reinterpret_cast<const float4*>(stat + (j & 1) * 128);
      const float4* st_d = st_l + 16;
      mbar_wait(&bars->sdp_full, j & 1);
      tc_fence_after();
      const bool mask_tile = p.causal && (k0 + kBT - 1 > qb);
      uint32_t pk_p[32], pk_ds[32];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t rs[32], rp[32];
        tmem_ld_32x32(tmem_s + lane_addr + half * 32, rs);
        tmem_ld_32x32(tmem_dp + lane_addr + half * 32, rp);
        if (bias_tk && half == 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) cbias[1][c] = __ldg(reinterpret_cast<const uint4*>(bias_tk + qb + 32 + 8 * c));
        }
        tmem_ld_wait();
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 l4 = st_l[half * 8 + i4];  // -lse (log2 domain) of the four queries
          const float4 d4 = st_d[half * 8 + i4];  // -delta
          float2 b01 = splat2(0.f), b23 = splat2(0.f);
          if (bias_tk) {
            const uint4 ub = cbias[half][i4 >> 1];
            const uint32_t w0 = (i4 & 1) ? ub.z : ub.x, w1 = (i4 & 1) ? ub.w : ub.y;
            b01 = __half22float2(*reinterpret_cast<const __half2*>(&w0));
            b23 = __half22float2(*reinterpret_cast<const __half2*>(&w1));
          }
          const int i = i4 * 4;
          const float2 e01 = fma2(add2(make_float2(__uint_as_float(rs[i]), __uint_as_float(rs[i + 1])), b01),
                                  splat2(kLog2eB), make_float2(l4.x, l4.y));
          const float2 e23 = fma2(add2(make_float2(__uint_as_float(rs[i + 2]), __uint_as_float(rs[i + 3])), b23),
                                  splat2(kLog2eB), make_float2(l4.z, l4.w));
          float2 p01 = make_float2(fast_exp2(e01.x), fast_exp2(e01.y));  // queries >= Tq carry -lse = -inf -> 0
          float2 p23 = make_float2(fast_exp2(e23.x), fast_exp2(e23.y));
          if (mask_tile) {  // the causal diagonal crosses this tile
            const int q = qb + half * 32 + i;
            if (key > q) p01.x = 0.f;
            if (key > q + 1) p01.y = 0.f;
            if (key > q + 2) p23.x = 0.f;
            if (key > q + 3) p23.y = 0.f;
          }
          const float2 s01 = mul2(p01, fma2(splat2(hs), make_float2(__uint_as_float(rp[i]), __uint_as_float(rp[i + 1])),
                                            make_float2(d4.x, d4.y)));
          const float2 s23 = mul2(p23, fma2(splat2(hs), make_float2(__uint_as_float(rp[i + 2]), __uint_as_float(rp[i + 3])),
                                            make_float2(d4.z, d4.w)));
          rs[i] = __float_as_uint(p01.x); rs[i + 1] = __float_as_uint(p01.y);
          rs[i + 2] = __float_as_uint(p23.x); rs[i + 3] = __float_as_uint(p23.y);
          rp[i] = __float_as_uint(s01.x); rp[i + 1] = __float_as_uint(s01.y);
          rp[i + 2] = __float_as_uint(s23.x); rp[i + 3] = __float_as_uint(s23.y);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          pk_p[half * 16 + i] = pack_bf16x2(__uint_as_float(rs[2 * i]), __uint_as_float(rs[2 * i + 1]));
          pk_ds[half * 16 + i] = pack_bf16x2(__uint_as_float(rp[2 * i]), __uint_as_float(rp[2 * i + 1]));
        }
      }
      if (!key_ok) {  // padded / out-of-range key: its whole row of P^T and dS^T is zero
#pragma unroll
        for (int i = 0; i < 32; ++i) pk_p[i] = pk_ds[i] = 0u;
      }
      tc_fence_before();
      mbar_arrive(&bars->sdp_empty);
      if (j >= 1) mbar_wait(&bars->pds_empty, (j - 1) & 1);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int o = (c ^ (tid & 7)) << 4;
        *reinterpret_cast<uint4*>(p_row + o) = make_uint4(pk_p[4 * c], pk_p[4 * c + 1], pk_p[4 * c + 2], pk_p[4 * c + 3]);
        *reinterpret_cast<uint4*>(ds_row + o) = make_uint4(pk_ds[4 * c], pk_ds[4 * c + 1], pk_ds[4 * c + 2], pk_ds[4 * c + 3]);
      }
      fence_proxy_async();
      mbar_arrive(&bars->pds_full);
    }
    if (n_t > 0) {
      mbar_wait(&bars->done, 0);
      tc_fence_after();
    }
    __nv_bfloat16* dst_v = reinterpret_cast<__nv_bfloat16*>(p.dv) + static_cast<int64_t>(b) * p.dv_batch_stride +
                           static_cast<int64_t>(key) * p.dv_row_stride + h * kHd;
    __nv_bfloat16* dst_k = reinterpret_cast<__nv_bfloat16*>(p.dk) + static_cast<int64_t>(b) * p.dk_batch_stride +
                           static_cast<int64_t>(key) * p.dk_row_stride + h * kHd;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const uint32_t tm = which == 0 ? tmem_dv : tmem_dk;
      __nv_bfloat16* dst = which == 0 ? dst_v : dst_k;
      const float sc = which == 0 ? hs : 1.0f;
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        uint32_t r[32];
        if (n_t > 0) {
          tmem_ld_32x32(tm + lane_addr + hb * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        if (key < p.Tk) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(r[8 * c + 0]) * sc, __uint_as_float(r[8 * c + 1]) * sc);
            u.y = pack_bf16x2(__uint_as_float(r[8 * c + 2]) * sc, __uint_as_float(r[8 * c + 3]) * sc);
            u.z = pack_bf16x2(__uint_as_float(r[8 * c + 4]) * sc, __uint_as_float(r[8 * c + 5]) * sc);
            u.w = pack_bf16x2(__uint_as_float(r[8 * c + 6]) * sc, __uint_as_float(r[8 * c + 7]) * sc);
            *reinterpret_cast<uint4*>(dst + hb * 32 + 8 * c) = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// out[h][c][r] = in[h][r][c] for 16-bit elements (the key-major copy of an fp16 bias for the dK/dV kernel); rows of
// `out` beyond R (up to the padded row length) are zero-filled.  64x64 tiles through shared memory, 128-bit accesses.
__global__ void __launch_bounds__(256) transpose16_batched_kernel(const uint16_t* __restrict__ in, int64_t in_hs,
                                                                 int64_t in_rs, int R, int C, uint16_t* __restrict__ out,
                                                                 int64_t out_hs, int64_t out_rs) {
  __shared__ uint16_t tile[64][72];
  const uint16_t* ih = in + static_cast<int64_t>(blockIdx.z) * in_hs;
  uint16_t* oh = out + static_cast<int64_t>(blockIdx.z) * out_hs;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tr = threadIdx.x >> 2, seg = (threadIdx.x & 3) * 16;
  {
    const int r = r0 + tr;
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    if (r < R) {
      if (c0 + seg + 8 <= C) a = *reinterpret_cast<const uint4*>(ih + static_cast<int64_t>(r) * in_rs + c0 + seg);
      if (c0 + seg + 16 <= C) b = *reinterpret_cast<const uint4*>(ih + static_cast<int64_t>(r) * in_rs + c0 + seg + 8);
    }
    *reinterpret_cast<uint4*>(&tile[tr][seg]) = a;
    *reinterpret_cast<uint4*>(&tile[tr][seg + 8]) = b;
  }
  __syncthreads();
  const int c = c0 + tr;
  if (c < C) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      w[i] = static_cast<uint32_t>(tile[seg + 2 * i][tr]) | (static_cast<uint32_t>(tile[seg + 2 * i + 1][tr]) << 16);
    uint16_t* dst = oh + static_cast<int64_t>(c) * out_rs + r0 + seg;
    *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4*>(dst + 8) = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

static int make_hd_map(CUtensorMap* m, const void* base, int64_t row_stride, int64_t batch_stride, int H, int T, int B,
                       int box_rows) {
  uint64_t dims[4] = {kHd, static_cast<uint64_t>(H), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
  uint64_t strides[3] = {kHd * 2, static_cast<uint64_t>(row_stride) * 2, static_cast<uint64_t>(batch_stride) * 2};
  uint32_t box[4] = {kHd, 1, static_cast<uint32_t>(box_rows), 1};
  return encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace sgf

using namespace sgf;

extern "C" int sgf_attention_bwd_bf16(const sgf_attention_bwd_args* a, void* stream) {
  SGF_REQUIRE(a != nullptr && a->q && a->k && a->v && a->out && a->dout && a->dq && a->dk && a->dv && a->lse && a->delta,
              "attention_bwd: null pointer");
  SGF_REQUIRE(a->B > 0 && a->H > 0 && a->H <= 64 && a->Tq > 0 && a->Tk > 0, "attention_bwd: bad shape");
  const int64_t strides[] = {a->q_row_stride, a->q_batch_stride, a->k_row_stride, a->k_batch_stride, a->v_row_stride,
                             a->v_batch_stride, a->o_row_stride, a->o_batch_stride, a->do_row_stride, a->do_batch_stride,
                             a->dq_row_stride, a->dq_batch_stride, a->dk_row_stride, a->dk_batch_stride,
                             a->dv_row_stride, a->dv_batch_stride};
  for (int64_t s : strides) SGF_REQUIRE(s % 8 == 0, "attention_bwd: strides must be multiples of 8 elements");
  const void* ptrs[] = {a->q, a->k, a->v, a->out, a->dout, a->dq, a->dk, a->dv};
  for (const void* ptr : ptrs) SGF_REQUIRE(reinterpret_cast<uintptr_t>(ptr) % 16 == 0, "attention_bwd: 16-byte alignment");
  if (a->bias)
    SGF_REQUIRE(a->bias_row_stride % 8 == 0 && a->bias_head_stride % 8 == 0 &&
                    reinterpret_cast<uintptr_t>(a->bias) % 16 == 0 && a->bias_row_stride >= a->Tk,
                "attention_bwd: the fp16 bias must be 16-byte aligned with row/head strides multiples of 8 elements");
  SGF_REQUIRE(!a->d_head_scale || a->head_scale, "attention_bwd: d_head_scale needs head_scale");
  SGF_REQUIRE(!a->bias || a->bias_t, "attention_bwd: a bias needs its key-major copy bias_t (sgf_transpose16_batched)");
  if (a->bias_t)
    SGF_REQUIRE(a->bias && a->bias_t_row_stride % 8 == 0 && a->bias_t_head_stride % 8 == 0 &&
                    reinterpret_cast<uintptr_t>(a->bias_t) % 16 == 0 && a->bias_t_row_stride >= (a->Tq + 63) / 64 * 64,
                "attention_bwd: bias_t must be the 16-byte aligned [H,Tk,>=roundup(Tq,64)] transpose of bias");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  {
    const int rows = a->B * a->Tq;
    attn_delta_kernel<<<(rows + 7) / 8, 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(a->out), a->o_row_stride, a->o_batch_stride,
        reinterpret_cast<const __nv_bfloat16*>(a->dout), a->do_row_stride, a->do_batch_stride, a->delta, a->head_scale,
        a->d_head_scale, a->B, a->H, a->Tq);
    SGF_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  AttnBwdParams p{reinterpret_cast<const __half*>(a->bias), a->bias_head_stride, a->bias_row_stride, a->head_scale, a->key_padding_mask, a->lse,
                  a->delta, a->dq, a->dq_row_stride, a->dq_batch_stride, a->dq_scale, a->dk, a->dk_row_stride,
                  a->dk_batch_stride, a->dv, a->dv_row_stride, a->dv_batch_stride, a->B, a->H, a->Tq, a->Tk, a->causal,
                  a->dbias, reinterpret_cast<const __half*>(a->bias_t), a->bias_t_head_stride, a->bias_t_row_stride};
  SGF_REQUIRE(!a->dbias || (a->bias && reinterpret_cast<uintptr_t>(a->dbias) % 16 == 0 && a->bias_row_stride % 64 == 0),
              "attention_bwd: dbias needs bias (same strides), 16-byte alignment and a row stride padded to 64");
  static bool configured = false;
  if (!configured) {
    SGF_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem::kTotal));
    SGF_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvSmem::kTotal));
    configured = true;
  }
  {
    CUtensorMap tmQ, tmDO, tmK, tmV, tmB;
    if (int rc = make_hd_map(&tmQ, a->q, a->q_row_stride, a->q_batch_stride, a->H, a->Tq, a->B, kBT)) return rc;
    if (int rc = make_hd_map(&tmDO, a->dout, a->do_row_stride, a->do_batch_stride, a->H, a->Tq, a->B, kBT)) return rc;
    if (int rc = make_hd_map(&tmK, a->k, a->k_row_stride, a->k_batch_stride, a->H, a->Tk, a->B, kBS)) return rc;
    if (int rc = make_hd_map(&tmV, a->v, a->v_row_stride, a->v_batch_stride, a->H, a->Tk, a->B, kBS)) return rc;
    memset(&tmB, 0, sizeof(tmB));
    if (a->bias) {
      uint64_t dims[3] = {static_cast<uint64_t>(a->bias_row_stride), static_cast<uint64_t>(a->Tq), static_cast<uint64_t>(a->H)};
      uint64_t strides[2] = {static_cast<uint64_t>(a->bias_row_stride) * 2, static_cast<uint64_t>(a->bias_head_stride) * 2};
      uint32_t box[3] = {kBS, kBT, 1};
      if (int rc = encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, a->bias, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))
        return rc;
    }
    dim3 grid(a->B, (a->Tq + kBT - 1) / kBT, a->H);
    SGF_CHECK_CUDA(launch_pdl(attn_bwd_dq_kernel, grid, dim3(kBwdThreads), static_cast<size_t>(DqSmem::kTotal), st, tmQ,
                              tmDO, tmK, tmV, tmB, p));
    count_launch();
  }
  {
    CUtensorMap tmQ, tmDO, tmK, tmV;
    if (int rc = make_hd_map(&tmQ, a->q, a->q_row_stride, a->q_batch_stride, a->H, a->Tq, a->B, kBS)) return rc;
    if (int rc = make_hd_map(&tmDO, a->dout, a->do_row_stride, a->do_batch_stride, a->H, a->Tq, a->B, kBS)) return rc;
    if (int rc = make_hd_map(&tmK, a->k, a->k_row_stride, a->k_batch_stride, a->H, a->Tk, a->B, kBT)) return rc;
    if (int rc = make_hd_map(&tmV, a->v, a->v_row_stride, a->v_batch_stride, a->H, a->Tk, a->B, kBT)) return rc;
    dim3 grid((a->Tk + kBT - 1) / kBT, a->H, a->B);
    SGF_CHECK_CUDA(launch_pdl(attn_bwd_dkv_kernel, grid, dim3(kBwdThreads), static_cast<size_t>(DkvSmem::kTotal), st,
                              tmQ, tmDO, tmK, tmV, p));
    count_launch();
  }
  return SGF_OK;
}

extern "C" int sgf_transpose16_batched(const void* in, int64_t in_head_stride, int64_t in_row_stride, int32_t H, int32_t R,
                                       int32_t C_, void* out, int64_t out_head_stride, int64_t out_row_stride,
                                       void* stream) {
  SGF_REQUIRE(in && out && H > 0 && R > 0 && C_ > 0, "transpose16_batched: bad arguments");
  SGF_REQUIRE(C_ % 8 == 0 && in_row_stride % 8 == 0 && in_row_stride >= C_ && in_head_stride % 8 == 0 &&
                  out_row_stride % 8 == 0 && out_row_stride >= (R + 63) / 64 * 64 && out_head_stride % 8 == 0 &&
                  reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
              "transpose16_batched: 16-byte alignment; C a multiple of 8; out rows padded to a multiple of 64");
  dim3 grid((C_ + 63) / 64, (R + 63) / 64, H);
  transpose16_batched_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint16_t*>(in), in_head_stride, in_row_stride, R, C_, reinterpret_cast<uint16_t*>(out),
      out_head_stride, out_row_stride);
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}
