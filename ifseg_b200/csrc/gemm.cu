// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )     bf16 operands, fp32 accumulation in TMEM.
//
// This file: the one-tile-per-CTA kernel -- the general path (any N, batched operands, runtime epilogue) and the
// fallback of the persistent CTA-pair kernel in gemm_ts.cu, which takes every hot GEMM / convolution of the segofa path --
// and the mixed-major kernel of the dense adjoints.
// One CTA per 128 x BN output tile, 192/320 threads, warp-specialised:
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (tcgen05.commit frees stages)
//   warps 2..   : epilogue       (tcgen05.ld 32x32b -> column ops -> smem staging -> coalesced
//                                 residual add / activation / store), 4 or 8 warps
// The same main loop serves the 3x3/s1 convolution: the A tile of filter tap (ky,kx) is a 4-D
// TMA box [64 ch, bw, bh, 1] of the NHWC input shifted by (kx-1, ky-1); TMA's out-of-bounds
// zero fill is the convolution padding, so no im2col matrix ever exists.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gemm_common.cuh"

namespace sgf {

// Epilogue shared by the 1-CTA and the 2-CTA kernels.  Executed by the kEpiWarps epilogue warps
// (warp index 2..): `stage_bytes_avail` bytes at `smem` (the retired pipeline stages) hold the staging tiles.
template <int BN, int kEpiWarps, int kStageBytesAvail, bool kConv, int kEpi>
SGF_DEVICE void gemm_epilogue(uint8_t* smem, uint32_t tmem_base, uint64_t* accum_bar, const GemmShape& shp,
                              const GemmEpilogue& ep, int m0, int n0, int z, int img, int h0, int w0, int warp,
                              int lane) {
  constexpr bool kRt = (kEpi & kEpiRuntime) != 0;
  // ------------------------------ epilogue warps ------------------------------
  // kEpiWarps warps; warp (quarter, part) owns TMEM lanes [32*quarter, +32) and columns
  // [part*kCols, +kCols) of the tile.
  // Phase 1: TMEM -> registers (thread = row), column-wise ops (row-norm, scale, bias, q-scale, GELU),
  //          row parked in a per-warp smem staging tile (the pipeline stages are dead by then).
  // Phase 2: the tile is re-read with lanes along the columns, so the residual add and the output
  //          stores are fully coalesced 16/32-byte-per-lane row segments.
  constexpr int kSplit = kEpiWarps / 4;
  constexpr int kCols = BN / kSplit;
  constexpr int kRowPitch = kCols * 4 + 16;  // bytes; +16 keeps the thread-per-row float4 writes conflict-free
  constexpr int kLanesPerRow = kCols / 8;
  constexpr int kRowsPerIter = 32 / kLanesPerRow;
  constexpr int kIters = 32 / kRowsPerIter;
  static_assert(kEpiWarps * 32 * kRowPitch <= kStageBytesAvail, "epilogue staging must fit in the pipeline smem");
  const int ew = warp - 2;
  const int quarter = warp & 3;  // TMEM lane quarter this warp may access
  const int part = ew >> 2;
  uint8_t* stage = smem + ew * (32 * kRowPitch);
  const int col = part * kCols + (lane % kLanesPerRow) * 8;
  const int c = n0 + col;
  // specialised kernels are only dispatched for N % 32 == 0: a lane's 8 columns are all in or all out
  const bool vec_ok = kRt ? ((shp.N % 8) == 0 && (c + 8 <= shp.N)) : (c < shp.N);
  const bool out_f32 = epi_has<kEpi, kEpiOutF32>(ep.c_dtype == SGF_F32);
  const bool res_f32 = epi_has<kEpi, kEpiResF32>(ep.residual && ep.r_dtype == SGF_F32);
  const bool res_b16 = epi_has<kEpi, kEpiResBf16>(ep.residual && ep.r_dtype != SGF_F32);
  const bool do_relu = epi_has<kEpi, kEpiRelu>(ep.act == SGF_ACT_RELU);
  const int csz = out_f32 ? 4 : 2;
  const int rsz = res_f32 ? 4 : 2;

  // folded-LayerNorm row statistics of this thread's row: summed (fixed order -> deterministic) while the
  // main loop is still running
  float rn_mean = 0.f, rn_rstd = 1.f;
  const bool rn_on = !kConv && epi_has<kEpi, kEpiRowNorm>(ep.rownorm_stats != nullptr);
  if (rn_on) {
    const int mrow = min(m0 + quarter * 32 + lane, shp.M - 1);
    const float4* sp = reinterpret_cast<const float4*>(ep.rownorm_stats) +
                       (static_cast<int64_t>(z) * shp.M + mrow) * (ep.rownorm_parts / 2);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
    for (int q = 0; q < ep.rownorm_parts / 2; ++q) {
      const float4 t = __ldg(sp + q);
      s0 += t.x; s1 += t.y; s0 += t.z; s1 += t.w;
    }
    rn_mean = s0 * ep.rownorm_inv_dim;
    rn_rstd = rsqrtf(fmaxf(s1 * ep.rownorm_inv_dim - rn_mean * rn_mean, 0.f) + 1e-5f);
  }

  mbar_wait(accum_bar, 0);
  tc_fence_after();

  // ---- phase 1 ----
#pragma unroll 1
  for (int ch = 0; ch < kCols / 32; ++ch) {
    const int c0 = n0 + part * kCols + ch * 32;
    if (c0 >= shp.N) break;  // warp-uniform
    uint32_t acc[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + part * kCols + ch * 32, acc);
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
    const bool full32 = kRt ? (c0 + 32 <= shp.N && (shp.N % 4) == 0) : true;
    if (full32) {
      if (rn_on) {  // folded LayerNorm of the A operand: v = rstd * (v - mean * u[n])
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 u4 = __ldg(reinterpret_cast<const float4*>(ep.rownorm_u + c0 + j));
          v[j] = rn_rstd * (v[j] - rn_mean * u4.x); v[j + 1] = rn_rstd * (v[j + 1] - rn_mean * u4.y);
          v[j + 2] = rn_rstd * (v[j + 2] - rn_mean * u4.z); v[j + 3] = rn_rstd * (v[j + 3] - rn_mean * u4.w);
        }
      }
      if (epi_has<kEpi, kEpiScale>(ep.col_scale != nullptr)) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(ep.col_scale + c0 + j));
          v[j] *= s4.x; v[j + 1] *= s4.y; v[j + 2] *= s4.z; v[j + 3] *= s4.w;
        }
      }
      if (epi_has<kEpi, kEpiBias>(ep.col_bias != nullptr)) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.col_bias + c0 + j));
          v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
        }
      }
    } else if constexpr (kRt) {
#pragma unroll 4
      for (int j = 0; j < 32; ++j) {
        if (c0 + j < shp.N) {
          if (rn_on) v[j] = rn_rstd * (v[j] - rn_mean * ep.rownorm_u[c0 + j]);
          if (ep.col_scale) v[j] *= ep.col_scale[c0 + j];
          if (ep.col_bias) v[j] += ep.col_bias[c0 + j];
        }
      }
    }
    if (epi_has<kEpi, kEpiAlpha>(ep.alpha_cols > 0)) {
      if (ep.alpha_cols >= c0 + 32) {  // warp-uniform: the whole chunk is scaled (q columns)
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= ep.alpha;
      } else if (ep.alpha_cols > c0) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < ep.alpha_cols) v[j] *= ep.alpha;
      }
    }
    if (epi_has<kEpi, kEpiGelu>(ep.act == SGF_ACT_GELU)) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float2 y = gelu_erf2(make_float2(v[j], v[j + 1]));
        v[j] = y.x;
        v[j + 1] = y.y;
      }
    }
    float4* dst = reinterpret_cast<float4*>(stage + lane * kRowPitch + ch * 128);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  __syncwarp();

  // ---- phase 2a: row map + all residual loads of this warp issued back to back (one exposed latency) ----
  int out_rows[kIters];
  uint4 res_lo[kIters], res_hi[kIters];
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const int r = quarter * 32 + it * kRowsPerIter + lane / kLanesPerRow;
    int64_t out_row;
    bool row_ok;
    if constexpr (kConv) {
      const int hh = h0 + r / shp.bw, ww = w0 + r % shp.bw;
      row_ok = (hh < shp.H) && (ww < shp.W);
      out_row = (static_cast<int64_t>(img) * shp.H + hh) * shp.W + ww;
    } else {
      row_ok = (m0 + r) < shp.M;
      out_row = m0 + r;
    }
    out_rows[it] = (row_ok && c < shp.N) ? static_cast<int>(out_row) : -1;
    res_lo[it] = make_uint4(0, 0, 0, 0);
    res_hi[it] = make_uint4(0, 0, 0, 0);
    if ((res_f32 || res_b16) && vec_ok && out_rows[it] >= 0) {
      const uint8_t* rptr = reinterpret_cast<const uint8_t*>(ep.residual) +
                            (static_cast<int64_t>(z) * ep.r_batch_stride + out_row * ep.ldr + c) * rsz;
      res_lo[it] = *reinterpret_cast<const uint4*>(rptr);
      if (res_f32) res_hi[it] = *reinterpret_cast<const uint4*>(rptr + 16);
    }
  }
  // ---- phase 2b ----
  const int scol = (lane % kLanesPerRow) * 8;  // column inside this warp's staging tile
  const bool want_stats = epi_has<kEpi, kEpiRowStats>(ep.rowstats_out != nullptr);
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    float st_sum = 0.f, st_sq = 0.f;
    if (out_rows[it] >= 0) {
      const int rl = it * kRowsPerIter + lane / kLanesPerRow;
      float v[8];
      {
        const float4 a4 = *reinterpret_cast<const float4*>(stage + rl * kRowPitch + scol * 4);
        const float4 b4 = *reinterpret_cast<const float4*>(stage + rl * kRowPitch + scol * 4 + 16);
        v[0] = a4.x; v[1] = a4.y; v[2] = a4.z; v[3] = a4.w; v[4] = b4.x; v[5] = b4.y; v[6] = b4.z; v[7] = b4.w;
      }
      uint8_t* cptr = reinterpret_cast<uint8_t*>(ep.c) +
                      (static_cast<int64_t>(z) * ep.c_batch_stride + static_cast<int64_t>(out_rows[it]) * ep.ldc + c) * csz;
      if (vec_ok) {
        if (res_f32) {
          v[0] += __uint_as_float(res_lo[it].x); v[1] += __uint_as_float(res_lo[it].y);
          v[2] += __uint_as_float(res_lo[it].z); v[3] += __uint_as_float(res_lo[it].w);
          v[4] += __uint_as_float(res_hi[it].x); v[5] += __uint_as_float(res_hi[it].y);
          v[6] += __uint_as_float(res_hi[it].z); v[7] += __uint_as_float(res_hi[it].w);
        } else if (res_b16) {
          const float2 a = unpack_bf16x2(res_lo[it].x), b2 = unpack_bf16x2(res_lo[it].y),
                       c2 = unpack_bf16x2(res_lo[it].z), d = unpack_bf16x2(res_lo[it].w);
          v[0] += a.x; v[1] += a.y; v[2] += b2.x; v[3] += b2.y; v[4] += c2.x; v[5] += c2.y; v[6] += d.x; v[7] += d.y;
        }
        if (do_relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        if constexpr (!kRt && (kEpi & kEpiAtomic) != 0) {  // split-K partial: fp32 reduction into C
          red_add_f32x4(reinterpret_cast<float*>(cptr), v[0], v[1], v[2], v[3]);
          red_add_f32x4(reinterpret_cast<float*>(cptr) + 4, v[4], v[5], v[6], v[7]);
        } else if (out_f32) {
          *reinterpret_cast<float4*>(cptr) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(cptr + 16) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
          uint4 o;
          o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
          o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
          *reinterpret_cast<uint4*>(cptr) = o;
          if (want_stats) {  // statistics of exactly what was stored
            const float2 a = unpack_bf16x2(o.x), b2 = unpack_bf16x2(o.y), c2 = unpack_bf16x2(o.z), d = unpack_bf16x2(o.w);
            const float w[8] = {a.x, a.y, b2.x, b2.y, c2.x, c2.y, d.x, d.y};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              st_sum += w[j];
              st_sq = fmaf(w[j], w[j], st_sq);
            }
          }
        }
      } else if constexpr (kRt) {
        const uint8_t* rptr = ep.residual ? reinterpret_cast<const uint8_t*>(ep.residual) +
                                                (static_cast<int64_t>(z) * ep.r_batch_stride +
                                                 static_cast<int64_t>(out_rows[it]) * ep.ldr + c) * rsz
                                          : nullptr;
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          if (c + j < shp.N) {
            float t = v[j];
            if (rptr)
              t += res_f32 ? reinterpret_cast<const float*>(rptr)[j]
                           : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rptr)[j]);
            if (do_relu) t = fmaxf(t, 0.0f);
            if (out_f32)
              reinterpret_cast<float*>(cptr)[j] = t;
            else
              reinterpret_cast<__nv_bfloat16*>(cptr)[j] = __float2bfloat16_rn(t);
          }
        }
      }
    }  // valid row
    if (want_stats) {  // warp-uniform: reduce over the lanes that share a row, one slot per row segment
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {  // 8 lanes = one 64-column block of a row
        st_sum += __shfl_xor_sync(0xffffffffu, st_sum, o);
        st_sq += __shfl_xor_sync(0xffffffffu, st_sq, o);
      }
      if ((lane & 7) == 0 && out_rows[it] >= 0) {
        // deterministic: one (sum, sumsq) slot per 64-column block of the row, summed in order by the consumer
        const int nparts = (shp.N + 63) / 64;
        float2* sp = reinterpret_cast<float2*>(ep.rowstats_out) +
                     (static_cast<int64_t>(z) * shp.M + out_rows[it]) * nparts + c / 64;
        *sp = make_float2(st_sum, st_sq);
      }
    }
  }
}

template <int BN, int kStages, bool kConv, int kEpi>
__global__ void __launch_bounds__(gemm_threads<BN>(), (BN <= 128 ? 2 : 1)) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmB,
                                                                    const GemmShape shp, const GemmEpilogue ep) {
  using S = GemmSmem<BN>;
  constexpr int kEpiWarps = gemm_epi_warps<BN>();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem is only guaranteed 16B aligned: round up to the 1024B the 128B swizzle needs
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* accum_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int mt = blockIdx.y;
  const int z = blockIdx.z;
  const int num_kb = (shp.K + BK - 1) / BK;

  // conv tile decomposition
  int img = 0, h0 = 0, w0 = 0;
  if constexpr (kConv) {
    const int per_img = shp.tiles_w * shp.tiles_h;
    img = mt / per_img;
    const int t = mt - img * per_img;
    h0 = (t / shp.tiles_w) * shp.bh;
    w0 = (t % shp.tiles_w) * shp.bw;
  }
  const int m0 = mt * BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory from here on

  if (warp == 0) {
    if (lane == 0) {
#pragma unroll 1
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * S::kStageBytes;
        uint8_t* sb = sa + S::kABytes;
        mbar_expect_tx(&full_bar[s], S::kStageBytes);
        if constexpr (kConv) {
          const int tap = kb / shp.cin_blocks;
          const int kc = kb - tap * shp.cin_blocks;
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          tma_load_4d(sa, &tmA, &full_bar[s], kc * BK, w0 + dx, h0 + dy, img);
        } else {
          tma_load_3d(sa, &tmA, &full_bar[s], kb * BK, m0, z);
        }
        tma_load_3d(sb, &tmB, &full_bar[s], kb * BK, n0, z);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
#pragma unroll 1
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * S::kStageBytes);
        const uint32_t sb = sa + S::kABytes;
        const uint64_t da = make_smem_desc_sw128(sa);
        const uint64_t db = make_smem_desc_sw128(sb);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 K-elements = 32 bytes inside the 128B swizzle row: +2 in the (addr>>4) field
          umma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees this smem stage when the MMAs above retire
      }
      umma_commit(accum_bar);  // accumulator complete
    }
  } else {
    gemm_epilogue<BN, kEpiWarps, kStages * S::kStageBytes, kConv, kEpi>(smem, tmem_base, accum_bar, shp, ep, m0, n0, z,
                                                                        img, h0, w0, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------------
// Mixed-major kernel for the dense adjoints: either operand may be MN-major, i.e. stored with the
// CONTRACTION index as the row index (A[k][m] / B[k][n]), which is how the saved activations and
// output gradients lie in HBM for dW[n,k] = sum_t dY[t,n] X[t,k] (contraction over tokens) and how
// W[n,k] lies for dX = dY W (contraction over n).  No transposed copies are materialised: an
// MN-major tile is staged as 64-column TMA boxes ([64 contraction rows] x [64 MN elements], 128B
// swizzle) and consumed through an MN-major UMMA descriptor (LBO = distance between the 64-wide
// column blocks, SBO = 8-row group stride).  blockIdx.z = split-K slice; slices reduce into C with
// vector fp32 atomics (kEpiAtomic).
// ----------------------------------------------------------------------------------------
SGF_DEVICE uint64_t make_smem_desc_sw128_mn(uint32_t smem_addr, uint32_t block_stride_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(block_stride_bytes >> 4) << 16;  // LBO: next 64-element block along M/N
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                // SBO: next group of 8 contraction rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int BN, int kStages, bool kAmn, bool kBmn, int kEpi>
__global__ void __launch_bounds__(gemm_threads<BN>(), (BN <= 128 ? 2 : 1)) gemm_mm_tcgen05_kernel(
    const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmShape shp,
    const GemmEpilogue ep, const int kb_per_split) {
  using S = GemmSmem<BN>;
  constexpr int kEpiWarps = gemm_epi_warps<BN>();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* accum_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int num_kb_all = (shp.K + BK - 1) / BK;
  const int kb0 = blockIdx.z * kb_per_split;
  const int num_kb = min(kb_per_split, num_kb_all - kb0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
#pragma unroll 1
      for (int i = 0; i < num_kb; ++i) {
        const int kb = kb0 + i;
        const int s = i % kStages;
        const uint32_t ph = (i / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * S::kStageBytes;
        uint8_t* sb = sa + S::kABytes;
        mbar_expect_tx(&full_bar[s], S::kStageBytes);
        if constexpr (kAmn) {
#pragma unroll
          for (int blk = 0; blk < BM / 64; ++blk) tma_load_2d(sa + blk * 8192, &tmA, &full_bar[s], m0 + blk * 64, kb * BK);
        } else {
          tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, m0);
        }
        if constexpr (kBmn) {
#pragma unroll
          for (int blk = 0; blk < BN / 64; ++blk) tma_load_2d(sb + blk * 8192, &tmB, &full_bar[s], n0 + blk * 64, kb * BK);
        } else {
          tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, kAmn ? 1 : 0, kBmn ? 1 : 0);
#pragma unroll 1
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % kStages;
        const uint32_t ph = (i / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * S::kStageBytes);
        const uint32_t sb = sa + S::kABytes;
        const uint64_t da = kAmn ? make_smem_desc_sw128_mn(sa, 8192) : make_smem_desc_sw128(sa);
        const uint64_t db = kBmn ? make_smem_desc_sw128_mn(sb, 8192) : make_smem_desc_sw128(sb);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major: 16 contraction elements = 32 B inside the swizzle row (+2); MN-major: 16 rows of 128 B (+128)
          umma_f16(tmem_base, da + (kAmn ? 128 : 2) * k, db + (kBmn ? 128 : 2) * k, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(accum_bar);
    }
  } else {
    gemm_epilogue<BN, kEpiWarps, kStages * S::kStageBytes, false, kEpi>(smem, tmem_base, accum_bar, shp, ep, m0, n0, 0, 0,
                                                                        0, 0, warp, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
template <int BN, int kStages>
static constexpr int gemm_smem_bytes() {
  return kStages * GemmSmem<BN>::kStageBytes + (2 * kStages + 1) * 8 + 16 + 1024;
}

template <int BN, int kStages, bool kConv, int kEpi>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& shp, const GemmEpilogue& ep,
                       dim3 grid, cudaStream_t st) {
  auto kern = gemm_tcgen05_kernel<BN, kStages, kConv, kEpi>;
  constexpr int smem = gemm_smem_bytes<BN, kStages>();
  static bool configured = false;
  if (!configured) {
    SGF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  SGF_CHECK_CUDA(launch_pdl(kern, grid, dim3(gemm_threads<BN>()), smem, st, tmA, tmB, shp, ep));
  count_launch();
  return SGF_OK;
}

#define SGF_EPI_CASE(MASK)                                                                              \
  case (MASK):                                                                                          \
    return launch_gemm<BN, kStages, kConv, (MASK)>(tmA, tmB, shp, ep, grid, st);

template <int BN, int kStages, bool kConv>
static int dispatch_epilogue(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& shp, const GemmEpilogue& ep,
                             dim3 grid, cudaStream_t st) {
  if (shp.N % 32 == 0) {
    switch (epilogue_mask(ep)) {
      SGF_EPI_CASE(kEpiScale | kEpiBias | kEpiRelu)                 // stem conv + BN + ReLU
      SGF_EPI_CASE(kEpiScale | kEpiBias | kEpiRelu | kEpiResBf16)   // bottleneck conv3 + BN + residual + ReLU
      SGF_EPI_CASE(kEpiScale | kEpiBias)                            // downsample conv + BN
      default: break;
    }
    if constexpr (!kConv) {
      switch (epilogue_mask(ep)) {
        SGF_EPI_CASE(kEpiBias | kEpiAlpha)                          // fused QKV / cross q
        SGF_EPI_CASE(kEpiBias | kEpiOutF32)                         // out_proj, image_proj
        SGF_EPI_CASE(kEpiBias)                                      // cross k/v, position projections
        SGF_EPI_CASE(kEpiBias | kEpiGelu | kEpiRowStats)            // fc1 (+ ffn_layernorm statistics)
        SGF_EPI_CASE(kEpiBias | kEpiGelu)                           // fc1
        SGF_EPI_CASE(kEpiBias | kEpiRowNorm | kEpiResF32 | kEpiOutF32)  // fc2 with folded ffn_layernorm
        SGF_EPI_CASE(kEpiBias | kEpiResF32 | kEpiOutF32)            // fc2
        SGF_EPI_CASE(0)                                             // dX = dY W (bf16)
        SGF_EPI_CASE(kEpiOutF32)                                    // dW = dY^T X (fp32), d(encoder_out)
        SGF_EPI_CASE(kEpiOutF32 | kEpiResF32)                       // dW += (gradient accumulation)
        default: break;
      }
    }
  }
  return launch_gemm<BN, kStages, kConv, kEpiRuntime>(tmA, tmB, shp, ep, grid, st);
}

static int check_epilogue_alignment(const GemmEpilogue& ep, int N) {
  if (N % 8 == 0) {
    const int cs = ep.c_dtype == SGF_F32 ? 4 : 2;
    SGF_REQUIRE((reinterpret_cast<uintptr_t>(ep.c) % 16) == 0 && (ep.ldc * cs) % 16 == 0 &&
                    (ep.c_batch_stride * cs) % 16 == 0,
                "gemm: C must be 16-byte aligned with 16-byte multiple row stride");
    if (ep.residual) {
      const int rs = ep.r_dtype == SGF_F32 ? 4 : 2;
      SGF_REQUIRE((reinterpret_cast<uintptr_t>(ep.residual) % 16) == 0 && (ep.ldr * rs) % 16 == 0 &&
                      (ep.r_batch_stride * rs) % 16 == 0,
                  "gemm: residual must be 16-byte aligned with 16-byte multiple row stride");
    }
    if (ep.col_scale) SGF_REQUIRE((reinterpret_cast<uintptr_t>(ep.col_scale) % 16) == 0, "gemm: col_scale align");
    if (ep.col_bias) SGF_REQUIRE((reinterpret_cast<uintptr_t>(ep.col_bias) % 16) == 0, "gemm: col_bias align");
  }
  return SGF_OK;
}

// Kernel-family selection can be pinned from the environment for A/B measurements (tools/bench_gemm.py):
//   SGF_GEMM_FAMILY = tile (one 128 x BN tile per CTA, this file) | ts (persistent CTA pair, TMA-store epilogue: gemm_ts.cu)
//   unset = ts wherever it applies.  (The r01/r02 persistent kernels with direct TMEM -> global stores were removed after
//   the A/B in profiles/r02_gemm_family_ab.txt: their epilogue was the bottleneck.)
enum { kFamAuto = 0, kFamTile = 1, kFamTs = 4 };
static int gemm_family_override() {
  const char* e = getenv("SGF_GEMM_FAMILY");
  if (!e) return kFamAuto;
  if (!strcmp(e, "tile")) return kFamTile;
  if (!strcmp(e, "ts")) return kFamTs;
  return kFamAuto;
}

static int pick_bn(int M_tiles, int N, int batch) {
  // prefer the widest tile that still gives >= 1 wave of CTAs on 148 SMs
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  const long ctas128 = static_cast<long>(M_tiles) * ((N + 127) / 128) * batch;
  if (N % 128 != 0 && N % 64 == 0 && N < 256) return 64;
  if (ctas128 < 148) return 64;
  return 128;
}

}  // namespace sgf

using namespace sgf;

extern "C" int sgf_gemm_bf16(const sgf_gemm_args* a, void* stream) {
  SGF_REQUIRE(a != nullptr, "gemm: null args");
  SGF_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->batch > 0, "gemm: bad shape M=%d N=%d K=%d batch=%d", a->M,
              a->N, a->K, a->batch);
  SGF_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0 && a->a_batch_stride % 8 == 0 && a->b_batch_stride % 8 == 0,
              "gemm: lda/ldb/batch strides must be multiples of 8 elements (TMA 16-byte strides)");
  SGF_REQUIRE((reinterpret_cast<uintptr_t>(a->a) % 16) == 0 && (reinterpret_cast<uintptr_t>(a->b) % 16) == 0,
              "gemm: A/B must be 16-byte aligned");
  SGF_REQUIRE(a->lda >= a->K && a->ldb >= a->K, "gemm: leading dimension smaller than K");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  GemmShape shp{};
  shp.M = a->M; shp.N = a->N; shp.K = a->K;
  GemmEpilogue ep{a->c, a->ldc, a->c_batch_stride, a->c_dtype, a->col_scale, a->col_bias, a->residual,
                  a->ldr, a->r_batch_stride, a->r_dtype, a->act, a->alpha, a->alpha_cols, a->rowstats_out,
                  a->rownorm_stats, a->rownorm_u, a->rownorm_dim > 0 ? 1.0f / static_cast<float>(a->rownorm_dim) : 0.f,
                  (a->rownorm_dim + 63) / 64};
  SGF_REQUIRE(!a->rowstats_out || (a->c_dtype == SGF_BF16 && a->N % 64 == 0 && a->N >= 128),
              "gemm: rowstats_out needs a bf16 output with N a multiple of 64 (>= 128)");
  SGF_REQUIRE(!a->rownorm_stats || (a->rownorm_u && a->rownorm_dim > 0 && a->rownorm_dim % 128 == 0),
              "gemm: rownorm_stats needs rownorm_u and rownorm_dim (a multiple of 128)");
  if (int rc = check_epilogue_alignment(ep, a->N)) return rc;

  const int m_tiles = (a->M + BM - 1) / BM;
  int bn = pick_bn(m_tiles, a->N, a->batch);
  const int fam = gemm_family_override();
  // persistent CTA-pair kernel with the TMA-store epilogue: everything with N % 64 == 0 and at least a few tiles
  if (a->batch == 1 && (fam == kFamTs || (fam == kFamAuto && static_cast<long>(a->M) * a->N >= 128L * 1024))) {
    const int rc = gemm_ts_dispatch(shp, ep, a->a, a->lda, a->b, a->ldb, nullptr, st);
    if (rc >= 0) return rc;
  }
  if (a->rowstats_out && bn != 64 && bn != 128) bn = 128;  // statistics are kept per 64-column block

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->M), static_cast<uint64_t>(a->batch)};
    uint64_t strides[2] = {static_cast<uint64_t>(a->lda) * 2,
                           static_cast<uint64_t>(a->batch > 1 ? a->a_batch_stride : a->lda * (int64_t)a->M) * 2};
    uint32_t box[3] = {BK, BM, 1};
    if (int rc = encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->a, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->N), static_cast<uint64_t>(a->batch)};
    uint64_t strides[2] = {static_cast<uint64_t>(a->ldb) * 2,
                           static_cast<uint64_t>(a->batch > 1 ? a->b_batch_stride : a->ldb * (int64_t)a->N) * 2};
    uint32_t box[3] = {BK, static_cast<uint32_t>(bn), 1};
    if (int rc = encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->b, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  dim3 grid((a->N + bn - 1) / bn, m_tiles, a->batch);
  switch (bn) {
    case 32: return dispatch_epilogue<32, 4, false>(tmA, tmB, shp, ep, grid, st);
    case 64: return dispatch_epilogue<64, 4, false>(tmA, tmB, shp, ep, grid, st);
    default: return dispatch_epilogue<128, 3, false>(tmA, tmB, shp, ep, grid, st);
  }
}


template <int BN, bool kAmn, bool kBmn, int kEpi>
static int launch_gemm_mm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& shp, const GemmEpilogue& ep,
                          dim3 grid, int kb_per_split, cudaStream_t st) {
  constexpr int kStages = BN <= 64 ? 4 : 3;
  auto kern = gemm_mm_tcgen05_kernel<BN, kStages, kAmn, kBmn, kEpi>;
  constexpr int smem = gemm_smem_bytes<BN, kStages>();
  static bool configured = false;
  if (!configured) {
    SGF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  SGF_CHECK_CUDA(launch_pdl(kern, grid, dim3(gemm_threads<BN>()), smem, st, tmA, tmB, shp, ep, kb_per_split));
  count_launch();
  return SGF_OK;
}

template <int BN, bool kAmn, bool kBmn>
static int dispatch_gemm_mm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmShape& shp, const GemmEpilogue& ep,
                            dim3 grid, int kb_per_split, bool atomic, cudaStream_t st) {
  if (atomic) return launch_gemm_mm<BN, kAmn, kBmn, kEpiOutF32 | kEpiAtomic>(tmA, tmB, shp, ep, grid, kb_per_split, st);
  if (ep.c_dtype == SGF_F32) return launch_gemm_mm<BN, kAmn, kBmn, kEpiOutF32>(tmA, tmB, shp, ep, grid, kb_per_split, st);
  return launch_gemm_mm<BN, kAmn, kBmn, 0>(tmA, tmB, shp, ep, grid, kb_per_split, st);
}

extern "C" int sgf_gemm_bf16_ex(const sgf_gemm_args* a, int32_t a_mn_major, int32_t b_mn_major, int32_t split_k,
                                void* stream) {
  SGF_REQUIRE(a != nullptr && a->a && a->b && a->c, "gemm_ex: null pointer");
  SGF_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->batch == 1, "gemm_ex: bad shape (batch must be 1)");
  SGF_REQUIRE(!a->col_scale && !a->col_bias && a->act == SGF_ACT_NONE && a->alpha_cols == 0 && !a->rowstats_out &&
                  !a->rownorm_stats && (!a->residual || (a->residual == a->c && a->r_dtype == SGF_F32 && a->c_dtype == SGF_F32)),
              "gemm_ex: the mixed-major kernel has a plain epilogue (store, or fp32 accumulate in place: residual == c)");
  SGF_REQUIRE(a->N % 32 == 0, "gemm_ex: N must be a multiple of 32 (N=%d)", a->N);
  SGF_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0 && (reinterpret_cast<uintptr_t>(a->a) % 16) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->b) % 16) == 0,
              "gemm_ex: operand alignment");
  SGF_REQUIRE(a->lda >= (a_mn_major ? a->M : a->K) && a->ldb >= (b_mn_major ? a->N : a->K), "gemm_ex: leading dimension");
  if (split_k < 1) split_k = 1;
  SGF_REQUIRE(split_k == 1 || a->c_dtype == SGF_F32, "gemm_ex: split-K needs an fp32 output (accumulated in place)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GemmShape shp{};
  shp.M = a->M; shp.N = a->N; shp.K = a->K;
  GemmEpilogue ep{a->c, a->ldc, 0, a->c_dtype, nullptr, nullptr, nullptr, 0, 0, SGF_BF16, SGF_ACT_NONE, 1.0f, 0,
                  nullptr, nullptr, nullptr, 0.f, 0};
  if (int rc = check_epilogue_alignment(ep, a->N)) return rc;
  const int m_tiles = (a->M + BM - 1) / BM;
  const int bn = (a->N % 128 == 0 || a->N > 128) && (static_cast<long>(m_tiles) * ((a->N + 127) / 128) * split_k >= 96) ? 128 : 64;
  auto make_map = [&](CUtensorMap* tm, const void* base, int64_t ld, int rows_mn, bool mn_major, int box_mn) -> int {
    if (mn_major) {  // [K rows][MN contiguous]: box = 64 MN elements x 64 contraction rows
      uint64_t dims[2] = {static_cast<uint64_t>(rows_mn), static_cast<uint64_t>(a->K)};
      uint64_t strides[1] = {static_cast<uint64_t>(ld) * 2};
      uint32_t box[2] = {64, BK};
      return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    }
    uint64_t dims[2] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(rows_mn)};
    uint64_t strides[1] = {static_cast<uint64_t>(ld) * 2};
    uint32_t box[2] = {BK, static_cast<uint32_t>(box_mn)};
    return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  CUtensorMap tmA, tmB;
  if (int rc = make_map(&tmA, a->a, a->lda, a->M, a_mn_major != 0, BM)) return rc;
  if (int rc = make_map(&tmB, a->b, a->ldb, a->N, b_mn_major != 0, bn)) return rc;
  const int num_kb = (a->K + BK - 1) / BK;
  if (split_k > num_kb) split_k = num_kb;
  const int kb_per_split = (num_kb + split_k - 1) / split_k;
  split_k = (num_kb + kb_per_split - 1) / kb_per_split;
  dim3 grid((a->N + bn - 1) / bn, m_tiles, split_k);
  const bool atomic = split_k > 1 || a->residual != nullptr;  // residual == c: accumulate in place
#define SGF_MM_CASE(AMN, BMN)                                                                                   \
  if ((a_mn_major != 0) == AMN && (b_mn_major != 0) == BMN)                                                      \
    return bn == 128 ? dispatch_gemm_mm<128, AMN, BMN>(tmA, tmB, shp, ep, grid, kb_per_split, atomic, st)        \
                     : dispatch_gemm_mm<64, AMN, BMN>(tmA, tmB, shp, ep, grid, kb_per_split, atomic, st);
  SGF_MM_CASE(true, true)
  SGF_MM_CASE(false, true)
#undef SGF_MM_CASE
  set_last_error("gemm_ex: only (A MN-major, B MN-major) and (A K-major, B MN-major) are built");
  return SGF_ERR_UNSUPPORTED;
}

extern "C" int sgf_conv3x3_s1_nhwc(const sgf_conv3x3_args* a, void* stream) {
  SGF_REQUIRE(a != nullptr, "conv3x3: null args");
  SGF_REQUIRE(a->cin % 64 == 0, "conv3x3: Cin must be a multiple of 64 (got %d)", a->cin);
  SGF_REQUIRE(a->cout % 8 == 0, "conv3x3: Cout must be a multiple of 8 (got %d)", a->cout);
  SGF_REQUIRE(a->n > 0 && a->h > 0 && a->w_ > 0, "conv3x3: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  // 128-pixel tile = bw x bh rectangle; bw = smallest power of two >= min(W,128)
  int bw = 1;
  while (bw < a->w_ && bw < 128) bw <<= 1;
  const int bh = 128 / bw;
  GemmShape shp{};
  shp.M = a->n * a->h * a->w_;
  shp.N = a->cout;
  shp.K = 9 * a->cin;
  shp.H = a->h; shp.W = a->w_; shp.bw = bw; shp.bh = bh;
  shp.tiles_w = (a->w_ + bw - 1) / bw;
  shp.tiles_h = (a->h + bh - 1) / bh;
  shp.cin_blocks = a->cin / 64;
  shp.kw = 3; shp.stride_h = shp.stride_w = 1; shp.pad_h = shp.pad_w = 1;
  GemmEpilogue ep{a->y, a->cout, 0, SGF_BF16, a->col_scale, a->col_bias, nullptr, 0, 0, SGF_BF16, a->act, 1.0f, 0,
                  nullptr, nullptr, nullptr, 0.f, 0};
  if (int rc = check_epilogue_alignment(ep, a->cout)) return rc;

  const int m_tiles = a->n * shp.tiles_w * shp.tiles_h;
  const int fam = gemm_family_override();
  if (fam == kFamTs || fam == kFamAuto) {
    const ConvInput cv{a->x, a->n, a->h, a->w_, a->cin, a->cin, static_cast<int64_t>(a->cin) * a->w_,
                       static_cast<int64_t>(a->cin) * a->w_ * a->h};
    const int rc = gemm_ts_dispatch(shp, ep, nullptr, 0, a->w, shp.K, &cv, st);
    if (rc >= 0) return rc;
  }
  const int bn = pick_bn(m_tiles, a->cout, 1);
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(a->cin), static_cast<uint64_t>(a->w_), static_cast<uint64_t>(a->h),
                        static_cast<uint64_t>(a->n)};
    uint64_t strides[3] = {static_cast<uint64_t>(a->cin) * 2, static_cast<uint64_t>(a->cin) * a->w_ * 2,
                           static_cast<uint64_t>(a->cin) * a->w_ * a->h * 2};
    uint32_t box[4] = {BK, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), 1};
    if (int rc = encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->x, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  {
    uint64_t dims[3] = {static_cast<uint64_t>(shp.K), static_cast<uint64_t>(a->cout), 1};
    uint64_t strides[2] = {static_cast<uint64_t>(shp.K) * 2, static_cast<uint64_t>(shp.K) * a->cout * 2};
    uint32_t box[3] = {BK, static_cast<uint32_t>(bn), 1};
    if (int rc = encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a->w, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  }
  dim3 grid((a->cout + bn - 1) / bn, m_tiles, 1);
  switch (bn) {
    case 32: return dispatch_epilogue<32, 4, true>(tmA, tmB, shp, ep, grid, st);
    case 64: return dispatch_epilogue<64, 4, true>(tmA, tmB, shp, ep, grid, st);
    default: return dispatch_epilogue<128, 3, true>(tmA, tmB, shp, ep, grid, st);
  }
}

extern "C" int sgf_conv2d_nhwc(const sgf_conv2d_args* a, void* stream) {
  SGF_REQUIRE(a != nullptr && a->x && a->w && a->y, "conv2d: null pointer");
  SGF_REQUIRE(a->cin % 64 == 0, "conv2d: the innermost input extent must be a multiple of 64 (got %d)", a->cin);
  SGF_REQUIRE(a->cout % 64 == 0, "conv2d: Cout must be a multiple of 64 (got %d)", a->cout);
  SGF_REQUIRE(a->n > 0 && a->h > 0 && a->w_ > 0 && a->ho > 0 && a->wo > 0 && a->kh > 0 && a->kw > 0 && a->stride_h > 0 &&
                  a->stride_w > 0 && a->pad_h >= 0 && a->pad_w >= 0,
              "conv2d: bad shape");
  SGF_REQUIRE(a->x_w_stride % 8 == 0 && a->x_h_stride % 8 == 0 && a->x_n_stride % 8 == 0 &&
                  reinterpret_cast<uintptr_t>(a->x) % 16 == 0 && reinterpret_cast<uintptr_t>(a->w) % 16 == 0,
              "conv2d: input strides must be multiples of 8 elements and x / w 16-byte aligned (TMA)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // 128-pixel output tile = bw x bh rectangle; bw = smallest power of two >= min(Wo, 128)
  int bw = 1;
  while (bw < a->wo && bw < 128) bw <<= 1;
  const int bh = 128 / bw;
  SGF_REQUIRE(bw * a->stride_w <= 256 && bh * a->stride_h <= 256, "conv2d: stride too large for one TMA box");
  GemmShape shp{};
  shp.M = a->n * a->ho * a->wo;
  shp.N = a->cout;
  shp.K = a->kh * a->kw * a->cin;
  shp.H = a->ho; shp.W = a->wo; shp.bw = bw; shp.bh = bh;
  shp.tiles_w = (a->wo + bw - 1) / bw;
  shp.tiles_h = (a->ho + bh - 1) / bh;
  shp.cin_blocks = a->cin / 64;
  shp.kw = a->kw; shp.stride_h = a->stride_h; shp.stride_w = a->stride_w; shp.pad_h = a->pad_h; shp.pad_w = a->pad_w;
  GemmEpilogue ep{a->y, a->cout, 0, SGF_BF16, a->col_scale, a->col_bias, nullptr, 0, 0, SGF_BF16, a->act, 1.0f, 0,
                  nullptr, nullptr, nullptr, 0.f, 0};
  if (int rc = check_epilogue_alignment(ep, a->cout)) return rc;
  const ConvInput cv{a->x, a->n, a->h, a->w_, a->cin, a->x_w_stride, a->x_h_stride, a->x_n_stride};
  const int rc = gemm_ts_dispatch(shp, ep, nullptr, 0, a->w, shp.K, &cv, st);
  if (rc >= 0) return rc;
  set_last_error("conv2d: no kernel for this epilogue (supported: BN affine with or without ReLU)");
  return SGF_ERR_UNSUPPORTED;
}
