// HBM-bound row / layout kernels: fused LayerNorm(+gather, +residual, +second LayerNorm),
// attention rel-pos bias assembly, stem layout helpers (NCHW->NHWC, zero-padded NHWC8 for the 7x7/2 stem convolution; the strided
// convolutions, max-pool).  Warp-shuffle reductions, 128-bit global accesses.
#include <stdlib.h>

#include <cuda_fp16.h>

#include "common.cuh"

namespace sgf {

// ----------------------------------------------------------------------------------------
// fused row LayerNorm: one warp per row, row kept in registers (CH chunks of 8 per lane)
// ----------------------------------------------------------------------------------------
struct RowLnParams {
  const void* x; int64_t ldx; int x_dtype;
  const int64_t* gather_idx;
  const float* pre_add;
  const float* g1; const float* b1;
  const void* residual; int64_t ldr; int r_dtype;
  void* out1; int64_t ld1; int out1_dtype;
  const float* g2; const float* b2;
  void* out2; int64_t ld2;
  const uint8_t* zero_row;
  int rows, D;
  int seg_len, seg_stride, seg_off;
  float* clear_rowstats;
  int x_act;
  float drop_p, droppath_p; uint32_t drop_seed, drop_site; int rows_per_sample; const int32_t* drop_step;
};

SGF_DEVICE void load8(const void* base, int dtype, int64_t elem_off, float (&v)[8]) {
  if (dtype == SGF_F32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
    const float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
}
SGF_DEVICE void store8(void* base, int dtype, int64_t elem_off, const float (&v)[8]) {
  if (dtype == SGF_F32) {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem_off);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + elem_off) = u;
  }
}

// Rows are staged through shared memory with 1-D bulk async copies (cp.async.bulk + mbarrier):
// the bytes in flight per SM are bounded by shared memory (4 CTAs x 48 KB), not by registers, which
// is what an HBM-latency-bound row kernel needs.  One warp then owns one row and makes a few
// conflict-free 128-bit passes over it; the only cross-lane traffic is 2 shuffle reductions per LN.
static constexpr int kLnMaxRows = 8;  // rows (= warps) per CTA (fewer when a row needs > 25 KB of staging)

SGF_DEVICE void smem_load8(const uint8_t* row, int dtype, int e, float (&v)[8]) {
  if (dtype == SGF_F32) {
    const float4 a = *reinterpret_cast<const float4*>(row + e * 4);
    const float4 b = *reinterpret_cast<const float4*>(row + e * 4 + 16);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(row + e * 2);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
}

__global__ void __launch_bounds__(kLnMaxRows * 32) row_layernorm_kernel(const RowLnParams p, const int x_row_bytes,
                                                                        const int r_row_bytes, const int s_row_bytes) {
  // Persistent CTAs; every warp owns one row slot and runs its own two-deep pipeline: while it works on the row of
  // group g it has the bulk copies of its row of group g + gridDim.x in flight (per-warp mbarriers: no CTA-wide
  // synchronisation anywhere in the loop).
  pdl_trigger();
  const int kLnRows = blockDim.x >> 5;
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int in_bytes = x_row_bytes + r_row_bytes;
  uint8_t* inbuf = ln_smem;                                              // [2][kLnRows][x | residual]
  uint8_t* sbuf = inbuf + 2 * kLnRows * in_bytes;                        // [kLnRows][s_row_bytes] (fp32 stash for LN2)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbuf + kLnRows * s_row_bytes);  // [2][kLnRows]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int xs = p.x_dtype == SGF_F32 ? 4 : 2, rs = p.r_dtype == SGF_F32 ? 4 : 2;
  const int ngroups = (p.rows + kLnRows - 1) / kLnRows;
  if (lane == 0) {
    mbar_init(&bars[warp], 1);
    mbar_init(&bars[kLnRows + warp], 1);
    fence_mbar_init();
  }
  __syncwarp();
  pdl_wait();
  const float invD = 1.0f / static_cast<float>(p.D);
  const int nchunk = p.D >> 3;
  const DropCtx drop = make_drop_ctx(p.drop_p, p.droppath_p, p.drop_seed, p.drop_site, p.drop_step, p.rows_per_sample);
  const bool two_stage = p.out2 && (p.g1 || p.residual || p.pre_add || p.out1 || p.x_act || drop.on);
  float* sr = reinterpret_cast<float*>(sbuf + warp * s_row_bytes);

  auto dst_of = [&](int row) -> int64_t {
    return p.seg_len > 0 ? static_cast<int64_t>(row / p.seg_len) * p.seg_stride + p.seg_off + row % p.seg_len : row;
  };
  auto issue = [&](int g, int buf) {  // lane 0: bulk copies of this warp's row of group g into buffer `buf`
    const int row = g * kLnRows + warp;
    if (row >= p.rows) return;
    uint8_t* dst = inbuf + (buf * kLnRows + warp) * in_bytes;
    uint64_t* bar = &bars[buf * kLnRows + warp];
    mbar_expect_tx(bar, in_bytes);
    const int64_t src_row = p.gather_idx ? p.gather_idx[row] : row;
    bulk_load_1d(dst, reinterpret_cast<const uint8_t*>(p.x) + src_row * p.ldx * xs, x_row_bytes, bar);
    if (p.residual)
      bulk_load_1d(dst + x_row_bytes, reinterpret_cast<const uint8_t*>(p.residual) + dst_of(row) * p.ldr * rs, r_row_bytes,
                   bar);
  };
  if (lane == 0 && static_cast<int>(blockIdx.x) < ngroups) issue(blockIdx.x, 0);

  int it = 0;
  for (int g = blockIdx.x; g < ngroups; g += gridDim.x, ++it) {
    const int buf = it & 1;
    __syncwarp();  // every lane is done with the other buffer (previous group) before lane 0 refills it
    if (lane == 0 && g + static_cast<int>(gridDim.x) < ngroups) {
      fence_proxy_async();  // generic-proxy reads of that buffer are ordered before the async-proxy bulk writes
      issue(g + gridDim.x, buf ^ 1);
    }
    const int row = g * kLnRows + warp;
    if (row >= p.rows) continue;  // (only in the last group)
    const int64_t dst_row = dst_of(row);
    if (p.clear_rowstats && lane < 2) p.clear_rowstats[static_cast<int64_t>(row) * 2 + lane] = 0.f;
    mbar_wait(&bars[buf * kLnRows + warp], (it >> 1) & 1);
    const uint8_t* xr = inbuf + (buf * kLnRows + warp) * in_bytes;
    const uint8_t* rr = xr + x_row_bytes;
    const bool zero = p.zero_row && p.zero_row[row];

  // ---- first LayerNorm statistics (over t = x + pre_add) ----
  float mean1 = 0.f, rstd1 = 1.f;
  if (p.g1) {
    float s = 0.f;
    for (int c = lane; c < nchunk; c += 32) {
      float v[8];
      smem_load8(xr, p.x_dtype, c * 8, v);
      if (p.x_act == SGF_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
      }
      if (p.pre_add) {
        float a[8];
        load8(p.pre_add, SGF_F32, c * 8, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += a[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[j];
    }
    mean1 = warp_sum(s) * invD;
    float q = 0.f;
    for (int c = lane; c < nchunk; c += 32) {
      float v[8];
      smem_load8(xr, p.x_dtype, c * 8, v);
      if (p.x_act == SGF_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
      }
      if (p.pre_add) {
        float a[8];
        load8(p.pre_add, SGF_F32, c * 8, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += a[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[j] - mean1;
        q += d * d;
      }
    }
    rstd1 = rsqrtf(warp_sum(q) * invD + 1e-5f);
  }
  // ---- v = LN1(t) + residual ; out1 ; stash for LN2 ----
  float s2 = 0.f;
  if (two_stage || p.out1) {
    for (int c = lane; c < nchunk; c += 32) {
      float v[8];
      smem_load8(xr, p.x_dtype, c * 8, v);
      if (p.x_act == SGF_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
      }
      if (p.pre_add) {
        float a[8];
        load8(p.pre_add, SGF_F32, c * 8, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += a[j];
      }
      if (p.g1) {
        float g[8], b[8];
        load8(p.g1, SGF_F32, c * 8, g);
        load8(p.b1, SGF_F32, c * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (v[j] - mean1) * rstd1 * g[j] + b[j];
      }
      if (drop.on) {
        float m[8];
        drop_mult8(drop, dst_row, c, m);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= m[j];
      }
      if (p.residual) {
        float r[8];
        smem_load8(rr, p.r_dtype, c * 8, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += r[j];
      }
      if (zero) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      if (p.out1) {
        store8(p.out1, p.out1_dtype, dst_row * p.ld1 + c * 8, v);
        if (p.out1_dtype == SGF_BF16) {  // second LN sees exactly what was stored
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
        }
      }
      if (p.out2) {
        *reinterpret_cast<float4*>(sr + c * 8) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(sr + c * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s2 += v[j];
      }
    }
  }
  if (!p.out2) continue;
  // ---- second LayerNorm (over the stash, or directly over x for the plain x -> LN -> out2 form) ----
  const uint8_t* vrow = two_stage ? reinterpret_cast<const uint8_t*>(sr) : xr;
  const int vdt = two_stage ? SGF_F32 : p.x_dtype;
  if (!two_stage) {
    for (int c = lane; c < nchunk; c += 32) {
      float v[8];
      smem_load8(vrow, vdt, c * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s2 += v[j];
    }
  }
  const float mean2 = warp_sum(s2) * invD;
  float q2 = 0.f;
  for (int c = lane; c < nchunk; c += 32) {
    float v[8];
    smem_load8(vrow, vdt, c * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[j] - mean2;
      q2 += d * d;
    }
  }
  const float rstd2 = rsqrtf(warp_sum(q2) * invD + 1e-5f);
  for (int c = lane; c < nchunk; c += 32) {
    float v[8], g[8], b[8];
    smem_load8(vrow, vdt, c * 8, v);
    load8(p.g2, SGF_F32, c * 8, g);
    load8(p.b2, SGF_F32, c * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (v[j] - mean2) * rstd2 * g[j] + b[j];
    store8(p.out2, SGF_BF16, dst_row * p.ld2 + c * 8, v);
  }
  }  // persistent loop over row groups
}

// ----------------------------------------------------------------------------------------
// Register-resident variant for model-width rows (D <= 1280, no activation): one warp per row, lane l owns the
// 8-column chunks l, l+32, ...; x and the residual row arrive through direct 128-bit global loads issued back to
// back, both LayerNorms run on registers (packed fp32x2 arithmetic), nothing is staged in shared memory, so the
// SM holds 24-32 warps of independent rows.
// ----------------------------------------------------------------------------------------
SGF_DEVICE void load8p(const void* base, int dtype, int64_t off, float2 (&v)[4]) {
  if (dtype == SGF_F32) {
    const float4* q = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
    const float4 a = q[0], b = q[1];
    v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w); v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
    v[0] = unpack_bf16x2(u.x); v[1] = unpack_bf16x2(u.y); v[2] = unpack_bf16x2(u.z); v[3] = unpack_bf16x2(u.w);
  }
}
SGF_DEVICE void store8p(void* base, int dtype, int64_t off, const float2 (&v)[4]) {
  if (dtype == SGF_F32) {
    float4* q = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
    q[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    q[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
  } else {
    uint4 u;
    u.x = pack_bf16x2(v[0].x, v[0].y); u.y = pack_bf16x2(v[1].x, v[1].y);
    u.z = pack_bf16x2(v[2].x, v[2].y); u.w = pack_bf16x2(v[3].x, v[3].y);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = u;
  }
}

template <int NC>
__global__ void __launch_bounds__(128, NC <= 3 ? 5 : 3) row_layernorm_reg_kernel(const RowLnParams p) {
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float invD = 1.0f / static_cast<float>(p.D);
  const int nchunk = p.D >> 3;
  const DropCtx drop = make_drop_ctx(p.drop_p, p.droppath_p, p.drop_seed, p.drop_site, p.drop_step, p.rows_per_sample);
  const float2 zero2 = splat2(0.f);
  pdl_wait();
  for (int row = blockIdx.x * 4 + warp; row < p.rows; row += gridDim.x * 4) {
    const int64_t dst_row =
        p.seg_len > 0 ? static_cast<int64_t>(row / p.seg_len) * p.seg_stride + p.seg_off + row % p.seg_len : row;
    const int64_t src_row = p.gather_idx ? p.gather_idx[row] : row;
    if (p.clear_rowstats && lane < 2) p.clear_rowstats[static_cast<int64_t>(row) * 2 + lane] = 0.f;
    const bool zero = p.zero_row && p.zero_row[row];
    float2 y[NC][4], r[NC][4];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int c = lane + 32 * k;
      if (c < nchunk) {
        load8p(p.x, p.x_dtype, src_row * p.ldx + c * 8, y[k]);
        if (p.residual) load8p(p.residual, p.r_dtype, dst_row * p.ldr + c * 8, r[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int c = lane + 32 * k;
      const bool ok = c < nchunk;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!ok) y[k][j] = zero2;
        if (!ok || !p.residual) r[k][j] = zero2;
      }
      if (ok && p.pre_add) {
        float2 a[4];
        load8p(p.pre_add, SGF_F32, c * 8, a);
#pragma unroll
        for (int j = 0; j < 4; ++j) y[k][j] = add2(y[k][j], a[j]);
      }
    }
    // ---- first LayerNorm ----
    if (p.g1) {
      float2 s2 = zero2;
#pragma unroll
      for (int k = 0; k < NC; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) s2 = add2(s2, y[k][j]);
      const float mean = warp_sum(s2.x + s2.y) * invD;
      const float2 nm = splat2(-mean);
      float2 q = zero2;
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        if (lane + 32 * k < nchunk) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            y[k][j] = add2(y[k][j], nm);
            q = fma2(y[k][j], y[k][j], q);
          }
        }
      }
      const float2 rs = splat2(rsqrtf(warp_sum(q.x + q.y) * invD + 1e-5f));
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const int c = lane + 32 * k;
        if (c < nchunk) {
          float2 g[4], b[4];
          load8p(p.g1, SGF_F32, c * 8, g);
          load8p(p.b1, SGF_F32, c * 8, b);
#pragma unroll
          for (int j = 0; j < 4; ++j) y[k][j] = fma2(mul2(y[k][j], rs), g[j], b[j]);
        }
      }
    }
    // ---- dropout / DropPath, residual, padding rows, first output ----
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int c = lane + 32 * k;
      if (c < nchunk) {
        if (drop.on) {
          float m[8];
          drop_mult8(drop, dst_row, c, m);
#pragma unroll
          for (int j = 0; j < 4; ++j) y[k][j] = mul2(y[k][j], make_float2(m[2 * j], m[2 * j + 1]));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          y[k][j] = add2(y[k][j], r[k][j]);
          if (zero) y[k][j] = zero2;
        }
        if (p.out1) {
          store8p(p.out1, p.out1_dtype, dst_row * p.ld1 + c * 8, y[k]);
          if (p.out1_dtype == SGF_BF16) {  // second LN sees exactly what was stored
#pragma unroll
            for (int j = 0; j < 4; ++j) y[k][j] = unpack_bf16x2(pack_bf16x2(y[k][j].x, y[k][j].y));
          }
        }
      }
    }
    if (!p.out2) continue;
    // ---- second LayerNorm ----
    float2 s2 = zero2;
#pragma unroll
    for (int k = 0; k < NC; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) s2 = add2(s2, y[k][j]);
    const float mean2 = warp_sum(s2.x + s2.y) * invD;
    const float2 nm2 = splat2(-mean2);
    float2 q2 = zero2;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      if (lane + 32 * k < nchunk) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          y[k][j] = add2(y[k][j], nm2);
          q2 = fma2(y[k][j], y[k][j], q2);
        }
      }
    }
    const float2 rs2 = splat2(rsqrtf(warp_sum(q2.x + q2.y) * invD + 1e-5f));
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int c = lane + 32 * k;
      if (c < nchunk) {
        float2 g[4], b[4];
        load8p(p.g2, SGF_F32, c * 8, g);
        load8p(p.b2, SGF_F32, c * 8, b);
#pragma unroll
        for (int j = 0; j < 4; ++j) y[k][j] = fma2(mul2(y[k][j], rs2), g[j], b[j]);
        store8p(p.out2, SGF_BF16, dst_row * p.ld2 + c * 8, y[k]);
      }
    }
  }
}

// ----------------------------------------------------------------------------------------
// attention bias: out = abs (+ dense_add) + rel-pos table lookups on up to two square blocks
// ----------------------------------------------------------------------------------------
struct BiasParams {
  float* out; const float* abs; int64_t head_stride, row_stride; const float* dense_add;
  int H, Tq, Tk, num_blocks;
  sgf_relblock blk[2];
  __half* out16;
};
__global__ void __launch_bounds__(256) build_attn_bias_kernel(const BiasParams p) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= p.Tk) return;
  const float* t = nullptr;
  for (int b = 0; b < p.num_blocks; ++b) {
    const sgf_relblock& k = p.blk[b];
    if (i >= k.lo && i < k.hi && j >= k.lo && j < k.hi)
      t = k.table + k.bucket[k.ids[i - k.lo] * k.bucket_ld + k.ids[j - k.lo]] * p.H;
  }
  const int64_t off = static_cast<int64_t>(i) * p.row_stride + j;
  for (int h = 0; h < p.H; ++h) {
    float v = p.abs[h * p.head_stride + off];
    if (p.dense_add) v += p.dense_add[h * p.head_stride + off];
    if (t) v += t[h];
    if (p.out16) {  // the attention kernels read fp16; an fp32 copy, when asked for, holds exactly the same values
      const __half hv = __float2half_rn(v);
      p.out16[h * p.head_stride + off] = hv;
      v = __half2float(hv);
    }
    if (p.out) p.out[h * p.head_stride + off] = v;
  }
}

// adjoint of build_attn_bias_kernel.  Shared-memory fp32 atomics are CAS loops on this architecture, so the
// relative-position table gradient is a GATHER: the (i,j) positions of every bucket are a static CSR list (built
// once per shape on the host); one warp sums the positions of one (bucket, head) and adds the result to the table
// gradient (unique owner: no atomics, deterministic).  A second elementwise kernel folds the layer's gradient into
// the abs-term accumulator and clears the per-layer buffer.
__global__ void __launch_bounds__(256) bias_table_gather_kernel(const float* __restrict__ dbias, int64_t head_stride,
                                                                const int32_t* __restrict__ order,
                                                                const int32_t* __restrict__ offsets, int num_rel, int H,
                                                                float* __restrict__ dtable) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int h = blockIdx.y;
  if (r >= num_rel) return;
  const int lane = threadIdx.x & 31;
  const int lo = offsets[r], hi = offsets[r + 1];
  if (lo == hi) return;
  const float* src = dbias + h * head_stride;
  float s = 0.f;
  for (int t = lo + lane; t < hi; t += 32) s += __ldg(src + order[t]);
  s = warp_sum(s);
  if (lane == 0) dtable[static_cast<int64_t>(r) * H + h] += s;
}

__global__ void __launch_bounds__(256) bias_fold_clear_kernel(float4* __restrict__ dbias, float4* __restrict__ dabs,
                                                              int64_t n4) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n4) return;
  const float4 v = dbias[i];
  if (dabs) {
    float4 a = dabs[i];
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    dabs[i] = a;
  }
  dbias[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ----------------------------------------------------------------------------------------
// stem helpers
// ----------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int c, int h,
                                    int w) {
  const int64_t total = static_cast<int64_t>(n) * h * w;
  const int64_t pix = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (pix >= total) return;
  const int64_t img = pix / (static_cast<int64_t>(h) * w);
  const int64_t hw = pix - img * h * w;
  for (int cc = 0; cc < c; ++cc)
    y[pix * c + cc] = __float2bfloat16_rn(x[(img * c + cc) * static_cast<int64_t>(h) * w + hw]);
}

// [N,C<=8,H,W] fp32 -> zero-padded [N, H+2*pad, Wp, 8] bf16 (channels C..7 and the border are zero): the input layout of
// the im2col-free 7x7/2 stem convolution (one 16-byte store per padded pixel)
__global__ void nchw_to_nhwc8_padded_kernel(const float* __restrict__ x, uint4* __restrict__ y, int n, int c, int h, int w,
                                            int pad, int hp, int wp) {
  const int64_t total = static_cast<int64_t>(n) * hp * wp;
  const int64_t pix = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (pix >= total) return;
  const int xx = static_cast<int>(pix % wp) - pad;
  const int yy = static_cast<int>((pix / wp) % hp) - pad;
  const int64_t img = pix / (static_cast<int64_t>(wp) * hp);
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (xx >= 0 && xx < w && yy >= 0 && yy < h) {
    for (int cc = 0; cc < c; ++cc) v[cc] = x[((img * c + cc) * h + yy) * static_cast<int64_t>(w) + xx];
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  y[pix] = o;
}

__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int n, int h,
                                    int w, int c, int ho, int wo) {
  const int cg = c / 8;
  const int64_t total = static_cast<int64_t>(n) * ho * wo * cg;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int g = idx % cg;
  int64_t t = idx / cg;
  const int ox = t % wo;
  t /= wo;
  const int oy = t % ho;
  const int img = t / ho;
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 - 1 + ky;
    if (iy < 0 || iy >= h) continue;
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 - 1 + kx;
      if (ix < 0 || ix >= w) continue;
      float v[8];
      load8(x, SGF_BF16, ((static_cast<int64_t>(img) * h + iy) * w + ix) * c + g * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
    }
  }
  store8(y, SGF_BF16, ((static_cast<int64_t>(img) * ho + oy) * wo + ox) * c + g * 8, m);
}

// ragged EmbeddingBag(mean): one warp per bag
__global__ void __launch_bounds__(256) embedding_bag_mean_kernel(const int64_t* __restrict__ tokens, int64_t ld_tokens,
                                                                 const int64_t* __restrict__ ends, int B, int P,
                                                                 const void* table, int table_dtype, int64_t ld_table,
                                                                 int D, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int bag = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (bag >= B * P) return;
  const int b = bag / P, i = bag - b * P;
  const int64_t lo = i == 0 ? 0 : ends[bag - 1];
  const int64_t hi = ends[bag];
  const int64_t* row = tokens + static_cast<int64_t>(b) * ld_tokens;
  const float inv = hi > lo ? 1.0f / static_cast<float>(hi - lo) : 0.f;
  for (int e = lane * 8; e < D; e += 256) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t t = lo; t < hi; ++t) {
      float v[8];
      load8(table, table_dtype, row[t] * ld_table + e, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    store8(out, SGF_F32, static_cast<int64_t>(bag) * D + e, acc);
  }
}

}  // namespace sgf

using namespace sgf;

extern "C" int sgf_embedding_bag_mean(const int64_t* tokens, int64_t ld_tokens, const int64_t* ends, int32_t B,
                                      int32_t P, const void* table, int32_t table_dtype, int64_t ld_table, int32_t D,
                                      float* out, void* stream) {
  SGF_REQUIRE(tokens && ends && table && out && B > 0 && P > 0 && D > 0 && D % 8 == 0 && ld_table % 8 == 0,
              "embedding_bag_mean: bad arguments");
  const int nb = B * P;
  embedding_bag_mean_kernel<<<(nb + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      tokens, ld_tokens, ends, B, P, table, table_dtype, ld_table, D, out);
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}

namespace sgf {
int launch_gelu_ln_fwd_wide(const void* h, int64_t ldh, const float* g, const float* b, void* z, int64_t ldz, int rows,
                            int F, cudaStream_t st);  // train.cu
int launch_ln_fwd_wide_plain(const void* h, int64_t ldh, const float* g, const float* b, void* z, int64_t ldz, int rows,
                             int F, cudaStream_t st);  // train.cu
}

extern "C" int sgf_row_layernorm(const sgf_rowln_args* a, void* stream) {
  SGF_REQUIRE(a != nullptr, "row_layernorm: null args");
  if (a->x_act == SGF_ACT_GELU && a->x_dtype == SGF_BF16 && !a->gather_idx && !a->pre_add && !a->g1 && !a->residual &&
      !a->out1 && a->out2 && a->g2 && a->b2 && !a->zero_row && a->seg_len == 0 && !a->clear_rowstats && a->D >= 1024 &&
      a->drop_p == 0.f && a->droppath_p == 0.f &&
      a->D <= 4096 && a->D % 8 == 0 && a->rows > 0 && a->ldx % 8 == 0 && a->ld2 % 8 == 0 &&
      reinterpret_cast<uintptr_t>(a->x) % 16 == 0 && reinterpret_cast<uintptr_t>(a->out2) % 16 == 0)
    return launch_gelu_ln_fwd_wide(a->x, a->ldx, a->g2, a->b2, a->out2, a->ld2, a->rows, a->D,
                                   reinterpret_cast<cudaStream_t>(stream));
  // plain wide LayerNorm (the inference FFN: GELU already applied by the GEMM epilogue)
  if (a->x_act == SGF_ACT_NONE && a->x_dtype == SGF_BF16 && !a->gather_idx && !a->pre_add && !a->g1 && !a->residual &&
      !a->out1 && a->out2 && a->g2 && a->b2 && !a->zero_row && a->seg_len == 0 && !a->clear_rowstats && a->D >= 1536 &&
      a->drop_p == 0.f && a->droppath_p == 0.f && a->D <= 8192 && a->D % 512 == 0 && a->rows > 0 && a->ldx % 8 == 0 &&
      a->ld2 % 8 == 0 && reinterpret_cast<uintptr_t>(a->x) % 16 == 0 && reinterpret_cast<uintptr_t>(a->out2) % 16 == 0)
    return launch_ln_fwd_wide_plain(a->x, a->ldx, a->g2, a->b2, a->out2, a->ld2, a->rows, a->D,
                                    reinterpret_cast<cudaStream_t>(stream));
  SGF_REQUIRE(a->rows > 0 && a->D > 0 && a->D % 8 == 0, "row_layernorm: D must be a positive multiple of 8 (D=%d)", a->D);
  SGF_REQUIRE(a->D <= 5120, "row_layernorm: D=%d exceeds the 5120 register-resident limit", a->D);
  SGF_REQUIRE((a->g1 == nullptr) == (a->b1 == nullptr), "row_layernorm: g1/b1 must both be set or both null");
  SGF_REQUIRE(!a->out2 || (a->g2 && a->b2), "row_layernorm: out2 needs g2/b2");
  SGF_REQUIRE(a->out1 || a->out2, "row_layernorm: no output");
  SGF_REQUIRE(a->ldx % 8 == 0 && (!a->out1 || a->ld1 % 8 == 0) && (!a->out2 || a->ld2 % 8 == 0) &&
                  (!a->residual || a->ldr % 8 == 0),
              "row_layernorm: row strides must be multiples of 8 elements");
  RowLnParams p{a->x, a->ldx, a->x_dtype, a->gather_idx, a->pre_add, a->g1, a->b1, a->residual, a->ldr, a->r_dtype,
                a->out1, a->ld1, a->out1_dtype, a->g2, a->b2, a->out2, a->ld2, a->zero_row, a->rows, a->D,
                a->seg_len, a->seg_stride, a->seg_off, a->clear_rowstats, a->x_act, a->drop_p, a->droppath_p,
                a->drop_seed, a->drop_site, a->rows_per_sample, a->drop_step};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int xs = a->x_dtype == SGF_F32 ? 4 : 2, rs = a->r_dtype == SGF_F32 ? 4 : 2;
  SGF_REQUIRE(reinterpret_cast<uintptr_t>(a->x) % 16 == 0 && (a->ldx * xs) % 16 == 0 &&
                  (!a->residual || (reinterpret_cast<uintptr_t>(a->residual) % 16 == 0 && (a->ldr * rs) % 16 == 0)),
              "row_layernorm: x/residual rows must be 16-byte aligned");
  const bool out_ok = (!a->out1 || reinterpret_cast<uintptr_t>(a->out1) % 16 == 0) &&
                      (!a->out2 || reinterpret_cast<uintptr_t>(a->out2) % 16 == 0);
  if (a->x_act == SGF_ACT_NONE && a->D <= 1280 && out_ok) {  // model-width rows: register-resident kernel
    const int nc = (a->D + 255) / 256;
    const int ngrp = (a->rows + 3) / 4;
    const int resident = 148 * (nc <= 3 ? 5 : 3);
    static const int grid_policy = getenv("SGF_LN_GRID") ? atoi(getenv("SGF_LN_GRID")) : 0;  // A/B: 1 = one CTA per 4 rows
    const dim3 grid(grid_policy == 1 ? ngrp : (ngrp < resident ? ngrp : resident)), block(128);
    switch (nc) {
      case 1: SGF_CHECK_CUDA(launch_pdl(row_layernorm_reg_kernel<1>, grid, block, size_t(0), st, p)); break;
      case 2: SGF_CHECK_CUDA(launch_pdl(row_layernorm_reg_kernel<2>, grid, block, size_t(0), st, p)); break;
      case 3: SGF_CHECK_CUDA(launch_pdl(row_layernorm_reg_kernel<3>, grid, block, size_t(0), st, p)); break;
      case 4: SGF_CHECK_CUDA(launch_pdl(row_layernorm_reg_kernel<4>, grid, block, size_t(0), st, p)); break;
      default: SGF_CHECK_CUDA(launch_pdl(row_layernorm_reg_kernel<5>, grid, block, size_t(0), st, p)); break;
    }
    count_launch();
    return SGF_OK;
  }
  const int x_row_bytes = a->D * xs;
  const int r_row_bytes = a->residual ? a->D * rs : 0;
  const bool two_stage = a->out2 && (a->g1 || a->residual || a->pre_add || a->out1 || a->x_act || a->drop_p > 0.f ||
                                     a->droppath_p > 0.f);
  const int s_row_bytes = two_stage ? a->D * 4 : 0;
  // per row: two staging buffers (double-buffered bulk copies) + the fp32 stash; three CTAs per SM when the rows allow
  const int per_row = 2 * (x_row_bytes + r_row_bytes) + s_row_bytes + 16;
  int kLnRows = kLnMaxRows;
  while (kLnRows > 1 && kLnRows * per_row > 72 * 1024) kLnRows >>= 1;
  const int smem = kLnRows * per_row;
  SGF_REQUIRE(smem <= 200 * 1024, "row_layernorm: D=%d too wide for this operand combination (%d B smem)", a->D, smem);
  SGF_REQUIRE(reinterpret_cast<uintptr_t>(a->x) % 16 == 0 && (a->ldx * xs) % 16 == 0 &&
                  (!a->residual || (reinterpret_cast<uintptr_t>(a->residual) % 16 == 0 && (a->ldr * rs) % 16 == 0)),
              "row_layernorm: x/residual rows must be 16-byte aligned");
  static int configured_smem = 0;
  if (smem > configured_smem) {
    SGF_CHECK_CUDA(cudaFuncSetAttribute(row_layernorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured_smem = 200 * 1024;
  }
  const int ngroups = (a->rows + kLnRows - 1) / kLnRows;
  int per_sm = (220 * 1024) / (smem + 1024);
  per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);
  const int resident = 148 * per_sm;
  dim3 block(kLnRows * 32), grid(ngroups < resident ? ngroups : resident);
  SGF_CHECK_CUDA(launch_pdl(row_layernorm_kernel, grid, block, smem, st, p, x_row_bytes, r_row_bytes, s_row_bytes));
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}

extern "C" int sgf_build_attn_bias(const sgf_bias_args* a, void* stream) {
  SGF_REQUIRE(a != nullptr && (a->out || a->out_f16) && a->abs, "build_attn_bias: null pointer");
  SGF_REQUIRE(a->H > 0 && a->Tq > 0 && a->Tk > 0 && a->row_stride >= a->Tk, "build_attn_bias: bad shape");
  SGF_REQUIRE(a->num_blocks >= 0 && a->num_blocks <= 2, "build_attn_bias: at most 2 blocks");
  BiasParams p{a->out, a->abs, a->head_stride, a->row_stride, a->dense_add, a->H, a->Tq, a->Tk, a->num_blocks, {},
               reinterpret_cast<__half*>(a->out_f16)};
  for (int b = 0; b < a->num_blocks; ++b) {
    const sgf_relblock& k = a->blocks[b];
    SGF_REQUIRE(k.bucket && k.ids && k.table && k.lo >= 0 && k.hi > k.lo && k.hi <= a->Tq && k.hi <= a->Tk,
                "build_attn_bias: bad block %d [%d,%d)", b, k.lo, k.hi);
    p.blk[b] = k;
  }
  dim3 grid((a->Tk + 255) / 256, a->Tq), block(256);
  build_attn_bias_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}

extern "C" int sgf_attn_bias_bwd(const sgf_bias_bwd_args* a, void* stream) {
  SGF_REQUIRE(a != nullptr && a->dbias, "attn_bias_bwd: null pointer");
  SGF_REQUIRE(a->H > 0 && a->Tq > 0 && a->row_stride % 4 == 0 && a->head_stride == static_cast<int64_t>(a->Tq) * a->row_stride &&
                  reinterpret_cast<uintptr_t>(a->dbias) % 16 == 0,
              "attn_bias_bwd: dbias must be a dense [H,Tq,row_stride] buffer with row_stride %% 4 == 0");
  SGF_REQUIRE(a->num_blocks >= 0 && a->num_blocks <= 2, "attn_bias_bwd: at most 2 blocks");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int b = 0; b < a->num_blocks; ++b) {
    SGF_REQUIRE(a->order[b] && a->offsets[b] && a->dtable[b] && a->num_rel[b] > 0, "attn_bias_bwd: bad block %d", b);
    dim3 grid((a->num_rel[b] + 7) / 8, a->H);
    bias_table_gather_kernel<<<grid, 256, 0, st>>>(a->dbias, a->head_stride, a->order[b], a->offsets[b], a->num_rel[b],
                                                   a->H, a->dtable[b]);
    SGF_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  const int64_t n4 = static_cast<int64_t>(a->H) * a->head_stride / 4;
  bias_fold_clear_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(
      reinterpret_cast<float4*>(a->dbias), reinterpret_cast<float4*>(a->dabs_acc), n4);
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}

extern "C" int sgf_nchw_f32_to_nhwc_bf16(const float* x, void* y, int32_t n, int32_t c, int32_t h, int32_t w,
                                         void* stream) {
  SGF_REQUIRE(x && y && n > 0 && c > 0 && h > 0 && w > 0, "nchw_to_nhwc: bad args");
  const int64_t total = static_cast<int64_t>(n) * h * w;
  nchw_to_nhwc_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<__nv_bfloat16*>(y), n, c, h, w);
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}

extern "C" int sgf_nchw_f32_to_nhwc8_padded(const float* x, void* y, int32_t n, int32_t c, int32_t h, int32_t w, int32_t pad,
                                            int32_t hp, int32_t wp, void* stream) {
  SGF_REQUIRE(x && y && n > 0 && c > 0 && c <= 8 && h > 0 && w > 0 && pad >= 0 && hp >= h + pad && wp >= w + pad,
              "nchw_to_nhwc8_padded: bad args");
  SGF_REQUIRE(reinterpret_cast<uintptr_t>(y) % 16 == 0, "nchw_to_nhwc8_padded: y must be 16-byte aligned");
  const int64_t total = static_cast<int64_t>(n) * hp * wp;
  nchw_to_nhwc8_padded_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<uint4*>(y), n, c, h, w, pad, hp, wp);
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}

extern "C" int sgf_maxpool3x3s2_nhwc(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t ho,
                                     int32_t wo, void* stream) {
  SGF_REQUIRE(x && y && c % 8 == 0, "maxpool: C must be a multiple of 8");
  const int64_t total = static_cast<int64_t>(n) * ho * wo * (c / 8);
  maxpool3x3s2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), n, h, w, c, ho, wo);
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}
