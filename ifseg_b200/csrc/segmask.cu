// Patch-grid -> mask: bilinear x(h/hp) upsample fused with the class argmax (and optional
// per-class area histograms).  The full-resolution [B,h,w,C] logits that the reference
// materialises (seg_criterion.py:237-244; 138 MB per 480x480 image at C=150) never exist:
// each thread evaluates the C interpolated logits of its pixel in registers.  HBM traffic is
// the low-res logits (L2-resident) in and one int64 per pixel out.
//
// Arithmetic mirrors ATen's upsample_bilinear2d(align_corners=False): source index
// s = max(0, scale*(d+0.5)-0.5), i0 = (int)s, i1 = i0 + (i0 < in-1), l1 = s - i0, l0 = 1 - l1,
// value = h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11), evaluated without FMA contraction.
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"

namespace sgf {

struct SegmaskParams {
  const float* logits;
  int64_t batch_stride, tok_stride;
  int B, C, hp, wp, h, w;
  int64_t* mask;
  const int64_t* target;
  float* area_intersect;
  float* area_pred;
  float* area_label;
  float scale_h, scale_w;
};

// Rounding sequence of the interpolation (the argmax of near-tied classes depends on the last bit):
//   kArithPlain     every product and sum rounded separately (ATen's CPU contiguous kernel)
//   kArithFma + v   nvcc-contracted forms of ATen's CUDA expression (aten/native/cuda/UpSampleBilinear2d.cu)
//                     val = h0 * (w0*v00 + w1*v01) + h1 * (w0*v10 + w1*v11),   src = fma(scale, dst + 0.5, -0.5)
//                   each of the three sums becomes one multiply + one fma; which product stays a multiply is the
//                   compiler's choice and differs between ATen's two CUDA kernels, so all 8 forms exist:
//                     bit 0: top  = fma(w1, v01, w0*v00)   instead of fma(w0, v00, w1*v01)
//                     bit 1: bot  = fma(w1, v11, w0*v10)   instead of fma(w0, v10, w1*v11)
//                     bit 2: val  = fma(h1, bot, h0*top)   instead of fma(h0, top, h1*bot)
enum : int { kArithPlain = 0, kArithFma = 8 };

template <int kArith>
SGF_DEVICE float lerp4(float hl0, float hl1, float wl0, float wl1, float v00, float v01, float v10, float v11) {
  if constexpr (kArith == kArithPlain) {
    const float top = __fadd_rn(__fmul_rn(wl0, v00), __fmul_rn(wl1, v01));
    const float bot = __fadd_rn(__fmul_rn(wl0, v10), __fmul_rn(wl1, v11));
    return __fadd_rn(__fmul_rn(hl0, top), __fmul_rn(hl1, bot));
  } else {
    constexpr int v = kArith - kArithFma;
    const float top = (v & 1) ? __fmaf_rn(wl1, v01, __fmul_rn(wl0, v00)) : __fmaf_rn(wl0, v00, __fmul_rn(wl1, v01));
    const float bot = (v & 2) ? __fmaf_rn(wl1, v11, __fmul_rn(wl0, v10)) : __fmaf_rn(wl0, v10, __fmul_rn(wl1, v11));
    return (v & 4) ? __fmaf_rn(hl1, bot, __fmul_rn(hl0, top)) : __fmaf_rn(hl0, top, __fmul_rn(hl1, bot));
  }
}

template <int kArith = kArithPlain>
SGF_DEVICE void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = kArith == kArithPlain ? __fsub_rn(__fmul_rn(scale, static_cast<float>(dst) + 0.5f), 0.5f)
                                  : __fmaf_rn(scale, static_cast<float>(dst) + 0.5f, -0.5f);
  s = s < 0.f ? 0.f : s;
  i0 = static_cast<int>(s);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = __fsub_rn(s, static_cast<float>(i0));
  l0 = __fsub_rn(1.0f, l1);
}

template <int kArith>
__global__ void __launch_bounds__(256) upsample_argmax_kernel(const SegmaskParams p) {
  extern __shared__ float hist[];  // [3][C] when histograms are requested
  const bool want_hist = p.area_pred != nullptr;
  if (want_hist) {
    for (int i = threadIdx.x; i < 3 * p.C; i += blockDim.x) hist[i] = 0.f;
    __syncthreads();
  }
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (x < p.w) {
    int y0, y1, x0, x1;
    float hl0, hl1, wl0, wl1;
    src_index<kArith>(p.scale_h, y, p.hp, y0, y1, hl0, hl1);
    src_index<kArith>(p.scale_w, x, p.wp, x0, x1, wl0, wl1);
    const float* base = p.logits + static_cast<int64_t>(b) * p.batch_stride;
    const float* p00 = base + static_cast<int64_t>(y0 * p.wp + x0) * p.tok_stride;
    const float* p01 = base + static_cast<int64_t>(y0 * p.wp + x1) * p.tok_stride;
    const float* p10 = base + static_cast<int64_t>(y1 * p.wp + x0) * p.tok_stride;
    const float* p11 = base + static_cast<int64_t>(y1 * p.wp + x1) * p.tok_stride;
    float best = -INFINITY;
    int best_c = 0;
    for (int c = 0; c < p.C; ++c) {
      const float v = lerp4<kArith>(hl0, hl1, wl0, wl1, __ldg(p00 + c), __ldg(p01 + c), __ldg(p10 + c), __ldg(p11 + c));
      if (v > best || (c == 0)) {  // first maximum wins, NaN-free inputs assumed
        if (c == 0 || v > best) {
          best = v;
          best_c = c;
        }
      }
    }
    const int64_t pix = (static_cast<int64_t>(b) * p.h + y) * p.w + x;
    p.mask[pix] = best_c;
    if (want_hist) {
      bool counted = true;
      if (p.target) {
        const int64_t t = p.target[pix];
        counted = t >= 0 && t < p.C;  // ignored pixels (pad / "unknown") contribute to nothing
        if (counted) {
          atomicAdd(&hist[2 * p.C + static_cast<int>(t)], 1.f);
          if (t == best_c) atomicAdd(&hist[best_c], 1.f);
        }
      }
      if (counted) atomicAdd(&hist[p.C + best_c], 1.f);
    }
  }
  if (want_hist) {
    __syncthreads();
    for (int i = threadIdx.x; i < p.C; i += blockDim.x) {
      if (p.area_intersect && hist[i] != 0.f) atomicAdd(&p.area_intersect[i], hist[i]);
      if (hist[p.C + i] != 0.f) atomicAdd(&p.area_pred[i], hist[p.C + i]);
      if (p.area_label && hist[2 * p.C + i] != 0.f) atomicAdd(&p.area_label[i], hist[2 * p.C + i]);
    }
  }
}

// Four horizontally adjacent pixels per thread.  The scalar kernel above is LSU-bound for wide class counts (4 loads per
// class and pixel: 325 us at C = 150, 8 x 480^2).  At an upsampling factor >= 4 the four pixels of a thread see at most
// two adjacent tap-column pairs, (xa, xa+1) or (xa+1, xa+2): six loads per class serve four pixels, and every pixel still
// runs exactly the per-pixel rounding sequence of the scalar kernel (lerp4<kArith>), so the masks are bit-identical.
template <int kArith>
__global__ void __launch_bounds__(128) upsample_argmax4_kernel(const SegmaskParams p) {
  extern __shared__ float hist[];  // [3][C] when histograms are requested
  const bool want_hist = p.area_pred != nullptr;
  if (want_hist) {
    for (int i = threadIdx.x; i < 3 * p.C; i += blockDim.x) hist[i] = 0.f;
    __syncthreads();
  }
  const int xq = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (xq < p.w) {
    int y0, y1;
    float hl0, hl1;
    src_index<kArith>(p.scale_h, y, p.hp, y0, y1, hl0, hl1);
    int x0[4], x1[4];
    float wl0[4], wl1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) src_index<kArith>(p.scale_w, min(xq + i, p.w - 1), p.wp, x0[i], x1[i], wl0[i], wl1[i]);
    const int xa = x0[0], xb = min(xa + 1, p.wp - 1), xc = min(xa + 2, p.wp - 1);
    bool sel[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) sel[i] = x0[i] != xa;  // then x0 == xa + 1 (== xb) and x1 == xc
    const float* base = p.logits + static_cast<int64_t>(b) * p.batch_stride;
    const float* qa0 = base + static_cast<int64_t>(y0 * p.wp + xa) * p.tok_stride;
    const float* qb0 = base + static_cast<int64_t>(y0 * p.wp + xb) * p.tok_stride;
    const float* qc0 = base + static_cast<int64_t>(y0 * p.wp + xc) * p.tok_stride;
    const float* qa1 = base + static_cast<int64_t>(y1 * p.wp + xa) * p.tok_stride;
    const float* qb1 = base + static_cast<int64_t>(y1 * p.wp + xb) * p.tok_stride;
    const float* qc1 = base + static_cast<int64_t>(y1 * p.wp + xc) * p.tok_stride;
    float best[4];
    int best_c[4] = {0, 0, 0, 0};
#pragma unroll 2
    for (int c = 0; c < p.C; ++c) {
      const float a0 = __ldg(qa0 + c), b0 = __ldg(qb0 + c), c0 = __ldg(qc0 + c);
      const float a1 = __ldg(qa1 + c), b1 = __ldg(qb1 + c), c1 = __ldg(qc1 + c);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v = lerp4<kArith>(hl0, hl1, wl0[i], wl1[i], sel[i] ? b0 : a0, sel[i] ? c0 : b0, sel[i] ? b1 : a1,
                                      sel[i] ? c1 : b1);
        if (c == 0 || v > best[i]) {  // first maximum wins, NaN-free inputs assumed
          best[i] = v;
          best_c[i] = c;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (xq + i >= p.w) break;
      const int64_t pix = (static_cast<int64_t>(b) * p.h + y) * p.w + xq + i;
      p.mask[pix] = best_c[i];
      if (want_hist) {
        bool counted = true;
        if (p.target) {
          const int64_t t = p.target[pix];
          counted = t >= 0 && t < p.C;  // ignored pixels (pad / "unknown") contribute to nothing
          if (counted) {
            atomicAdd(&hist[2 * p.C + static_cast<int>(t)], 1.f);
            if (t == best_c[i]) atomicAdd(&hist[best_c[i]], 1.f);
          }
        }
        if (counted) atomicAdd(&hist[p.C + best_c[i]], 1.f);
      }
    }
  }
  if (want_hist) {
    __syncthreads();
    for (int i = threadIdx.x; i < p.C; i += blockDim.x) {
      if (p.area_intersect && hist[i] != 0.f) atomicAdd(&p.area_intersect[i], hist[i]);
      if (hist[p.C + i] != 0.f) atomicAdd(&p.area_pred[i], hist[p.C + i]);
      if (p.area_label && hist[2 * p.C + i] != 0.f) atomicAdd(&p.area_label[i], hist[2 * p.C + i]);
    }
  }
}

struct SeglossParams {
  const float* logits;
  int64_t batch_stride, tok_stride;
  int B, C, hp, wp, h, w;
  const int64_t* target;
  float eps;
  float* out;
  float scale_h, scale_w;
  float* lse_out;
};

__global__ void __launch_bounds__(256) upsample_ce_loss_kernel(const SeglossParams p) {
  __shared__ float red[2][8];
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  float loss = 0.f, cnt = 0.f;
  if (x < p.w) {
    const int64_t pix = (static_cast<int64_t>(b) * p.h + y) * p.w + x;
    const int64_t t = p.target[pix];
    if (t >= 0 && t < p.C) {
      int y0, y1, x0, x1;
      float hl0, hl1, wl0, wl1;
      src_index(p.scale_h, y, p.hp, y0, y1, hl0, hl1);
      src_index(p.scale_w, x, p.wp, x0, x1, wl0, wl1);
      const float* base = p.logits + static_cast<int64_t>(b) * p.batch_stride;
      const float* p00 = base + static_cast<int64_t>(y0 * p.wp + x0) * p.tok_stride;
      const float* p01 = base + static_cast<int64_t>(y0 * p.wp + x1) * p.tok_stride;
      const float* p10 = base + static_cast<int64_t>(y1 * p.wp + x0) * p.tok_stride;
      const float* p11 = base + static_cast<int64_t>(y1 * p.wp + x1) * p.tok_stride;
      // online logsumexp over the C interpolated logits
      float m = -INFINITY, ssum = 0.f, vt = 0.f, vsum = 0.f;
      for (int c = 0; c < p.C; ++c) {
        const float top = __fadd_rn(__fmul_rn(wl0, __ldg(p00 + c)), __fmul_rn(wl1, __ldg(p01 + c)));
        const float bot = __fadd_rn(__fmul_rn(wl0, __ldg(p10 + c)), __fmul_rn(wl1, __ldg(p11 + c)));
        const float v = __fadd_rn(__fmul_rn(hl0, top), __fmul_rn(hl1, bot));
        if (c == t) vt = v;
        vsum += v;
        const float mn = fmaxf(m, v);
        ssum = ssum * __expf(m - mn) + __expf(v - mn);
        m = mn;
      }
      const float lse = m + __logf(ssum);
      if (p.lse_out) p.lse_out[pix] = lse;
      loss = (1.f - p.eps) * (lse - vt) + p.eps * (lse - vsum / static_cast<float>(p.C));
      cnt = 1.f;
    }
  }
  loss = warp_sum(loss);
  cnt = warp_sum(cnt);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = loss;
    red[1][warp] = cnt;
  }
  __syncthreads();
  if (warp == 0) {
    loss = lane < 8 ? red[0][lane] : 0.f;
    cnt = lane < 8 ? red[1][lane] : 0.f;
    loss = warp_sum(loss);
    cnt = warp_sum(cnt);
    if (lane == 0 && cnt > 0.f) {
      atomicAdd(&p.out[0], loss);
      atomicAdd(&p.out[1], cnt);
    }
  }
}


// Four pixels per thread, same tap sharing as upsample_argmax4_kernel; the online logsumexp costs ONE exponential per
// class and pixel (exp(min(m, v) - max(m, v)); the other term of the textbook update is exp(0) = 1 exactly).
__global__ void __launch_bounds__(128) upsample_ce_loss4_kernel(const SeglossParams p) {
  __shared__ float red[2][4];
  const int xq = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  float loss = 0.f, cnt = 0.f;
  if (xq < p.w) {
    int64_t tg[4];
    bool live[4], any = false;
    const int64_t pix0 = (static_cast<int64_t>(b) * p.h + y) * p.w + xq;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      tg[i] = xq + i < p.w ? p.target[pix0 + i] : -1;
      live[i] = tg[i] >= 0 && tg[i] < p.C;
      any |= live[i];
    }
    if (any) {
      int y0, y1;
      float hl0, hl1;
      src_index(p.scale_h, y, p.hp, y0, y1, hl0, hl1);
      int x0[4], x1[4];
      float wl0[4], wl1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) src_index(p.scale_w, min(xq + i, p.w - 1), p.wp, x0[i], x1[i], wl0[i], wl1[i]);
      const int xa = x0[0], xb = min(xa + 1, p.wp - 1), xc = min(xa + 2, p.wp - 1);
      bool sel[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) sel[i] = x0[i] != xa;
      const float* base = p.logits + static_cast<int64_t>(b) * p.batch_stride;
      const float* qa0 = base + static_cast<int64_t>(y0 * p.wp + xa) * p.tok_stride;
      const float* qb0 = base + static_cast<int64_t>(y0 * p.wp + xb) * p.tok_stride;
      const float* qc0 = base + static_cast<int64_t>(y0 * p.wp + xc) * p.tok_stride;
      const float* qa1 = base + static_cast<int64_t>(y1 * p.wp + xa) * p.tok_stride;
      const float* qb1 = base + static_cast<int64_t>(y1 * p.wp + xb) * p.tok_stride;
      const float* qc1 = base + static_cast<int64_t>(y1 * p.wp + xc) * p.tok_stride;
      float m[4], ssum[4], vt[4], vsum[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        m[i] = -INFINITY;
        ssum[i] = vt[i] = vsum[i] = 0.f;
      }
#pragma unroll 2
      for (int c = 0; c < p.C; ++c) {
        const float a0 = __ldg(qa0 + c), b0 = __ldg(qb0 + c), c0 = __ldg(qc0 + c);
        const float a1 = __ldg(qa1 + c), b1 = __ldg(qb1 + c), c1 = __ldg(qc1 + c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float top = __fadd_rn(__fmul_rn(wl0[i], sel[i] ? b0 : a0), __fmul_rn(wl1[i], sel[i] ? c0 : b0));
          const float bot = __fadd_rn(__fmul_rn(wl0[i], sel[i] ? b1 : a1), __fmul_rn(wl1[i], sel[i] ? c1 : b1));
          const float v = __fadd_rn(__fmul_rn(hl0, top), __fmul_rn(hl1, bot));
          if (c == tg[i]) vt[i] = v;
          vsum[i] += v;
          const bool up = v > m[i];
          const float e = __expf(fminf(m[i], v) - fmaxf(m[i], v));
          ssum[i] = up ? fmaf(ssum[i], e, 1.0f) : ssum[i] + e;
          m[i] = fmaxf(m[i], v);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!live[i]) continue;
        const float lse = m[i] + __logf(ssum[i]);
        if (p.lse_out) p.lse_out[pix0 + i] = lse;
        loss += (1.f - p.eps) * (lse - vt[i]) + p.eps * (lse - vsum[i] / static_cast<float>(p.C));
        cnt += 1.f;
      }
    }
  }
  loss = warp_sum(loss);
  cnt = warp_sum(cnt);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = loss;
    red[1][warp] = cnt;
  }
  __syncthreads();
  if (warp == 0) {
    loss = lane < 4 ? red[0][lane] : 0.f;
    cnt = lane < 4 ? red[1][lane] : 0.f;
    loss = warp_sum(loss);
    cnt = warp_sum(cnt);
    if (lane == 0 && cnt > 0.f) {
      atomicAdd(&p.out[0], loss);
      atomicAdd(&p.out[1], cnt);
    }
  }
}

// ----------------------------------------------------------------------------------------
// adjoint of upsample + pixel cross-entropy w.r.t. the low-resolution logits (gather form)
// ----------------------------------------------------------------------------------------
struct SeglossBwdParams {
  const float* logits;
  int64_t batch_stride, tok_stride;
  int B, C, hp, wp, h, w;
  const int64_t* target;
  const float* lse;
  const float* count;
  float eps, grad_scale;
  __nv_bfloat16* dlogits;
  int64_t d_batch_stride, d_tok_stride;
  int d_tokens;
  float scale_h, scale_w;
};

// Stage 1 (whole CTA): everything that does not depend on the class goes to shared memory once -- per footprint row
// the two taps and their weights; the footprint columns re-laid out as RUNS of constant tap pair, each run padded to a
// multiple of four columns (padding carries weight 0 / logsumexp +inf, i.e. contributes exactly 0); per footprint
// pixel the weight w = wy*wx (0 for ignored pixels) and -log2e * logsumexp; and the class-dependent one-hot term
// sum_{pixels with target c} w as a fixed-point histogram (native integer shared-memory atomics).
// Stage 2 (thread = class): per footprint row the two row taps are blended once (3 values); inside a run four
// pixels cost four 128-bit shared loads, six packed-fp32x2 FMAs and four ex2 -- the kernel runs at the MUFU rate.
// Gather form: no global atomics, deterministic.
struct FootTap {
  float l0, l1, w;
  int i0, i1;  // tap indices relative to (p - 1): 0..2
};
static constexpr int kCeMaxRuns = 64;

__global__ void __launch_bounds__(256) upsample_ce_bwd_kernel(const SeglossBwdParams p, const int fy_max, const int fx_max,
                                                              const int nxp_max, const float qscale) {
  extern __shared__ __align__(16) uint8_t ce_smem[];
  float* pw = reinterpret_cast<float*>(ce_smem);                  // [nY][nXp] pixel weight
  float* pl = pw + fy_max * nxp_max;                               // [nY][nXp] -log2e * lse
  float* cl0 = pl + fy_max * nxp_max;                              // [nXp] column tap weights (run-padded layout)
  float* cl1 = cl0 + nxp_max;
  int4* runs = reinterpret_cast<int4*>(cl1 + nxp_max);             // (base, padded length, q0, q1)
  FootTap* rowp = reinterpret_cast<FootTap*>(runs + kCeMaxRuns);
  FootTap* colp = rowp + fy_max;
  int* colpos = reinterpret_cast<int*>(colp + fx_max);             // footprint column -> run-padded position
  int* hitq = colpos + fx_max;                                     // [C] fixed-point one-hot weights; [C] = total
  __shared__ int s_nruns, s_nxp;
  const int tok = blockIdx.x;
  const int b = blockIdx.y;
  __nv_bfloat16* drow = p.dlogits + static_cast<int64_t>(b) * p.d_batch_stride + static_cast<int64_t>(tok) * p.d_tok_stride;
  if (tok >= p.hp * p.wp) {  // eos slot / padding rows receive no gradient (seg_criterion.py:238)
    for (int c = threadIdx.x; c < p.d_tok_stride; c += blockDim.x) drow[c] = __float2bfloat16_rn(0.f);
    return;
  }
  const int py = tok / p.wp, px = tok - py * p.wp;
  // conservative pixel footprint of the patch: src = scale*(dst+0.5)-0.5 in (p-1, p+1)
  const float inv_h = static_cast<float>(p.h) / static_cast<float>(p.hp), inv_w = static_cast<float>(p.w) / static_cast<float>(p.wp);
  int ylo = static_cast<int>(floorf((py - 0.5f) * inv_h - 0.5f)) - 1, yhi = static_cast<int>(ceilf((py + 1.5f) * inv_h - 0.5f)) + 1;
  int xlo = static_cast<int>(floorf((px - 0.5f) * inv_w - 0.5f)) - 1, xhi = static_cast<int>(ceilf((px + 1.5f) * inv_w - 0.5f)) + 1;
  ylo = max(ylo, 0); xlo = max(xlo, 0); yhi = min(yhi, p.h - 1); xhi = min(xhi, p.w - 1);
  if (py == 0) ylo = 0;  // clamped border rows/columns all map onto the first / last patch
  if (px == 0) xlo = 0;
  if (py == p.hp - 1) yhi = p.h - 1;
  if (px == p.wp - 1) xhi = p.w - 1;
  const int nY = min(yhi - ylo + 1, fy_max), nX = min(xhi - xlo + 1, fx_max);
  for (int i = threadIdx.x; i < nY + nX; i += blockDim.x) {
    const bool is_row = i < nY;
    const int d = is_row ? ylo + i : xlo + (i - nY);
    int i0, i1;
    float l0, l1;
    src_index(is_row ? p.scale_h : p.scale_w, d, is_row ? p.hp : p.wp, i0, i1, l0, l1);
    const int want = is_row ? py : px;
    FootTap t;
    t.l0 = l0; t.l1 = l1;
    t.w = (i0 == want ? l0 : 0.f) + (i1 == want ? l1 : 0.f);
    t.i0 = min(max(i0 - (want - 1), 0), 2);
    t.i1 = min(max(i1 - (want - 1), 0), 2);
    (is_row ? rowp[i] : colp[i - nY]) = t;
  }
  for (int i = threadIdx.x; i <= p.C; i += blockDim.x) hitq[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {  // run layout (a handful of runs: the tap pair changes once per source column)
    int nr = 0, pos = 0;
    for (int ix = 0; ix < nX; ++ix) {
      const bool fresh = ix == 0 || colp[ix].i0 != colp[ix - 1].i0 || colp[ix].i1 != colp[ix - 1].i1;
      if (fresh && nr < kCeMaxRuns) {
        if (nr > 0) runs[nr - 1].y = ((pos + 3) & ~3) - runs[nr - 1].x;
        pos = (pos + 3) & ~3;
        runs[nr++] = make_int4(pos, 0, colp[ix].i0, colp[ix].i1);
      }
      colpos[ix] = pos;
      ++pos;
    }
    pos = (pos + 3) & ~3;
    if (nr > 0) runs[nr - 1].y = pos - runs[nr - 1].x;
    s_nruns = nr;
    s_nxp = min(pos, nxp_max);
  }
  __syncthreads();
  const int nXp = s_nxp, nruns = s_nruns;
  for (int i = threadIdx.x; i < nXp; i += blockDim.x) cl0[i] = cl1[i] = 0.f;
  for (int i = threadIdx.x; i < nY * nXp; i += blockDim.x) {
    pw[i] = 0.f;
    pl[i] = -INFINITY;
  }
  __syncthreads();
  for (int ix = threadIdx.x; ix < nX; ix += blockDim.x) {
    const int pos = colpos[ix];
    if (pos < nXp) {
      cl0[pos] = colp[ix].l0;
      cl1[pos] = colp[ix].l1;
    }
  }
  int swq = 0;
  for (int i = threadIdx.x; i < nY * nX; i += blockDim.x) {
    const int iy = i / nX, ix = i - iy * nX;
    const float w = rowp[iy].w * colp[ix].w;
    const int64_t pix = (static_cast<int64_t>(b) * p.h + (ylo + iy)) * p.w + (xlo + ix);
    const int64_t t = p.target[pix];
    const int pos = colpos[ix];
    if (w != 0.f && t >= 0 && t < p.C && pos < nXp) {
      pw[iy * nXp + pos] = w;
      pl[iy * nXp + pos] = p.lse[pix] * -1.4426950408889634f;
      const int q = __float2int_rn(w * qscale);
      atomicAdd(&hitq[static_cast<int>(t)], q);
      swq += q;
    }
  }
  swq = __reduce_add_sync(0xffffffffu, swq);
  if ((threadIdx.x & 31) == 0 && swq != 0) atomicAdd(&hitq[p.C], swq);
  __syncthreads();
  const float inv_cnt = p.grad_scale / fmaxf(p.count[0], 1.0f);
  const float inv_q = 1.0f / qscale;
  const float* base = p.logits + static_cast<int64_t>(b) * p.batch_stride;
  const float uni = p.eps / static_cast<float>(p.C);
  const float hit = 1.f - p.eps;
  for (int c = threadIdx.x; c < p.d_tok_stride; c += blockDim.x) {
    if (c >= p.C) {
      drow[c] = __float2bfloat16_rn(0.f);
      continue;
    }
    float nb[3][3];  // this class's logits on the 3x3 patch neighbourhood (clamped)
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int yy = min(max(py - 1 + dy, 0), p.hp - 1), xx = min(max(px - 1 + dx, 0), p.wp - 1);
        nb[dy][dx] = __ldg(base + static_cast<int64_t>(yy * p.wp + xx) * p.tok_stride + c);
      }
    float2 g2 = make_float2(0.f, 0.f);
    for (int iy = 0; iy < nY; ++iy) {
      const FootTap rp = rowp[iy];
      if (rp.w == 0.f) continue;  // CTA-uniform
      float a[3];  // log2e * (the two row taps blended), per neighbourhood column
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float top = rp.i0 == 0 ? nb[0][dx] : (rp.i0 == 1 ? nb[1][dx] : nb[2][dx]);
        const float bot = rp.i1 == 0 ? nb[0][dx] : (rp.i1 == 1 ? nb[1][dx] : nb[2][dx]);
        a[dx] = (rp.l0 * top + rp.l1 * bot) * 1.4426950408889634f;
      }
      for (int r = 0; r < nruns; ++r) {
        const int4 run = runs[r];
        const float2 vL = splat2(run.z == 0 ? a[0] : (run.z == 1 ? a[1] : a[2]));
        const float2 vR = splat2(run.w == 0 ? a[0] : (run.w == 1 ? a[1] : a[2]));
        const float4* c0 = reinterpret_cast<const float4*>(cl0 + run.x);
        const float4* c1 = reinterpret_cast<const float4*>(cl1 + run.x);
        const float4* wr = reinterpret_cast<const float4*>(pw + iy * nXp + run.x);
        const float4* lr = reinterpret_cast<const float4*>(pl + iy * nXp + run.x);
        const int n4 = run.y >> 2;
#pragma unroll 2
        for (int i = 0; i < n4; ++i) {
          const float4 A = c0[i], Bc = c1[i], W = wr[i], L = lr[i];
          const float2 v0 = fma2(make_float2(A.x, A.y), vL, fma2(make_float2(Bc.x, Bc.y), vR, make_float2(L.x, L.y)));
          const float2 v1 = fma2(make_float2(A.z, A.w), vL, fma2(make_float2(Bc.z, Bc.w), vR, make_float2(L.z, L.w)));
          g2 = fma2(make_float2(W.x, W.y), make_float2(fast_exp2(v0.x), fast_exp2(v0.y)), g2);
          g2 = fma2(make_float2(W.z, W.w), make_float2(fast_exp2(v1.x), fast_exp2(v1.y)), g2);
        }
      }
    }
    float g = g2.x + g2.y;
    g = fmaf(-hit * inv_q, static_cast<float>(hitq[c]), g);
    g = fmaf(-uni * inv_q, static_cast<float>(hitq[p.C]), g);
    drow[c] = __float2bfloat16_rn(g * inv_cnt);
  }
}


// ----------------------------------------------------------------------------------------
// Validation post-processing (seg_criterion.py:197-213): label propagation over the ResNet-feature
// nearest neighbours.  L2-normalised features -> cosine similarity (batched tcgen05 GEMM, gemm.cu) ->
// top-k neighbours per patch -> resnet_iters rounds of prob[p] = mean_k prob[nbr_k(p)].
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2_normalize_rows_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                                __nv_bfloat16* __restrict__ y, int64_t ldy, int rows, int D) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane * 8; c < D; c += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx + c);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    s += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + cc.x * cc.x + cc.y * cc.y + d.x * d.x + d.y * d.y;
  }
  s = warp_sum(s);
  const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);  // F.normalize eps
  for (int c = lane * 8; c < D; c += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx + c);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    uint4 o;
    o.x = pack_bf16x2(a.x * inv, a.y * inv); o.y = pack_bf16x2(b.x * inv, b.y * inv);
    o.z = pack_bf16x2(cc.x * inv, cc.y * inv); o.w = pack_bf16x2(d.x * inv, d.y * inv);
    *reinterpret_cast<uint4*>(y + static_cast<int64_t>(row) * ldy + c) = o;
  }
}

// top-k (k <= 8) column indices of every row, largest first, lowest index first among equals; one warp per row
template <int K>
__global__ void __launch_bounds__(256) row_topk_kernel(const float* __restrict__ x, int64_t ldx, int rows, int n,
                                                       int32_t* __restrict__ idx_out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float bv[K];
  int bi[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { bv[k] = -INFINITY; bi[k] = 0x7fffffff; }
  const float* r = x + static_cast<int64_t>(row) * ldx;
  for (int j = lane; j < n; j += 32) {
    float v = r[j];
    int id = j;
#pragma unroll
    for (int k = 0; k < K; ++k) {  // insertion into the lane-local sorted list
      const bool better = v > bv[k] || (v == bv[k] && id < bi[k]);
      if (better) {
        const float tv = bv[k]; const int ti = bi[k];
        bv[k] = v; bi[k] = id; v = tv; id = ti;
      }
    }
  }
  // K rounds: the warp-wide best head is emitted, its owner pops it
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float v = bv[0];
    int id = bi[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, id, o);
      if (ov > v || (ov == v && oi < id)) { v = ov; id = oi; }
    }
    if (lane == 0) idx_out[static_cast<int64_t>(row) * K + k] = id;
    if (bi[0] == id) {
#pragma unroll
      for (int q = 0; q + 1 < K; ++q) { bv[q] = bv[q + 1]; bi[q] = bi[q + 1]; }
      bv[K - 1] = -INFINITY; bi[K - 1] = 0x7fffffff;
    }
  }
}

// out[b,p,:] = softmax(x[b,p,:] * inv_temp)   (one warp per row)
__global__ void __launch_bounds__(256) row_softmax_kernel(const float* __restrict__ x, int64_t batch_stride, int64_t tok_stride,
                                                          int P, int C, float inv_temp, float* __restrict__ out, int rows) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* r = x + static_cast<int64_t>(row / P) * batch_stride + static_cast<int64_t>(row % P) * tok_stride;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, r[c] * inv_temp);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += __expf(r[c] * inv_temp - m);
  s = warp_sum(s);
  const float inv = 1.0f / s;
  for (int c = lane; c < C; c += 32) out[static_cast<int64_t>(row) * C + c] = __expf(r[c] * inv_temp - m) * inv;
}

// out[b,p,c] = mean_k in[b, nbr[b,p,k], c]
__global__ void __launch_bounds__(256) gather_mean_kernel(const float* __restrict__ in, const int32_t* __restrict__ nbr, int K,
                                                          int P, int C, int64_t total, float* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = i % C;
  const int64_t bp = i / C;
  const int64_t b = bp / P;
  float s = 0.f;
  for (int k = 0; k < K; ++k) s += in[(b * P + nbr[bp * K + k]) * C + c];
  out[i] = s / static_cast<float>(K);
}

}  // namespace sgf

using namespace sgf;

extern "C" int sgf_l2_normalize_rows(const void* x, int64_t ldx, void* y, int64_t ldy, int32_t rows, int32_t D, void* stream) {
  SGF_REQUIRE(x && y && rows > 0 && D > 0 && D % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "l2_normalize_rows: bad arguments");
  l2_normalize_rows_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<__nv_bfloat16*>(y), ldy, rows, D);
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}

extern "C" int sgf_row_topk(const float* x, int64_t ldx, int32_t rows, int32_t n, int32_t k, int32_t* idx_out, void* stream) {
  SGF_REQUIRE(x && idx_out && rows > 0 && n > 0 && k >= 1 && k <= 8 && k <= n, "row_topk: bad arguments (1 <= k <= 8)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((rows + 7) / 8);
  switch (k) {
    case 1: row_topk_kernel<1><<<grid, 256, 0, st>>>(x, ldx, rows, n, idx_out); break;
    case 2: row_topk_kernel<2><<<grid, 256, 0, st>>>(x, ldx, rows, n, idx_out); break;
    case 3: row_topk_kernel<3><<<grid, 256, 0, st>>>(x, ldx, rows, n, idx_out); break;
    case 4: row_topk_kernel<4><<<grid, 256, 0, st>>>(x, ldx, rows, n, idx_out); break;
    case 5: row_topk_kernel<5><<<grid, 256, 0, st>>>(x, ldx, rows, n, idx_out); break;
    case 6: row_topk_kernel<6><<<grid, 256, 0, st>>>(x, ldx, rows, n, idx_out); break;
    case 7: row_topk_kernel<7><<<grid, 256, 0, st>>>(x, ldx, rows, n, idx_out); break;
    default: row_topk_kernel<8><<<grid, 256, 0, st>>>(x, ldx, rows, n, idx_out); break;
  }
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}

extern "C" int sgf_label_propagation(const float* logits, int64_t batch_stride, int64_t tok_stride, int32_t B, int32_t P,
                                     int32_t C, float temperature, const int32_t* nbr, int32_t k, int32_t iters,
                                     float* prob_a, float* prob_b, void* stream) {
  SGF_REQUIRE(logits && nbr && prob_a && prob_b && B > 0 && P > 0 && C > 0 && k >= 1 && iters >= 0 && temperature > 0.f,
              "label_propagation: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rows = B * P;
  row_softmax_kernel<<<(rows + 7) / 8, 256, 0, st>>>(logits, batch_stride, tok_stride, P, C, 1.0f / temperature, prob_a, rows);
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  const int64_t total = static_cast<int64_t>(rows) * C;
  float* src = prob_a;
  float* dst = prob_b;
  for (int it = 0; it < iters; ++it) {
    gather_mean_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(src, nbr, k, P, C, total, dst);
    SGF_CHECK_CUDA(cudaGetLastError());
    count_launch();
    float* t = src; src = dst; dst = t;
  }
  return SGF_OK;  // result in prob_a when iters is even, prob_b when odd
}


extern "C" int sgf_upsample_ce_loss_bwd(const sgf_segloss_bwd_args* a, void* stream) {
  SGF_REQUIRE(a != nullptr && a->logits && a->target && a->lse && a->count && a->dlogits, "upsample_ce_loss_bwd: null pointer");
  SGF_REQUIRE(a->B > 0 && a->C > 0 && a->hp > 0 && a->wp > 0 && a->h > 0 && a->w > 0, "upsample_ce_loss_bwd: bad shape");
  SGF_REQUIRE(a->d_tok_stride >= a->C && a->d_tokens >= a->hp * a->wp, "upsample_ce_loss_bwd: dlogits too small");
  SeglossBwdParams p{a->logits, a->batch_stride, a->tok_stride, a->B, a->C, a->hp, a->wp, a->h, a->w, a->target,
                     a->lse, a->count, a->label_smoothing, a->grad_scale, reinterpret_cast<__nv_bfloat16*>(a->dlogits),
                     a->d_batch_stride, a->d_tok_stride, a->d_tokens,
                     static_cast<float>(a->hp) / static_cast<float>(a->h),
                     static_cast<float>(a->wp) / static_cast<float>(a->w)};
  int threads = static_cast<int>((a->d_tok_stride + 31) / 32 * 32);
  if (threads > 256) threads = 256;
  // footprint bound: 2 patches of pixels plus the conservative margins; border patches take the clamped rows too
  const int fy_max = 2 * ((a->h + a->hp - 1) / a->hp) + 6 + (a->h + a->hp - 1) / a->hp;
  const int fx_max = 2 * ((a->w + a->wp - 1) / a->wp) + 6 + (a->w + a->wp - 1) / a->wp;
  // run-padded row length: every run of constant tap pair is padded to a multiple of 4 columns; when up-sampling the
  // tap pair changes at most every (w/wp) columns, when down-sampling it may change at every column
  const int max_runs = a->w >= a->wp ? std::min(kCeMaxRuns, fx_max / std::max(1, a->w / a->wp) + 3) : std::min(kCeMaxRuns, fx_max);
  const int nxp_max = (fx_max + 3 * max_runs + 7) & ~3;
  // fixed-point scale of the one-hot histogram: the weights of one footprint sum to < 4 (h/hp)(w/wp)
  const double wsum_bound = 4.0 * ((a->h + a->hp - 1) / a->hp) * ((a->w + a->wp - 1) / a->wp);
  float qscale = 1.0f;
  while (static_cast<double>(qscale) * 2.0 * wsum_bound < 1073741824.0) qscale *= 2.0f;
  const size_t smem = static_cast<size_t>(2) * fy_max * nxp_max * 4 + static_cast<size_t>(2) * nxp_max * 4 +
                      kCeMaxRuns * sizeof(int4) + static_cast<size_t>(fy_max + fx_max) * sizeof(FootTap) +
                      static_cast<size_t>(fx_max) * 4 + static_cast<size_t>(a->C + 1) * 4 + 16;
  SGF_REQUIRE(smem <= 160 * 1024, "upsample_ce_loss_bwd: up-sampling factor too large (%zu B of shared memory)", smem);
  static size_t configured = 0;
  if (smem > configured) {
    SGF_CHECK_CUDA(cudaFuncSetAttribute(upsample_ce_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = 160 * 1024;
  }
  dim3 grid(a->d_tokens, a->B);
  upsample_ce_bwd_kernel<<<grid, threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p, fy_max, fx_max, nxp_max, qscale);
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}


extern "C" int sgf_upsample_ce_loss(const sgf_segloss_args* a, void* stream) {
  SGF_REQUIRE(a != nullptr && a->logits && a->target && a->out, "upsample_ce_loss: null pointer");
  SGF_REQUIRE(a->B > 0 && a->C > 0 && a->hp > 0 && a->wp > 0 && a->h > 0 && a->w > 0, "upsample_ce_loss: bad shape");
  SeglossParams p{a->logits, a->batch_stride, a->tok_stride, a->B, a->C, a->hp, a->wp, a->h, a->w, a->target,
                  a->label_smoothing, a->out, static_cast<float>(a->hp) / static_cast<float>(a->h),
                  static_cast<float>(a->wp) / static_cast<float>(a->w), a->lse_out};
  if (p.scale_w <= 0.25f) {  // upsampling factor >= 4: four pixels per thread share their tap columns
    dim3 block(128), grid((a->w + 511) / 512, a->h, a->B);
    upsample_ce_loss4_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  } else {
    dim3 block(256), grid((a->w + 255) / 256, a->h, a->B);
    upsample_ce_loss_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  }
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}

extern "C" int sgf_upsample_argmax(const sgf_segmask_args* a, void* stream) {
  SGF_REQUIRE(a != nullptr && a->logits && a->mask, "upsample_argmax: null pointer");
  SGF_REQUIRE(a->B > 0 && a->C > 0 && a->hp > 0 && a->wp > 0 && a->h > 0 && a->w > 0, "upsample_argmax: bad shape");
  SGF_REQUIRE(!a->target || (a->area_pred && a->area_label && a->area_intersect),
              "upsample_argmax: target given without histogram outputs");
  SegmaskParams p{a->logits, a->batch_stride, a->tok_stride, a->B, a->C, a->hp, a->wp, a->h, a->w, a->mask,
                  a->target, a->area_intersect, a->area_pred, a->area_label,
                  static_cast<float>(a->hp) / static_cast<float>(a->h),
                  static_cast<float>(a->wp) / static_cast<float>(a->w)};
  dim3 block(256), grid((a->w + 255) / 256, a->h, a->B);
  const size_t smem = a->area_pred ? 3 * a->C * sizeof(float) : 0;
  // default: what ATen executes on a GPU for the reference's input -- the NCHW kernel below 16 channels, the NHWC
  // kernel from 16 channels on (the reference hands F.interpolate a channels-last view, seg_criterion.py:238-240)
  int mode = a->arith;
  if (mode == SGF_LERP_DEFAULT) mode = a->C >= 16 ? SGF_LERP_ATEN_CUDA_NHWC : SGF_LERP_ATEN_CUDA;
  SGF_REQUIRE(mode == SGF_LERP_PLAIN || (mode >= 8 && mode < 16), "upsample_argmax: unknown arith mode %d", mode);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p.scale_w <= 0.25f) {  // upsampling factor >= 4: four pixels per thread share their tap columns (bit-identical masks)
    dim3 block4(128), grid4((a->w + 511) / 512, a->h, a->B);
    switch (mode) {
      case 0: upsample_argmax4_kernel<kArithPlain><<<grid4, block4, smem, st>>>(p); break;
      case 8: upsample_argmax4_kernel<kArithFma + 0><<<grid4, block4, smem, st>>>(p); break;
      case 9: upsample_argmax4_kernel<kArithFma + 1><<<grid4, block4, smem, st>>>(p); break;
      case 10: upsample_argmax4_kernel<kArithFma + 2><<<grid4, block4, smem, st>>>(p); break;
      case 11: upsample_argmax4_kernel<kArithFma + 3><<<grid4, block4, smem, st>>>(p); break;
      case 12: upsample_argmax4_kernel<kArithFma + 4><<<grid4, block4, smem, st>>>(p); break;
      case 13: upsample_argmax4_kernel<kArithFma + 5><<<grid4, block4, smem, st>>>(p); break;
      case 14: upsample_argmax4_kernel<kArithFma + 6><<<grid4, block4, smem, st>>>(p); break;
      default: upsample_argmax4_kernel<kArithFma + 7><<<grid4, block4, smem, st>>>(p); break;
    }
    SGF_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return SGF_OK;
  }
  switch (mode) {
    case 0: upsample_argmax_kernel<kArithPlain><<<grid, block, smem, st>>>(p); break;
    case 8: upsample_argmax_kernel<kArithFma + 0><<<grid, block, smem, st>>>(p); break;
    case 9: upsample_argmax_kernel<kArithFma + 1><<<grid, block, smem, st>>>(p); break;
    case 10: upsample_argmax_kernel<kArithFma + 2><<<grid, block, smem, st>>>(p); break;
    case 11: upsample_argmax_kernel<kArithFma + 3><<<grid, block, smem, st>>>(p); break;
    case 12: upsample_argmax_kernel<kArithFma + 4><<<grid, block, smem, st>>>(p); break;
    case 13: upsample_argmax_kernel<kArithFma + 5><<<grid, block, smem, st>>>(p); break;
    case 14: upsample_argmax_kernel<kArithFma + 6><<<grid, block, smem, st>>>(p); break;
    default: upsample_argmax_kernel<kArithFma + 7><<<grid, block, smem, st>>>(p); break;
  }
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}
