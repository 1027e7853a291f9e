// Patch-grid -> mask: bilinear x(h/hp) upsample fused with the class argmax (and optional
// per-class area histograms).  The full-resolution [B,h,w,C] logits that the reference
// materialises (seg_criterion.py:237-244; 138 MB per 480x480 image at C=150) never exist:
// each thread evaluates the C interpolated logits of its pixel in registers.  HBM traffic is
// the low-res logits (L2-resident) in and one int64 per pixel out.
//
// Arithmetic mirrors ATen's upsample_bilinear2d(align_corners=False): source index
// s = max(0, scale*(d+0.5)-0.5), i0 = (int)s, i1 = i0 + (i0 < in-1), l1 = s - i0, l0 = 1 - l1,
// value = h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11), evaluated without FMA contraction.
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

SGF_DEVICE void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = __fsub_rn(__fmul_rn(scale, static_cast<float>(dst) + 0.5f), 0.5f);
  s = s < 0.f ? 0.f : s;
  i0 = static_cast<int>(s);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = __fsub_rn(s, static_cast<float>(i0));
  l0 = __fsub_rn(1.0f, l1);
}

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
    src_index(p.scale_h, y, p.hp, y0, y1, hl0, hl1);
    src_index(p.scale_w, x, p.wp, x0, x1, wl0, wl1);
    const float* base = p.logits + static_cast<int64_t>(b) * p.batch_stride;
    const float* p00 = base + static_cast<int64_t>(y0 * p.wp + x0) * p.tok_stride;
    const float* p01 = base + static_cast<int64_t>(y0 * p.wp + x1) * p.tok_stride;
    const float* p10 = base + static_cast<int64_t>(y1 * p.wp + x0) * p.tok_stride;
    const float* p11 = base + static_cast<int64_t>(y1 * p.wp + x1) * p.tok_stride;
    float best = -INFINITY;
    int best_c = 0;
    for (int c = 0; c < p.C; ++c) {
      const float top = __fadd_rn(__fmul_rn(wl0, __ldg(p00 + c)), __fmul_rn(wl1, __ldg(p01 + c)));
      const float bot = __fadd_rn(__fmul_rn(wl0, __ldg(p10 + c)), __fmul_rn(wl1, __ldg(p11 + c)));
      const float v = __fadd_rn(__fmul_rn(hl0, top), __fmul_rn(hl1, bot));
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

struct SeglossParams {
  const float* logits;
  int64_t batch_stride, tok_stride;
  int B, C, hp, wp, h, w;
  const int64_t* target;
  float eps;
  float* out;
  float scale_h, scale_w;
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

}  // namespace sgf

using namespace sgf;

extern "C" int sgf_upsample_ce_loss(const sgf_segloss_args* a, void* stream) {
  SGF_REQUIRE(a != nullptr && a->logits && a->target && a->out, "upsample_ce_loss: null pointer");
  SGF_REQUIRE(a->B > 0 && a->C > 0 && a->hp > 0 && a->wp > 0 && a->h > 0 && a->w > 0, "upsample_ce_loss: bad shape");
  SeglossParams p{a->logits, a->batch_stride, a->tok_stride, a->B, a->C, a->hp, a->wp, a->h, a->w, a->target,
                  a->label_smoothing, a->out, static_cast<float>(a->hp) / static_cast<float>(a->h),
                  static_cast<float>(a->wp) / static_cast<float>(a->w)};
  dim3 block(256), grid((a->w + 255) / 256, a->h, a->B);
  upsample_ce_loss_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
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
  upsample_argmax_kernel<<<grid, block, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}
