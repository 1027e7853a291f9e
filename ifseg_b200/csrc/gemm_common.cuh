// Types shared by the GEMM translation units (gemm.cu: one-tile-per-CTA and mixed-major kernels; gemm_ts.cu: the
// persistent CTA-pair kernel with the TMA-store epilogue).
#pragma once
#include "common.cuh"

namespace sgf {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
static constexpr int UMMA_K = 16;
template <int BN>
static constexpr int gemm_epi_warps() { return BN >= 128 ? 8 : 4; }
template <int BN>
static constexpr int gemm_threads() { return 64 + 32 * gemm_epi_warps<BN>(); }

struct GemmEpilogue {
  void* c;
  int64_t ldc, c_batch_stride;
  int c_dtype;
  const float* col_scale;
  const float* col_bias;
  const void* residual;
  int64_t ldr, r_batch_stride;
  int r_dtype;
  int act;
  float alpha;
  int alpha_cols;
  float* rowstats_out;
  const float* rownorm_stats;
  const float* rownorm_u;
  float rownorm_inv_dim;
  int rownorm_parts;
};

struct GemmShape {
  int M, N, K;
  // conv mode only: H, W = OUTPUT height / width (the 128-pixel tiles are bw x bh rectangles of the output)
  int H, W, bw, bh, tiles_w, tiles_h, cin_blocks;
  // general convolution (gemm_ts.cu): filter width, stride and padding; the one-tile kernel is 3x3 / 1 / 1 only
  int kw, stride_h, stride_w, pad_h, pad_w;
};

// input side of a convolution launch of gemm_ts.cu: a 4-D TMA view (c, w, h, n) with explicit element strides
struct ConvInput {
  const void* x;
  int n, h, w, cin;                         // tensor-map extents along n, h, w and the 64-multiple innermost extent
  int64_t w_stride, h_stride, n_stride;     // in bf16 elements (w_stride < cin = overlapping windows is allowed)
};

template <int BN>
struct GemmSmem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
};

// Epilogue feature set.  Specialised kernels carry the flags as a template argument so that each
// instantiation contains only its own code path (the all-runtime kernel was 130 KB of SASS and
// stalled on instruction fetch: every CTA runs its epilogue exactly once); kEpiRuntime keeps the
// fully general path for odd shapes (N not a multiple of 32, rare operand combinations).
enum : int {
  kEpiScale = 1, kEpiBias = 2, kEpiAlpha = 4, kEpiGelu = 8, kEpiRelu = 16, kEpiResBf16 = 32, kEpiResF32 = 64,
  kEpiOutF32 = 128, kEpiRowStats = 256, kEpiRowNorm = 512, kEpiAtomic = 1024, kEpiRuntime = 1 << 15
};
template <int kEpi, int kFlag>
SGF_DEVICE bool epi_has(bool runtime_value) {
  if constexpr ((kEpi & kEpiRuntime) != 0) return runtime_value;
  else return (kEpi & kFlag) != 0;
}


// feature mask of a call; specialised kernels exist for the combinations the segofa path uses
static inline int epilogue_mask(const GemmEpilogue& ep) {
  int m = 0;
  if (ep.col_scale) m |= kEpiScale;
  if (ep.col_bias) m |= kEpiBias;
  if (ep.alpha_cols > 0) m |= kEpiAlpha;
  if (ep.act == SGF_ACT_GELU) m |= kEpiGelu;
  if (ep.act == SGF_ACT_RELU) m |= kEpiRelu;
  if (ep.residual) m |= (ep.r_dtype == SGF_F32 ? kEpiResF32 : kEpiResBf16);
  if (ep.c_dtype == SGF_F32) m |= kEpiOutF32;
  if (ep.rowstats_out) m |= kEpiRowStats;
  if (ep.rownorm_stats) m |= kEpiRowNorm;
  return m;
}

// gemm_ts.cu: returns -1 when the call is outside the kernel's envelope (the caller falls back to the tile kernels)
int gemm_ts_dispatch(const GemmShape& shp, const GemmEpilogue& ep, const void* a, int64_t lda, const void* b, int64_t ldb,
                     const ConvInput* conv, cudaStream_t st);

}  // namespace sgf
