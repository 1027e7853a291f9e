// Real-image side of the input pipeline on the device (data/mm_data/segmentation_dataset.py:210-301).
// The reference decodes a JPEG on the host and then runs, per sample and on one CPU thread (--num-workers=0),
//   mmseg Resize (mmcv.imrescale -> cv2.resize INTER_LINEAR on uint8)  ->  [RandomCrop -> RandomFlip]  ->
//   ToTensor (/255)  ->  Normalize(mean, std)                                   (:236-262, :155-156)
// and for the label map: remap (0 -> ignore, k -> k-1; :224-227), cv2 INTER_NEAREST resize, seg2code gather, and a
// torchvision NEAREST resize to the patch grid for prev_output_tokens (:246-262).
// Here the decoded uint8 image crosses PCIe as it is (1 byte per value instead of 4) and one kernel writes the
// normalised fp32 CHW tensor the model takes.  Everything is integer or IEEE fp32 arithmetic in the reference's order,
// so the result is bit-identical to cv2 + torchvision:
//   * cv2 INTER_LINEAR, 8-bit: 11-bit fixed-point coefficients (cvRound(c * 2048)), horizontal pass in int32,
//     vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2 >> 2; the x coefficients are clamped at the
//     borders, the y coefficients are not (the row index is); an exact 2x decimation takes the INTER_AREA 2x2 mean.
//   * source coordinates: fx = (float)((d + 0.5) * scale - 0.5) in double, scale = 1 / (dst / src).
// PhotoMetricDistortion (random, cv2 colour conversions) is not reproduced: training samples that need it stay on the
// reference's dataset; the validation path (MultiScaleFlipAug, keep-ratio resize, no flip) is complete.
// HBM-bound byte work: one thread per output pixel (3 channels), 4 source pixels read (L1/L2-resident rows).
#include <stdint.h>

#include "../../include/segofa_b200.h"
#include "common.cuh"

namespace sgf {

struct Tap {
  int i0, i1;  // clamped source indices
  int c0, c1;  // 11-bit fixed-point weights
};

// One axis of cv2's linear resize table (resize.cpp: resize() builds xofs/ialpha and yofs/ibeta this way).
SGF_DEVICE Tap linear_tap(int d, int src, double scale, bool clamp_weights) {
  const float f0 = static_cast<float>(__dsub_rn(__dmul_rn(static_cast<double>(d) + 0.5, scale), 0.5));
  int s = static_cast<int>(floorf(f0));
  float f = __fsub_rn(f0, static_cast<float>(s));
  if (clamp_weights) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  Tap t;
  t.c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  t.c1 = __float2int_rn(__fmul_rn(f, 2048.f));
  t.i0 = min(max(s, 0), src - 1);
  t.i1 = min(max(s + 1, 0), src - 1);
  return t;
}

__global__ void __launch_bounds__(256) resize_normalize_u8_kernel(const sgf_image_prep_args a, double scale_x, double scale_y,
                                                                  int area2x) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= a.out_w) return;
  const int rx = a.crop_x + (a.flip ? a.out_w - 1 - x : x);  // column of the resized image this output pixel shows
  const int ry = a.crop_y + y;
  int v[3];
  if (area2x) {
    const uint8_t* r0 = a.src + static_cast<int64_t>(2 * ry) * a.src_row_stride + 6 * rx;
    const uint8_t* r1 = r0 + a.src_row_stride;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (r0[c] + r0[3 + c] + r1[c] + r1[3 + c] + 2) >> 2;
  } else {
    const Tap tx = linear_tap(rx, a.src_w, scale_x, true);
    const Tap ty = linear_tap(ry, a.src_h, scale_y, false);
    const uint8_t* r0 = a.src + static_cast<int64_t>(ty.i0) * a.src_row_stride;
    const uint8_t* r1 = a.src + static_cast<int64_t>(ty.i1) * a.src_row_stride;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int s0 = r0[3 * tx.i0 + c] * tx.c0 + r0[3 * tx.i1 + c] * tx.c1;
      const int s1 = r1[3 * tx.i0 + c] * tx.c0 + r1[3 * tx.i1 + c] * tx.c1;
      const int o = (((ty.c0 * (s0 >> 4)) >> 16) + ((ty.c1 * (s1 >> 4)) >> 16) + 2) >> 2;
      v[c] = min(max(o, 0), 255);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // ToTensor: float(u8) / 255; Normalize: (t - mean) / std -- IEEE fp32, each step rounded
    const float t = __fdiv_rn(static_cast<float>(v[c]), 255.f);
    a.dst[c * a.dst_channel_stride + static_cast<int64_t>(y) * a.dst_row_stride + x] =
        __fdiv_rn(__fsub_rn(t, a.mean[c]), a.std[c]);
  }
}

// label remap of segmentation_dataset.py:224-227 (uint8 arithmetic): 0 -> 255, then -1, then 254 -> num_seg
SGF_DEVICE int remap_label(uint8_t raw, int num_seg) {
  const uint8_t v = static_cast<uint8_t>((raw == 0 ? 255 : raw) - 1);
  return v == 254 ? num_seg : v;
}

__global__ void __launch_bounds__(256) segmap_targets_kernel(const sgf_segmap_prep_args a, double ifx, double ify,
                                                             float down_sy, float down_sx) {
  const int64_t n_out = static_cast<int64_t>(a.out_h) * a.out_w;
  const int64_t n_grid = static_cast<int64_t>(a.grid_h) * a.grid_w;
  const int64_t n_ori = a.ori_classes ? static_cast<int64_t>(a.src_h) * a.src_w : 0;
  const int64_t total = n_out + 1 + n_grid + 1 + n_ori;
  // class id at pixel (y, x) of the augmented (resized, cropped, flipped) label map: cv2 INTER_NEAREST
  auto label_at = [&](int y, int x) {
    const int rx = a.crop_x + (a.flip ? a.out_w - 1 - x : x);
    const int ry = a.crop_y + y;
    const int sx = min(static_cast<int>(floor(__dmul_rn(static_cast<double>(rx), ifx))), a.src_w - 1);
    const int sy = min(static_cast<int>(floor(__dmul_rn(static_cast<double>(ry), ify))), a.src_h - 1);
    return remap_label(a.src[static_cast<int64_t>(sy) * a.src_row_stride + sx], a.num_seg);
  };
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (i < n_out) {
      a.target[i] = a.seg_id_offset + label_at(static_cast<int>(i / a.out_w), static_cast<int>(i % a.out_w));
    } else if (i == n_out) {
      a.target[i] = a.eos_id;
    } else if (i <= n_out + 1 + n_grid) {
      const int64_t g = i - n_out - 1;  // slot 0 = bos, then the patch grid
      int64_t code = a.bos_id;
      if (g > 0) {
        // torchvision Resize(NEAREST) = ATen 'nearest': src = min((int)floorf(dst * (float)in / out), in - 1)
        const int gy = static_cast<int>((g - 1) / a.grid_w), gx = static_cast<int>((g - 1) % a.grid_w);
        const int py = min(static_cast<int>(floorf(__fmul_rn(static_cast<float>(gy), down_sy))), a.out_h - 1);
        const int px = min(static_cast<int>(floorf(__fmul_rn(static_cast<float>(gx), down_sx))), a.out_w - 1);
        code = a.seg_id_offset + label_at(py, px);
        if (a.downsampled_target) a.downsampled_target[g - 1] = code;
      }
      a.prev_output_tokens[g] = code;
      if (g == n_grid && a.downsampled_target) a.downsampled_target[n_grid] = a.eos_id;
    } else {
      const int64_t o = i - (n_out + 1 + n_grid + 1);
      const int oy = static_cast<int>(o / a.src_w), ox = static_cast<int>(o % a.src_w);
      a.ori_classes[o] = remap_label(a.src[static_cast<int64_t>(oy) * a.src_row_stride + ox], a.num_seg);
    }
  }
}

}  // namespace sgf

using namespace sgf;

static bool window_ok(int rs_h, int rs_w, int crop_y, int crop_x, int out_h, int out_w) {
  return rs_h > 0 && rs_w > 0 && out_h > 0 && out_w > 0 && crop_y >= 0 && crop_x >= 0 && crop_y + out_h <= rs_h &&
         crop_x + out_w <= rs_w;
}

extern "C" int sgf_image_prep_u8(const sgf_image_prep_args* a, void* stream) {
  SGF_REQUIRE(a && a->src && a->dst && a->src_h > 0 && a->src_w > 0 && a->src_row_stride >= 3 * static_cast<int64_t>(a->src_w),
              "image_prep: bad source");
  SGF_REQUIRE(window_ok(a->rs_h, a->rs_w, a->crop_y, a->crop_x, a->out_h, a->out_w),
              "image_prep: the crop window must lie inside the resized image");
  SGF_REQUIRE(a->std[0] != 0.f && a->std[1] != 0.f && a->std[2] != 0.f, "image_prep: std must be non-zero");
  SGF_REQUIRE(a->dst_row_stride >= a->out_w && a->dst_channel_stride >= static_cast<int64_t>(a->out_h) * a->out_w,
              "image_prep: destination strides too small");
  // cv::resize: inv_scale = dsize / ssize (double), scale = 1 / inv_scale
  const double scale_x = 1.0 / (static_cast<double>(a->rs_w) / a->src_w);
  const double scale_y = 1.0 / (static_cast<double>(a->rs_h) / a->src_h);
  const int area2x = (a->src_w == 2 * a->rs_w && a->src_h == 2 * a->rs_h) ? 1 : 0;
  dim3 grid((a->out_w + 255) / 256, a->out_h);
  resize_normalize_u8_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*a, scale_x, scale_y, area2x);
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}

extern "C" int sgf_segmap_prep_u8(const sgf_segmap_prep_args* a, void* stream) {
  SGF_REQUIRE(a && a->src && a->target && a->prev_output_tokens && a->src_h > 0 && a->src_w > 0 &&
                  a->src_row_stride >= a->src_w && a->num_seg > 0 && a->num_seg < 255,
              "segmap_prep: bad arguments");
  SGF_REQUIRE(window_ok(a->rs_h, a->rs_w, a->crop_y, a->crop_x, a->out_h, a->out_w),
              "segmap_prep: the crop window must lie inside the resized map");
  SGF_REQUIRE(a->grid_h > 0 && a->grid_w > 0, "segmap_prep: bad patch grid");
  const double ifx = 1.0 / (static_cast<double>(a->rs_w) / a->src_w);
  const double ify = 1.0 / (static_cast<double>(a->rs_h) / a->src_h);
  const float down_sy = static_cast<float>(a->out_h) / static_cast<float>(a->grid_h);
  const float down_sx = static_cast<float>(a->out_w) / static_cast<float>(a->grid_w);
  const int64_t total = static_cast<int64_t>(a->out_h) * a->out_w + static_cast<int64_t>(a->grid_h) * a->grid_w + 2 +
                        (a->ori_classes ? static_cast<int64_t>(a->src_h) * a->src_w : 0);
  const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 8));
  segmap_targets_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*a, ifx, ify, down_sy, down_sx);
  SGF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return SGF_OK;
}
