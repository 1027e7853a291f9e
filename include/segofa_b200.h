/* segofa_b200 -- C ABI of the B200-native segofa hot path (libsegofa_b200.so).
 *
 * The reference (alinlab/ifseg) has NO native/FFI layer on this path: every op below is a
 * PyTorch ATen call issued from Python.  Each entry point therefore cites
 * the reference *Python call site* whose ATen ops it replaces; the binding a maintainer adds
 * on the reference side is a ctypes stub.
 *
 * Conventions
 *  - plain pointers + sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *  - every function takes the cudaStream_t to launch on as `void* stream` and is asynchronous;
 *  - return value: 0 = SGF_OK, otherwise an SGF_ERR_* code; sgf_last_error() gives the message
 *    (thread-local).  SGF_ERR_OOM messages contain "out of memory" so that the Python wrapper
 *    raises the RuntimeError that trainer.py:807-822 knows how to recover from;
 *  - no hidden allocations: scratch is passed in by the caller;
 *  - dtype enums: SGF_BF16 = 0, SGF_F32 = 1.
 */
#ifndef SEGOFA_B200_H_
#define SEGOFA_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SGF_OK = 0, SGF_ERR_INVALID = 1, SGF_ERR_CUDA = 2, SGF_ERR_OOM = 3, SGF_ERR_UNSUPPORTED = 4 };
enum { SGF_BF16 = 0, SGF_F32 = 1 };
enum { SGF_ACT_NONE = 0, SGF_ACT_RELU = 1, SGF_ACT_GELU = 2 };

const char* sgf_last_error(void);
int sgf_abi_version(void);
/* number of kernels launched by this library in the calling process since load / last reset */
int64_t sgf_launch_count(void);
void sgf_reset_launch_count(void);
/* Debug / measurement hooks (not on the product path; NULL switches them off again):
 *  - sgf_debug_set_gemm_trace: the next sgf_gemm_bf16 / sgf_conv* launches taken by the persistent CTA-pair kernel record a
 *    %globaltimer timeline of their first and last cluster into buf (2 x 8 roles x 64 events of uint64; tools/trace_gemm.py)
 *  - sgf_debug_set_attention_trace: sgf_attention_bf16 records a clock64 timeline of four CTAs into buf
 *    (4 x 4 roles x 32 tiles x 4 events of uint64; tools/trace_attn.py)
 * Environment switches read at launch time: SGF_GEMM_FAMILY = tile | ts, SGF_GEMM_TS_BN = 64..384, SGF_GEMM_TS_PRODUCERS = 1 | 2
 * (A/B measurements, tools/bench_gemm.py), SGF_NO_PDL (plain stream order instead of programmatic dependent launch). */
void sgf_debug_set_gemm_trace(void* buf);
void sgf_debug_set_attention_trace(void* buf);

/* ---------------------------------------------------------------------------------------
 * Dense contraction  C[z] = epilogue( A[z] (MxK, K-major) * B[z]^T (NxK, K-major) )
 *   tcgen05.mma (kind::f16, bf16 in / fp32 accumulate in TMEM), TMA-staged 128B-swizzled tiles.
 * epilogue(v)[m,n]:  [row-norm: v = rstd[m] * (v - mean[m] * rownorm_u[n])]
 *                    v *= col_scale[n]; v += col_bias[n]; if n < alpha_cols: v *= alpha;
 *                    GELU (act==2); v += residual[m,n]; ReLU (act==1); store as c_dtype;
 *                    [row-stats: rowstats_out[m] += (sum, sum of squares) of the STORED row segment]
 * The two bracketed options fold a LayerNorm that sits between two GEMMs into their epilogues
 * (LN(f) W^T = rstd (f (gamma.W)^T - mean * sum_k gamma_k W_nk) + W beta): the producer GEMM
 * writes per-row (sum, sumsq) of every 64-column block of its bf16 output to rowstats_out
 * [M, N/64, 2] (plain stores, no atomics: results are run-to-run deterministic), the consumer GEMM
 * reads them as rownorm_stats [M, rownorm_dim/64, 2], sums the blocks in order, with rownorm_dim =
 * row length of the normalised operand and rownorm_u[n] = sum_k W'[n,k].  Used for ffn_layernorm
 * (unify_transformer_layer.py:281-283, 558-560): the [M, F] LayerNorm pass disappears.
 * Replaces torch.nn.functional.linear / addmm at:
 *   models/segofa/unify_multihead_attention.py:328-345,513 (q/k/v/out_proj, q *= scaling :346)
 *   models/segofa/unify_transformer_layer.py:279-283,556-560 (fc1+gelu, fc2 + residual :289,566)
 *   models/segofa/encoder_module.py:416 (image_proj), :765-771 (pos_q/pos_k and their product)
 *   models/segofa/decoder_module.py:290-294 (seg_projection), :350-364 (self/cross pos bias)
 *   models/segofa/resnet.py:117-137 (1x1 convolutions as GEMMs over NHWC pixels, with the
 *   FrozenBatchNorm2d affine frozen_bn.py:40-45, ReLU and the residual add in the epilogue)
 * ------------------------------------------------------------------------------------- */
typedef struct {
  const void* a; int64_t lda; int64_t a_batch_stride; /* bf16, strides in elements */
  const void* b; int64_t ldb; int64_t b_batch_stride; /* bf16 [N,K] */
  void* c; int64_t ldc; int64_t c_batch_stride; int32_t c_dtype;
  int32_t M, N, K, batch;
  const float* col_scale; /* [N] or NULL */
  const float* col_bias;  /* [N] or NULL */
  const void* residual; int64_t ldr; int64_t r_batch_stride; int32_t r_dtype; /* [M,N] or NULL */
  int32_t act;
  float alpha; int32_t alpha_cols;
  float* rowstats_out;          /* [M, N/64, 2] fp32 or NULL */
  const float* rownorm_stats;   /* [M, rownorm_dim/64, 2] fp32 or NULL */
  const float* rownorm_u;       /* [N] fp32 (required with rownorm_stats) */
  int32_t rownorm_dim;
} sgf_gemm_args;
int sgf_gemm_bf16(const sgf_gemm_args* args, void* stream);
/* Mixed-major variant for the dense adjoints of the training path (no transposed copies):
 *   C[M,N] (+)= A_eff[M,K] * B_eff[N,K]^T   with  A_eff[m,k] = a_mn_major ? a[k*lda + m] : a[m*lda + k]
 *                                                B_eff[n,k] = b_mn_major ? b[k*ldb + n] : b[n*ldb + k]
 * (a_mn, b_mn) = (1,1): dW = dY^T X over the token dimension;  (0,1): dX = dY W with W as stored [N_out,K_in].
 * Plain epilogue only (no bias/activation; batch = 1; residual == c requests in-place accumulation).  split_k > 1 splits the contraction over
 * blockIdx.z and REDUCES the partial tiles into C with vector fp32 atomics: C must be fp32 and hold the value to
 * accumulate onto (zeros, or a running gradient for gradient accumulation).  N % 32 == 0. */
int sgf_gemm_bf16_ex(const sgf_gemm_args* args, int32_t a_mn_major, int32_t b_mn_major, int32_t split_k, void* stream);

/* ---------------------------------------------------------------------------------------
 * 3x3 / stride 1 / pad 1 convolution as an im2col-free implicit GEMM over NHWC bf16:
 * for every filter tap the A tile is a TMA box of the input shifted by the tap offset
 * (out-of-bounds pixels are zero-filled by TMA == the zero padding).  Same tcgen05 main loop
 * and epilogue as sgf_gemm_bf16 (BN affine + ReLU).
 * Replaces nn.Conv2d(3x3) + FrozenBatchNorm2d + ReLU at models/segofa/resnet.py:122-124.
 *   x   [N,H,W,Cin]  bf16      w [Cout, 3,3, Cin] bf16 (tap-major K)      y [N,H,W,Cout] bf16
 * ------------------------------------------------------------------------------------- */
typedef struct {
  const void* x; const void* w; void* y;
  int32_t n, h, w_, cin, cout;
  const float* col_scale; const float* col_bias; int32_t act;
} sgf_conv3x3_args;
int sgf_conv3x3_s1_nhwc(const sgf_conv3x3_args* args, void* stream);

/* ---------------------------------------------------------------------------------------
 * General im2col-free convolution + FrozenBatchNorm affine (+ ReLU) over bf16 NHWC as an implicit GEMM on the persistent
 * CTA-pair kernel: for filter tap (ky,kx) and 64-channel block the A tile is ONE 4-D TMA box of the input; a strided
 * convolution samples every stride-th pixel through the TMA element strides, out-of-bounds coordinates are zero fill
 * (= the padding).  No patch matrix is ever written.
 *   x   : a (c, w, h, n) view given by extents (cin % 64 == 0, w_, h, n) and ELEMENT strides x_w_stride / x_h_stride /
 *         x_n_stride (multiples of 8).  Plain NHWC: x_w_stride = cin, x_h_stride = cin*W, x_n_stride = cin*W*H.
 *         x_w_stride < cin describes overlapping windows: the 7x7/2 stem convolution over a zero-padded [N,Hp,Wp,8] image
 *         is kh = 7, kw = 1, cin = 64 (8 pixels x 8 channels = one 128-byte window), x_w_stride = 16 (two pixels),
 *         stride_h = 2, stride_w = 1, pad = 0 (the padding is physically there) -- see sgf_nchw_f32_to_nhwc8_padded.
 *   w   : [cout, kh*kw*cin] bf16, tap-major (ky, kx, c);   y : [n, ho, wo, cout] bf16 contiguous;   cout % 64 == 0
 *   y[n,oh,ow,:] = act( scale * sum_{ky,kx,c} x[n, oh*stride_h + ky - pad_h, ow*stride_w + kx - pad_w, c] w[:,ky,kx,c] + bias )
 * Replaces nn.Conv2d + FrozenBatchNorm2d (+ ReLU) at models/segofa/resnet.py:117-137 (3x3/1, 3x3/2), :161-170 (the 1x1/2
 * downsample convolutions) and :189-213 (conv1 7x7/2).
 * ------------------------------------------------------------------------------------- */
typedef struct {
  const void* x; int32_t n, h, w_, cin;
  int64_t x_w_stride, x_h_stride, x_n_stride;
  const void* w; void* y; int32_t ho, wo, cout;
  int32_t kh, kw, stride_h, stride_w, pad_h, pad_w;
  const float* col_scale; const float* col_bias; int32_t act;
} sgf_conv2d_args;
int sgf_conv2d_nhwc(const sgf_conv2d_args* args, void* stream);

/* ---------------------------------------------------------------------------------------
 * Stem helpers (HBM-bound layout/gather kernels), models/segofa/resnet.py:215-220:
 *  - sgf_nchw_f32_to_nhwc_bf16: patch_images [N,3,H,W] fp32 -> [N,H,W,C] bf16
 *  - sgf_nchw_f32_to_nhwc8_padded: the input layout of the im2col-free 7x7/2 conv1 (sgf_conv2d_nhwc window mode)
 *    (no patch matrix is written anywhere on the path: the r01 sgf_im2col_nhwc entry point is gone)
 *  - sgf_maxpool3x3s2_nhwc: nn.MaxPool2d(3,2,1)
 * ------------------------------------------------------------------------------------- */
int sgf_nchw_f32_to_nhwc_bf16(const float* x, void* y, int32_t n, int32_t c, int32_t h, int32_t w, void* stream);
/* patch_images [N,C<=8,H,W] fp32 -> zero-padded [N, hp, wp, 8] bf16 with the image at offset (pad, pad): channels C..7 and
 * the border are written as zeros (the whole buffer is written, it needs no clearing) */
int sgf_nchw_f32_to_nhwc8_padded(const float* x, void* y, int32_t n, int32_t c, int32_t h, int32_t w, int32_t pad,
                                 int32_t hp, int32_t wp, void* stream);
int sgf_maxpool3x3s2_nhwc(const void* x, void* y, int32_t n, int32_t h, int32_t w, int32_t c, int32_t ho,
                          int32_t wo, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused row kernel (warp-shuffle LayerNorm, 128-bit loads), one warp per row of width D:
 *     t    = x[r] (+ pre_add[d])                         x: bf16 or fp32
 *     u    = ln1 ? LN(t; g1,b1) : t
 *     v    = u (+ residual[r'])                          residual/out1 share the output row map
 *     out1[r'] = v            (optional, dtype out1_dtype)
 *     out2[r'] = LN(v; g2,b2) (optional, bf16)
 * Output row map: r' = (r / seg_len) * seg_stride + seg_off + r % seg_len  (seg_len==0: r'=r),
 * which writes straight into a concatenated [B, T, D] buffer (no torch.cat copy).
 * If `gather_idx` is given, input row r is x[gather_idx[r]] (token embedding lookup).
 * If `zero_row` is given and zero_row[r] != 0, v is forced to zero (padding rows, encoder_module.py:751-752);
 * out2 is then LN(0) = beta, exactly what the reference's next LayerNorm sees.
 * LayerNorm eps 1e-5, fp32 statistics (custom_fairseq/fairseq/modules/layer_norm.py:30-35).
 * Replaces LayerNorm/residual/dropout(p=0)/cat at unify_transformer_layer.py:256-291,463-568;
 * encoder_module.py:400-428,751-752,757-760,829-830; decoder_module.py:537,575-576,668-669.
 * ------------------------------------------------------------------------------------- */
typedef struct {
  const void* x; int64_t ldx; int32_t x_dtype;
  const int64_t* gather_idx;
  const float* pre_add;
  const float* g1; const float* b1;
  const void* residual; int64_t ldr; int32_t r_dtype;
  void* out1; int64_t ld1; int32_t out1_dtype;
  const float* g2; const float* b2;
  void* out2; int64_t ld2;
  const uint8_t* zero_row;
  int32_t rows, D;
  int32_t seg_len, seg_stride, seg_off;
  float* clear_rowstats; /* optional [rows,2] fp32 scratch zeroed as a side effect (row-stats of a following GEMM) */
  int32_t x_act;         /* SGF_ACT_GELU: t = gelu(x[r]) (the FFN's gelu -> ffn_layernorm pair of the training path) */
  /* training-mode dropout (fairseq_dropout.py) and DropPath (unify_transformer_layer.py:19-35) applied to u before the
   * residual add: v = u * keep_elem/(1-drop_p) * keep_path[sample]/(1-droppath_p) + residual.  Masks are a counter-based
   * hash of (drop_seed, drop_step[0], drop_site, output row, column): the adjoint kernel regenerates them, a CUDA-graph
   * replay with an incremented device-side step draws new ones.  sample = output row / rows_per_sample. */
  float drop_p; float droppath_p; uint32_t drop_seed; uint32_t drop_site; int32_t rows_per_sample;
  const int32_t* drop_step;
} sgf_rowln_args;
int sgf_row_layernorm(const sgf_rowln_args* args, void* stream);

/* ---------------------------------------------------------------------------------------
 * Attention bias assembly (batch-invariant, parameter-only), one launch per layer:
 *   out[h,i,j] = abs[h,i,j] + sum_{blocks b with lo_b <= i,j < hi_b}
 *                               table_b[ bucket_b[ ids_b[i-lo_b], ids_b[j-lo_b] ], h ]      (fp32)
 * `abs` is the layer-independent q_pos k_pos^T term (a batched sgf_gemm_bf16 over heads); the
 * blocks are the image-image and text-text squares of the encoder, or the whole seg grid of
 * the decoder.  out may alias abs.  Columns j >= Tk of the padded row stride are not touched.
 * Replaces the per-layer clone + F.embedding + slice-add at encoder_module.py:313-331,790-809
 * and decoder_module.py:327-333,601-627 for the identity-interpolation case (actual grid ==
 * orig/seg grid, true for every BASELINE config); for other grids the host precomputes the
 * interpolated table once and passes it as `dense_add` ([H,Tq,Tk] fp32, same strides as abs).
 * ------------------------------------------------------------------------------------- */
typedef struct {
  const int64_t* bucket; int64_t bucket_ld; /* int64 [*, bucket_ld] */
  const int64_t* ids;                       /* [hi-lo] position ids into bucket rows/cols */
  const float* table;                       /* fp32 [num_rel, H] */
  int32_t lo, hi;
} sgf_relblock;
typedef struct {
  float* out; const float* abs; int64_t head_stride; int64_t row_stride; /* [H,Tq,row_stride] fp32 */
  const float* dense_add;                                                /* optional */
  int32_t H, Tq, Tk;
  int32_t num_blocks; sgf_relblock blocks[2];
  void* out_f16; /* optional fp16 [H,Tq,row_stride] (same strides): what sgf_attention_bf16 reads.  When given, `out` may
                    be NULL, and if not NULL it receives float(half(value)) -- bit-identical to the fp16 copy (the
                    adjoint kernels read fp32) */
} sgf_bias_args;
int sgf_build_attn_bias(const sgf_bias_args* args, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused multi-head attention, head_dim 64:  O = softmax(Q K^T + bias + mask) V * head_scale
 *   QK^T and PV on tcgen05 (S and O tiles in TMEM), K/V tiles by TMA, fp32 online softmax by the
 *   row-owning threads, probabilities re-staged to shared memory as the bf16 A operand of PV.
 *   q is expected pre-scaled (the QKV GEMM epilogue applies (2*d_h)^-1/2 to the q columns).
 *   Element (b,t,h,d) of q/k/v/out lives at base + b*batch_stride + t*row_stride + h*64 + d.
 *   bias FP16 [H,Tq,row_stride] (batch-invariant; row stride a multiple of 8, streamed through double-buffered TMA
 *   tiles: half the L2 traffic of fp32 and room for a deeper K/V prefetch) or NULL; key_padding_mask uint8 [B,Tk] (1 = masked) or
 *   NULL; causal != 0 masks j > i (decoder_module.py:878-890).
 * Replaces unify_multihead_attention.py:459-512 (bmm, bias add, masks, fp32 softmax, bmm,
 * c_attn scale).
 * ------------------------------------------------------------------------------------- */
typedef struct {
  const void* q; int64_t q_row_stride; int64_t q_batch_stride;
  const void* k; int64_t k_row_stride; int64_t k_batch_stride;
  const void* v; int64_t v_row_stride; int64_t v_batch_stride;
  void* out; int64_t o_row_stride; int64_t o_batch_stride;
  const void* bias; int64_t bias_head_stride; int64_t bias_row_stride; /* fp16, strides in elements */
  const float* head_scale;
  const uint8_t* key_padding_mask;
  int32_t B, H, Tq, Tk, causal;
  float* lse; /* optional fp32 [B,H,Tq]: log2-domain log-sum-exp of every score row (saved for the backward) */
} sgf_attention_args;
int sgf_attention_bf16(const sgf_attention_args* args, void* stream);

/* ---------------------------------------------------------------------------------------
 * Patch-grid -> mask: bilinear (align_corners=False) upsample of the per-patch logits to
 * (h,w) fused with argmax over classes; full-resolution logits are never materialised.
 * Arithmetic order follows ATen's upsample_bilinear2d so that, on identical input logits,
 * the mask is bit-identical to F.interpolate(...).argmax(-1).
 *   logits fp32 [B, ld_tok (>= hp*wp), C] (token stride ld_c)  ->  mask int64 [B,h,w]
 *   optional per-class histograms (float [C] each, accumulated with atomics):
 *   area_pred, and with `target` (int64 [B,h,w], class ids, <0 or >=C = ignore): area_label,
 *   area_intersect  (seg_criterion.py:349-362).
 * `arith` selects the rounding sequence of the four-tap interpolation -- the argmax of two near-tied classes depends on
 * the last bit, and ATen rounds differently in each of its kernels:
 *   SGF_LERP_DEFAULT        what ATen executes on a GPU for the reference's call (mmseg.ops.resize -> F.interpolate on a
 *                           channels-last CUDA view): SGF_LERP_ATEN_CUDA below 16 classes, SGF_LERP_ATEN_CUDA_NHWC from 16 on
 *   SGF_LERP_ATEN_CUDA      src = fma(scale, d+0.5, -0.5); v = fma(h0, fma(w0,v00, w1*v01), h1*fma(w0,v10, w1*v11))
 *                           = nvcc's contraction of upsample_bilinear2d_out_frame (aten/native/cuda/UpSampleBilinear2d.cu)
 *   SGF_LERP_ATEN_CUDA_NHWC nvcc's contraction of upsample_bilinear2d_nhwc_out_frame: as above with top = fma(w1,v01, w0*v00)
 *                           (identified on a B200 against torch 2.11: 0 mismatches on adversarial near-tie logits,
 *                           profiles/r02_upsample_arith.txt)
 *   SGF_LERP_PLAIN          every product and sum rounded separately (ATen CPU, contiguous kernel)
 *   8 + v (v = 0..7)        the eight possible contractions (bit 0/1/2: operand order of the top / bottom / final sum)
 * Replaces seg_criterion.py:237-244 + :351 (and visualize_segmentation_web.ipynb cell 4).
 * ------------------------------------------------------------------------------------- */
#define SGF_LERP_DEFAULT (-1)
#define SGF_LERP_PLAIN 0
#define SGF_LERP_ATEN_CUDA 8
#define SGF_LERP_ATEN_CUDA_NHWC 9
typedef struct {
  const float* logits; int64_t batch_stride; int64_t tok_stride;
  int32_t B, C, hp, wp, h, w;
  int64_t* mask;
  const int64_t* target;
  float* area_intersect; float* area_pred; float* area_label;
  int32_t arith; /* SGF_LERP_* */
} sgf_segmask_args;
int sgf_upsample_argmax(const sgf_segmask_args* args, void* stream);

/* ---------------------------------------------------------------------------------------
 * Ragged EmbeddingBag(mode="mean") of the image-free branch: patch (b,i) is the mean of the
 * table rows of tokens[b, ends[b,i-1] : ends[b,i]] (ends = per-sample cumulative bag lengths,
 * ends[b,-1] = 0; pads sit at the tail of each token row, so no flatten/compact pass is needed).
 * One warp per bag, fp32 accumulation, out fp32 [B*P, D].
 * Replaces encoder_module.py:529-542 (mask-select, offset arithmetic, nn.EmbeddingBag) and
 * seg_criterion.py:386-393 (_lazy_initialization: B = 1, one bag per class name).
 * ------------------------------------------------------------------------------------- */
int sgf_embedding_bag_mean(const int64_t* tokens, int64_t ld_tokens, const int64_t* ends, int32_t B, int32_t P,
                           const void* table, int32_t table_dtype, int64_t ld_table, int32_t D, float* out,
                           void* stream);

/* ---------------------------------------------------------------------------------------
 * Pixel cross-entropy on the bilinearly upsampled logits without materialising them
 * (seg_criterion.py:246-267 / :340): for every pixel of the (h,w) grid the C interpolated logits
 * live in registers; loss_pix = logsumexp - logit[target] (label smoothing eps as in
 * F.cross_entropy); pixels whose target is outside [0,C) are ignored.
 *   out[0] += sum of pixel losses, out[1] += number of counted pixels (fp32 atomics; caller zeroes).
 *   logits fp32 [B, >= hp*wp, C]; target int64 [B,h,w] class ids (dictionary ids minus seg_id_offset).
 * ------------------------------------------------------------------------------------- */
typedef struct {
  const float* logits; int64_t batch_stride; int64_t tok_stride;
  int32_t B, C, hp, wp, h, w;
  const int64_t* target;
  float label_smoothing;
  float* out; /* [2] */
  float* lse_out; /* optional fp32 [B,h,w]: per-pixel logsumexp saved for sgf_upsample_ce_loss_bwd */
} sgf_segloss_args;
int sgf_upsample_ce_loss(const sgf_segloss_args* args, void* stream);

/* ---------------------------------------------------------------------------------------
 * Validation post-processing (seg_criterion.py:197-213, resnet_iters > 0): label propagation over the nearest
 * neighbours of the ResNet patch features.
 *   sgf_l2_normalize_rows : y[r] = x[r] / max(||x[r]||, 1e-12)        (F.normalize; bf16 rows)
 *   (cosine similarity = batched sgf_gemm_bf16 of the normalised features with themselves)
 *   sgf_row_topk          : the k (<= 8) largest columns of every fp32 row, largest first (torch.topk)
 *   sgf_label_propagation : prob = softmax(logits / temperature) over C, then `iters` rounds of
 *                           prob[b,p,:] = mean_k prob[b, nbr[b,p,k], :]; ping-pong buffers prob_a / prob_b
 *                           ([B,P,C] fp32): the result is in prob_a for even iters, prob_b for odd.
 * ------------------------------------------------------------------------------------- */
int sgf_l2_normalize_rows(const void* x, int64_t ldx, void* y, int64_t ldy, int32_t rows, int32_t D, void* stream);
int sgf_row_topk(const float* x, int64_t ldx, int32_t rows, int32_t n, int32_t k, int32_t* idx_out, void* stream);
int sgf_label_propagation(const float* logits, int64_t batch_stride, int64_t tok_stride, int32_t B, int32_t P, int32_t C,
                          float temperature, const int32_t* nbr, int32_t k, int32_t iters, float* prob_a, float* prob_b,
                          void* stream);

/* =======================================================================================
 * Training path (image-free branch, segofa.py:136-151 + seg_criterion.py:246-267): the
 * reference gets its backward from autograd over the ATen ops above; the entry points below are
 * the hand-written adjoint kernels.  Dense adjoints (dX = dY W, dW = dY^T X) reuse sgf_gemm_bf16
 * on operands transposed by sgf_transpose_cast.
 * ===================================================================================== */

/* Backward of sgf_upsample_ce_loss w.r.t. the low-resolution logits, gather form (no atomics,
 * deterministic): one CTA per (sample, patch), one thread per class; the thread walks the pixel
 * footprint of its patch, re-evaluates its class's interpolated logit, p = exp(v - lse[pixel]),
 * and accumulates weight * (p - (1-eps)[c == target] - eps/C).  The result is scaled by
 * grad_scale / count[0] (count = out[1] of the forward, read on the device: no host sync) and
 * written as bf16 dlogits[b, tok, c] (ld = ld_tok elements per token, batch stride in elements);
 * columns C..ld_tok-1 and token rows >= hp*wp (the eos slot) are zeroed by the kernel.
 * Adjoint of seg_criterion.py:237-244 (resize) + :263-267 (F.cross_entropy). */
typedef struct {
  const float* logits; int64_t batch_stride; int64_t tok_stride;
  int32_t B, C, hp, wp, h, w;
  const int64_t* target;
  const float* lse;      /* [B,h,w] from the forward */
  const float* count;    /* device pointer to the number of counted pixels */
  float label_smoothing; float grad_scale;
  void* dlogits; int64_t d_batch_stride; int64_t d_tok_stride; int32_t d_tokens; /* bf16 [B, d_tokens, d_tok_stride] */
} sgf_segloss_bwd_args;
int sgf_upsample_ce_loss_bwd(const sgf_segloss_bwd_args* args, void* stream);

/* Adjoint of sgf_row_layernorm.  Forward:  t = act(x[src]) (+ pre_add); u = g1 ? LN1(t) : t;
 * v = u (+ residual) = out1;  out2 = LN2(v).  Given dy2 = dL/d out2 (optional) and dv_in = dL/d out1
 * flowing in on the residual stream (optional):
 *     dv   = dv_in + LN2'(dy2; v)            -> d_res (fp32, optional; may alias dv_in)
 *     dt   = g1 ? LN1'(dv; t) : dv
 *     dx   = dt * act'(x)                     -> written (or accumulated) at the x row (src)
 *     dg1,db1,dg2,db2,d_pre_add (fp32 [D]) accumulated with atomics (per-CTA shared partials first).
 * LayerNorm statistics are recomputed from the saved rows (the row is staged in shared memory
 * anyway), so the forward saves no (mean, rstd).  v == NULL means v = t (no LN1, no residual).
 * Replaces autograd of LayerNorm / residual add / GELU at unify_transformer_layer.py:256-291,
 * 463-568 and the embedding LayerNorms of encoder_module.py:571-602, decoder_module.py:575-576. */
typedef struct {
  const void* x; int64_t ldx; int32_t x_dtype;
  const int64_t* gather_idx;
  int32_t x_act;
  const float* pre_add;
  const float* g1;
  const void* v; int64_t ldv; int32_t v_dtype;
  const float* g2;
  const void* dy2; int64_t ldy2; int32_t dy2_dtype;
  const float* dv_in; int64_t lddv;
  float* d_res; int64_t ldres;
  void* dx; int64_t lddx; int32_t dx_dtype; int32_t dx_accumulate;
  float* dg1; float* db1; float* dg2; float* db2; float* d_pre_add;
  int32_t rows, D;
  int32_t seg_len, seg_stride, seg_off;
  float* dx_colsum; /* optional fp32 [D] += column sums of dx (= bias gradient of the linear that produced x) */
  /* same dropout / DropPath description as the forward call: du = dv * mask (d_res = dv is not masked) */
  float drop_p; float droppath_p; uint32_t drop_seed; uint32_t drop_site; int32_t rows_per_sample;
  const int32_t* drop_step;
} sgf_rowln_bwd_args;
int sgf_row_layernorm_bwd(const sgf_rowln_bwd_args* args, void* stream);

/* in [M,N] (fp32 or bf16, row stride ld_in) -> out_t bf16 [N, ld_t >= M] (transposed; columns
 * M..ld_t-1 zero-filled up to the next multiple of 8), optional out_c bf16 [M, ld_c] (plain cast
 * copy), optional colsum fp32 [N] += sum_m in[m,n] (bias gradients).  Produces the operands of
 * the dense adjoints: W^T for dX = dY W, and dY^T / X^T for dW = dY^T X. */
int sgf_transpose_cast(const void* in, int32_t in_dtype, int64_t ld_in, int32_t M, int32_t N, void* out_t,
                       int64_t ld_t, void* out_c, int64_t ld_c, float* colsum, void* stream);

/* Backward of sgf_attention_bf16 (same operand addressing).  Three kernels:
 *   delta[b,h,i] = sum_d dout[b,i,h,d] * out[b,i,h,d]   (+ d_head_scale[h] += delta / head_scale[h])
 *   dQ  : CTA per (128 queries, head, batch): S = Q K^T and dP = dO V^T on tcgen05 into TMEM,
 *         P = exp2(S + bias - lse), dS = P (head_scale dP - delta) -> bf16 smem, dQ += dS K (TMEM)
 *   dKdV: CTA per (128 keys, head, batch): S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q
 * dq is multiplied by dq_scale (the q pre-scaling of the QKV epilogue).  The gradient w.r.t. the
 * additive position bias is accumulated into `dbias` when given (vector fp32 atomics from the dQ kernel).
 * The bias is the forward's FP16 tensor: the dQ kernel streams it as double-buffered 128B-swizzled TMA tiles.
 * delta: fp32 scratch [B,H,Tq]. */
typedef struct {
  const void* q; int64_t q_row_stride; int64_t q_batch_stride;
  const void* k; int64_t k_row_stride; int64_t k_batch_stride;
  const void* v; int64_t v_row_stride; int64_t v_batch_stride;
  const void* out; int64_t o_row_stride; int64_t o_batch_stride;
  const void* dout; int64_t do_row_stride; int64_t do_batch_stride;
  void* dq; int64_t dq_row_stride; int64_t dq_batch_stride;
  void* dk; int64_t dk_row_stride; int64_t dk_batch_stride;
  void* dv; int64_t dv_row_stride; int64_t dv_batch_stride;
  const void* bias; int64_t bias_head_stride; int64_t bias_row_stride; /* FP16 (the tensor the forward streamed), strides in elements, multiples of 8 */
  const float* head_scale; float* d_head_scale;
  const uint8_t* key_padding_mask;
  const float* lse; float* delta;
  float dq_scale;
  int32_t B, H, Tq, Tk, causal;
  float* dbias; /* optional fp32 [H,Tq,bias_row_stride] (bias strides; row stride a multiple of 64): += dS summed over
                   the batch, i.e. the gradient w.r.t. the additive position bias */
  const void* bias_t; int64_t bias_t_head_stride; int64_t bias_t_row_stride; /* FP16 key-major copy of bias, required with bias
                   [H,Tk,>=roundup(Tq,64)] (sgf_transpose16_batched): the dK/dV kernel, whose threads own keys, then reads
                   128 contiguous bytes per tile instead of 64 strided elements */
} sgf_attention_bwd_args;
int sgf_attention_bwd_bf16(const sgf_attention_bwd_args* args, void* stream);

/* out[h][c][r] = in[h][r][c] for 16-bit elements, H matrices of R x C (C a multiple of 8); the rows of `out` are padded
 * to a multiple of 64 elements (zero-filled beyond R).  Strides in elements, multiples of 8; 16-byte aligned pointers. */
int sgf_transpose16_batched(const void* in, int64_t in_head_stride, int64_t in_row_stride, int32_t H, int32_t R, int32_t C,
                            void* out, int64_t out_head_stride, int64_t out_row_stride, void* stream);

/* Adjoint of sgf_build_attn_bias for one layer.  dbias[h,i,j] is the batch-summed dS of that layer's attention
 * (written by sgf_attention_bwd_bf16).  For every block b the table gradient is a gather over a static CSR list of
 * the block's positions grouped by bucket (built once per shape by the caller from the same bucket/ids tensors the
 * forward uses): dtable_b[r,h] += sum_{t in [offsets_b[r], offsets_b[r+1])} dbias[h].flat[order_b[t]], with
 * order_b[t] = i*row_stride + j -- one warp per (r,h), no atomics, deterministic.  Then dabs_acc += dbias (the
 * gradient of the layer-independent abs term; optional) and dbias is cleared for the next layer.
 * Adjoint of encoder_module.py:313-331,790-809 / decoder_module.py:327-333,601-627 (identity-interpolation case). */
typedef struct {
  float* dbias; float* dabs_acc; int64_t head_stride; int64_t row_stride;
  int32_t H, Tq;
  int32_t num_blocks; const int32_t* order[2]; const int32_t* offsets[2]; float* dtable[2]; int32_t num_rel[2];
} sgf_bias_bwd_args;
int sgf_attn_bias_bwd(const sgf_bias_bwd_args* args, void* stream);

/* Device-side generation of the image-free training sample (data/mm_data/segmentation_dataset.py:303-329,
 * artificial_image_type 'rand_k-lo-hi'): per sample a random sh x sw label grid (sh, sw in
 * [lo, hi), labels in [0, C)) nearest-resized to the hp x hp patch grid -> bag_tokens int64 [B, bag_ld] (the
 * class-name tokens of every patch, back to back, tail = pad_id; bag_ld >= hp*hp*max name length) and bag_ends int64
 * [B, hp*hp] (cumulative bag lengths) = aux_input["patch_images"] / ["patch_masks"]; and to the S x S pixel grid ->
 * target int64 [B, S*S+1] (seg_id_offset + label, then eos_id) = sample["text2seg_target"].  name_tokens int64
 * [C, name_ld] / name_lens int32 [C] are the BPE ids of the class names.  Counter-based RNG keyed by (seed, step[0]).
 * grid_out (optional int32 [B, 2 + 32*32]) receives (sh, sw, labels) for inspection. */
typedef struct {
  const int64_t* name_tokens; int32_t name_ld; const int32_t* name_lens;
  int32_t C, B, hp, S, lo, hi;
  uint32_t seed; const int32_t* step;
  int64_t seg_id_offset, eos_id, pad_id;
  int64_t* bag_tokens; int64_t bag_ld; int64_t* bag_ends; int64_t* target; int32_t* grid_out;
} sgf_artsample_args;
int sgf_artificial_sample(const sgf_artsample_args* args, void* stream);

/* Real-image side of the input pipeline (data/mm_data/segmentation_dataset.py:210-301): the decoded
 * uint8 image / label PNG crosses PCIe as bytes and the per-sample transforms run on the device, bit-identical to the
 * host libraries the reference calls.
 * sgf_image_prep_u8 replaces mmseg Resize (mmcv.imrescale -> cv2.resize INTER_LINEAR, 8-bit fixed point) ->
 * RandomCrop window -> RandomFlip (horizontal) -> transforms.ToTensor (/255) -> transforms.Normalize (:155-156, :236-240,
 * :253-257): src = uint8 [src_h, src_w, 3] (row stride in bytes; channel order as it should come out), resized to
 * rs_h x rs_w, window (crop_y, crop_x, out_h, out_w) of it, mirrored when flip != 0, written as fp32 [3, out_h, out_w].
 * Validation (MultiScaleFlipAug, keep-ratio resize): crop = (0, 0, rs_h, rs_w), flip = 0.  The random choices
 * (scale, window, flip) are the caller's; PhotoMetricDistortion is not reproduced. */
typedef struct {
  const uint8_t* src; int64_t src_row_stride; int32_t src_h, src_w;
  int32_t rs_h, rs_w;
  int32_t crop_y, crop_x, out_h, out_w;
  int32_t flip;
  float mean[3]; float std[3];
  float* dst; int64_t dst_channel_stride; int64_t dst_row_stride;
} sgf_image_prep_args;
int sgf_image_prep_u8(const sgf_image_prep_args* args, void* stream);

/* Label side of the same sample (:224-227, :241-262, :264-265): src = the raw uint8 label PNG [src_h, src_w] (0 =
 * unlabelled, k = class k-1) -> remap (0 and 255 -> num_seg, k -> k-1) -> cv2 INTER_NEAREST resize to rs_h x rs_w ->
 * the same window / flip -> target int64 [out_h*out_w + 1] = seg_id_offset + class, then eos_id (sample["target"]);
 * prev_output_tokens int64 [grid_h*grid_w + 1] = bos_id, then the codes of torchvision's NEAREST resize of that map to
 * the patch grid; downsampled_target (optional) int64 [grid_h*grid_w + 1] = those codes, then eos_id; ori_classes
 * (optional) int64 [src_h, src_w] = the remapped map at the original resolution (sample["ori_semantic_seg"]). */
typedef struct {
  const uint8_t* src; int64_t src_row_stride; int32_t src_h, src_w;
  int32_t num_seg;
  int32_t rs_h, rs_w;
  int32_t crop_y, crop_x, out_h, out_w;
  int32_t flip;
  int32_t grid_h, grid_w;
  int64_t seg_id_offset, bos_id, eos_id;
  int64_t* target; int64_t* prev_output_tokens; int64_t* downsampled_target; int64_t* ori_classes;
} sgf_segmap_prep_args;
int sgf_segmap_prep_u8(const sgf_segmap_prep_args* args, void* stream);

/* Multi-tensor-free fused Adam(W) step over one flat fp32 master buffer (cf/optim/adam.py, fp32
 * master weights of cf/optim/fp16_optimizer.py:108-222): p -= lr*(m_hat/(sqrt(v_hat)+eps) + wd*p)
 * with grads scaled by grad_scale[0] (device scalar: 1/sample_size * clip coefficient).  The update
 * number is `step`, or step_dev[0] when step_dev != NULL (device-resident counter: the launch can then be
 * replayed from a CUDA graph). */
int sgf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int32_t step, const int32_t* step_dev,
                  const float* grad_scale, void* stream);
/* out[0] += sum of squares of x[0..n) (gradient-norm for clip_grad_norm, trainer.py:886) */
int sgf_sumsq(const float* x, int64_t n, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGOFA_B200_H_ */
