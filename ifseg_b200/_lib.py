"""ctypes binding of libsegofa_b200.so (the C ABI declared in include/segofa_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  SGF_ERR_OOM is reported with "out of memory" in the message so that
the reference trainer's OOM recovery (trainer.py:807-822) keeps working.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsegofa_b200.so")

SGF_BF16, SGF_F32 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2

_vp, _i64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", _vp), ("lda", _i64), ("a_batch_stride", _i64),
        ("b", _vp), ("ldb", _i64), ("b_batch_stride", _i64),
        ("c", _vp), ("ldc", _i64), ("c_batch_stride", _i64), ("c_dtype", _i32),
        ("M", _i32), ("N", _i32), ("K", _i32), ("batch", _i32),
        ("col_scale", _vp), ("col_bias", _vp),
        ("residual", _vp), ("ldr", _i64), ("r_batch_stride", _i64), ("r_dtype", _i32),
        ("act", _i32), ("alpha", _f32), ("alpha_cols", _i32),
        ("rowstats_out", _vp), ("rownorm_stats", _vp), ("rownorm_u", _vp), ("rownorm_dim", _i32),
    ]


class Conv3x3Args(C.Structure):
    _fields_ = [
        ("x", _vp), ("w", _vp), ("y", _vp),
        ("n", _i32), ("h", _i32), ("w_", _i32), ("cin", _i32), ("cout", _i32),
        ("col_scale", _vp), ("col_bias", _vp), ("act", _i32),
    ]


class Conv2dArgs(C.Structure):
    _fields_ = [
        ("x", _vp), ("n", _i32), ("h", _i32), ("w_", _i32), ("cin", _i32),
        ("x_w_stride", _i64), ("x_h_stride", _i64), ("x_n_stride", _i64),
        ("w", _vp), ("y", _vp), ("ho", _i32), ("wo", _i32), ("cout", _i32),
        ("kh", _i32), ("kw", _i32), ("stride_h", _i32), ("stride_w", _i32), ("pad_h", _i32), ("pad_w", _i32),
        ("col_scale", _vp), ("col_bias", _vp), ("act", _i32),
    ]


class RowLnArgs(C.Structure):
    _fields_ = [
        ("x", _vp), ("ldx", _i64), ("x_dtype", _i32),
        ("gather_idx", _vp), ("pre_add", _vp),
        ("g1", _vp), ("b1", _vp),
        ("residual", _vp), ("ldr", _i64), ("r_dtype", _i32),
        ("out1", _vp), ("ld1", _i64), ("out1_dtype", _i32),
        ("g2", _vp), ("b2", _vp),
        ("out2", _vp), ("ld2", _i64),
        ("zero_row", _vp),
        ("rows", _i32), ("D", _i32),
        ("seg_len", _i32), ("seg_stride", _i32), ("seg_off", _i32),
        ("clear_rowstats", _vp),
        ("x_act", _i32),
        ("drop_p", _f32), ("droppath_p", _f32), ("drop_seed", C.c_uint32), ("drop_site", C.c_uint32),
        ("rows_per_sample", _i32), ("drop_step", _vp),
    ]


class RelBlock(C.Structure):
    _fields_ = [("bucket", _vp), ("bucket_ld", _i64), ("ids", _vp), ("table", _vp), ("lo", _i32), ("hi", _i32)]


class BiasArgs(C.Structure):
    _fields_ = [
        ("out", _vp), ("abs", _vp), ("head_stride", _i64), ("row_stride", _i64), ("dense_add", _vp),
        ("H", _i32), ("Tq", _i32), ("Tk", _i32), ("num_blocks", _i32), ("blocks", RelBlock * 2),
        ("out_f16", _vp),
    ]


class AttentionArgs(C.Structure):
    _fields_ = [
        ("q", _vp), ("q_row_stride", _i64), ("q_batch_stride", _i64),
        ("k", _vp), ("k_row_stride", _i64), ("k_batch_stride", _i64),
        ("v", _vp), ("v_row_stride", _i64), ("v_batch_stride", _i64),
        ("out", _vp), ("o_row_stride", _i64), ("o_batch_stride", _i64),
        ("bias", _vp), ("bias_head_stride", _i64), ("bias_row_stride", _i64),
        ("head_scale", _vp), ("key_padding_mask", _vp),
        ("B", _i32), ("H", _i32), ("Tq", _i32), ("Tk", _i32), ("causal", _i32),
        ("lse", _vp),
    ]


class SegmaskArgs(C.Structure):
    _fields_ = [
        ("logits", _vp), ("batch_stride", _i64), ("tok_stride", _i64),
        ("B", _i32), ("C", _i32), ("hp", _i32), ("wp", _i32), ("h", _i32), ("w", _i32),
        ("mask", _vp), ("target", _vp),
        ("area_intersect", _vp), ("area_pred", _vp), ("area_label", _vp),
        ("arith", _i32),
    ]


class SeglossArgs(C.Structure):
    _fields_ = [
        ("logits", _vp), ("batch_stride", _i64), ("tok_stride", _i64),
        ("B", _i32), ("C", _i32), ("hp", _i32), ("wp", _i32), ("h", _i32), ("w", _i32),
        ("target", _vp), ("label_smoothing", _f32), ("out", _vp), ("lse_out", _vp),
    ]


class SeglossBwdArgs(C.Structure):
    _fields_ = [
        ("logits", _vp), ("batch_stride", _i64), ("tok_stride", _i64),
        ("B", _i32), ("C", _i32), ("hp", _i32), ("wp", _i32), ("h", _i32), ("w", _i32),
        ("target", _vp), ("lse", _vp), ("count", _vp),
        ("label_smoothing", _f32), ("grad_scale", _f32),
        ("dlogits", _vp), ("d_batch_stride", _i64), ("d_tok_stride", _i64), ("d_tokens", _i32),
    ]


class RowLnBwdArgs(C.Structure):
    _fields_ = [
        ("x", _vp), ("ldx", _i64), ("x_dtype", _i32),
        ("gather_idx", _vp),
        ("x_act", _i32),
        ("pre_add", _vp),
        ("g1", _vp),
        ("v", _vp), ("ldv", _i64), ("v_dtype", _i32),
        ("g2", _vp),
        ("dy2", _vp), ("ldy2", _i64), ("dy2_dtype", _i32),
        ("dv_in", _vp), ("lddv", _i64),
        ("d_res", _vp), ("ldres", _i64),
        ("dx", _vp), ("lddx", _i64), ("dx_dtype", _i32), ("dx_accumulate", _i32),
        ("dg1", _vp), ("db1", _vp), ("dg2", _vp), ("db2", _vp), ("d_pre_add", _vp),
        ("rows", _i32), ("D", _i32),
        ("seg_len", _i32), ("seg_stride", _i32), ("seg_off", _i32),
        ("dx_colsum", _vp),
        ("drop_p", _f32), ("droppath_p", _f32), ("drop_seed", C.c_uint32), ("drop_site", C.c_uint32),
        ("rows_per_sample", _i32), ("drop_step", _vp),
    ]


class AttentionBwdArgs(C.Structure):
    _fields_ = [
        ("q", _vp), ("q_row_stride", _i64), ("q_batch_stride", _i64),
        ("k", _vp), ("k_row_stride", _i64), ("k_batch_stride", _i64),
        ("v", _vp), ("v_row_stride", _i64), ("v_batch_stride", _i64),
        ("out", _vp), ("o_row_stride", _i64), ("o_batch_stride", _i64),
        ("dout", _vp), ("do_row_stride", _i64), ("do_batch_stride", _i64),
        ("dq", _vp), ("dq_row_stride", _i64), ("dq_batch_stride", _i64),
        ("dk", _vp), ("dk_row_stride", _i64), ("dk_batch_stride", _i64),
        ("dv", _vp), ("dv_row_stride", _i64), ("dv_batch_stride", _i64),
        ("bias", _vp), ("bias_head_stride", _i64), ("bias_row_stride", _i64),
        ("head_scale", _vp), ("d_head_scale", _vp),
        ("key_padding_mask", _vp),
        ("lse", _vp), ("delta", _vp),
        ("dq_scale", _f32),
        ("B", _i32), ("H", _i32), ("Tq", _i32), ("Tk", _i32), ("causal", _i32),
        ("dbias", _vp),
        ("bias_t", _vp), ("bias_t_head_stride", _i64), ("bias_t_row_stride", _i64),
    ]


class BiasBwdArgs(C.Structure):
    _fields_ = [
        ("dbias", _vp), ("dabs_acc", _vp), ("head_stride", _i64), ("row_stride", _i64),
        ("H", _i32), ("Tq", _i32),
        ("num_blocks", _i32), ("order", _vp * 2), ("offsets", _vp * 2), ("dtable", _vp * 2), ("num_rel", _i32 * 2),
    ]


class ArtSampleArgs(C.Structure):
    _fields_ = [
        ("name_tokens", _vp), ("name_ld", _i32), ("name_lens", _vp),
        ("C", _i32), ("B", _i32), ("hp", _i32), ("S", _i32), ("lo", _i32), ("hi", _i32),
        ("seed", C.c_uint32), ("step", _vp),
        ("seg_id_offset", _i64), ("eos_id", _i64), ("pad_id", _i64),
        ("bag_tokens", _vp), ("bag_ld", _i64), ("bag_ends", _vp), ("target", _vp), ("grid_out", _vp),
    ]


class ImagePrepArgs(C.Structure):
    _fields_ = [
        ("src", _vp), ("src_row_stride", _i64), ("src_h", _i32), ("src_w", _i32),
        ("rs_h", _i32), ("rs_w", _i32),
        ("crop_y", _i32), ("crop_x", _i32), ("out_h", _i32), ("out_w", _i32),
        ("flip", _i32),
        ("mean", _f32 * 3), ("std", _f32 * 3),
        ("dst", _vp), ("dst_channel_stride", _i64), ("dst_row_stride", _i64),
    ]


class SegmapPrepArgs(C.Structure):
    _fields_ = [
        ("src", _vp), ("src_row_stride", _i64), ("src_h", _i32), ("src_w", _i32),
        ("num_seg", _i32),
        ("rs_h", _i32), ("rs_w", _i32),
        ("crop_y", _i32), ("crop_x", _i32), ("out_h", _i32), ("out_w", _i32),
        ("flip", _i32),
        ("grid_h", _i32), ("grid_w", _i32),
        ("seg_id_offset", _i64), ("bos_id", _i64), ("eos_id", _i64),
        ("target", _vp), ("prev_output_tokens", _vp), ("downsampled_target", _vp), ("ori_classes", _vp),
    ]


# every symbol include/segofa_b200.h declares: (name, restype, argtypes)
EXPORTS = [
    ("sgf_last_error", C.c_char_p, []),
    ("sgf_abi_version", C.c_int, []),
    ("sgf_launch_count", _i64, []),
    ("sgf_reset_launch_count", None, []),
    ("sgf_debug_set_gemm_trace", None, [_vp]),
    ("sgf_debug_set_attention_trace", None, [_vp]),
    ("sgf_gemm_bf16", C.c_int, [C.POINTER(GemmArgs), _vp]),
    ("sgf_gemm_bf16_ex", C.c_int, [C.POINTER(GemmArgs), _i32, _i32, _i32, _vp]),
    ("sgf_conv3x3_s1_nhwc", C.c_int, [C.POINTER(Conv3x3Args), _vp]),
    ("sgf_conv2d_nhwc", C.c_int, [C.POINTER(Conv2dArgs), _vp]),
    ("sgf_nchw_f32_to_nhwc_bf16", C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    ("sgf_nchw_f32_to_nhwc8_padded", C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    ("sgf_maxpool3x3s2_nhwc", C.c_int, [_vp, _vp] + [_i32] * 6 + [_vp]),
    ("sgf_row_layernorm", C.c_int, [C.POINTER(RowLnArgs), _vp]),
    ("sgf_build_attn_bias", C.c_int, [C.POINTER(BiasArgs), _vp]),
    ("sgf_attention_bf16", C.c_int, [C.POINTER(AttentionArgs), _vp]),
    ("sgf_upsample_argmax", C.c_int, [C.POINTER(SegmaskArgs), _vp]),
    ("sgf_embedding_bag_mean", C.c_int, [_vp, _i64, _vp, _i32, _i32, _vp, _i32, _i64, _i32, _vp, _vp]),
    ("sgf_upsample_ce_loss", C.c_int, [C.POINTER(SeglossArgs), _vp]),
    ("sgf_l2_normalize_rows", C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _vp]),
    ("sgf_row_topk", C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    ("sgf_label_propagation", C.c_int, [_vp, _i64, _i64, _i32, _i32, _i32, _f32, _vp, _i32, _i32, _vp, _vp, _vp]),
    ("sgf_upsample_ce_loss_bwd", C.c_int, [C.POINTER(SeglossBwdArgs), _vp]),
    ("sgf_row_layernorm_bwd", C.c_int, [C.POINTER(RowLnBwdArgs), _vp]),
    ("sgf_transpose_cast", C.c_int, [_vp, _i32, _i64, _i32, _i32, _vp, _i64, _vp, _i64, _vp, _vp]),
    ("sgf_attention_bwd_bf16", C.c_int, [C.POINTER(AttentionBwdArgs), _vp]),
    ("sgf_transpose16_batched", C.c_int, [_vp, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _i64, _vp]),
    ("sgf_attn_bias_bwd", C.c_int, [C.POINTER(BiasBwdArgs), _vp]),
    ("sgf_artificial_sample", C.c_int, [C.POINTER(ArtSampleArgs), _vp]),
    ("sgf_image_prep_u8", C.c_int, [C.POINTER(ImagePrepArgs), _vp]),
    ("sgf_segmap_prep_u8", C.c_int, [C.POINTER(SegmapPrepArgs), _vp]),
    ("sgf_adam_step", C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _vp, _vp, _vp]),
    ("sgf_sumsq", C.c_int, [_vp, _i64, _vp, _vp]),
]

_lib = None


def load():
    """Loads the library (once).  Raises RuntimeError with build instructions if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the segofa_b200 CUDA extension is not built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (or ifseg_b200/csrc/build.py). "
            "There is no CPU/PyTorch fallback for the hot path."
        )
    lib = C.CDLL(LIB_PATH)
    for name, res, args in EXPORTS:
        fn = getattr(lib, name)  # AttributeError -> missing export
        fn.restype = res
        fn.argtypes = args
    if lib.sgf_abi_version() != 3:
        raise RuntimeError("libsegofa_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what):
    if rc == 0:
        return
    msg = load().sgf_last_error().decode("utf-8", "replace")
    if rc == 3:
        raise RuntimeError(f"CUDA out of memory in {what}: {msg}")
    if rc == 1:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what} failed (code {rc}): {msg}")
