"""ctypes binding of include/conmamba_b200.h.

The structures mirror the header field by field.  ``lib()`` loads ``lib/libconmamba_b200.so`` (built in-tree by
``python -m mamba_asr_b200.build`` / ``__graft_entry__.build()``) and raises if it is missing: there is no
CPU or Triton fallback behind this package.
"""
import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CM_LIB_PATH") or os.path.join(_PKG, "lib", "libconmamba_b200.so")   # env: tuning experiments

CM_F32, CM_BF16, CM_F16 = 0, 1, 2
CM_ERR_BAD_ARG, CM_ERR_UNSUPPORTED = -1, -2
CM_FLAG_DELTA_SOFTPLUS = 1
CM_FLAG_SILU = 1
CM_SCAN_CKPT_STEPS = 8
CM_ABI_VERSION = 17
CM_LN_ACT_LEAKY_RELU, CM_LN_ACT_GELU = 0, 1
CM_LN_OUT_GELU = 1

EXPORTS = (
    "cm_version", "cm_scan_num_ckpt", "cm_scan_slab_channels", "cm_scan_pick_lanes", "cm_scan_pick_lanes_bwd", "cm_scan_bwd_slab_channels", "cm_scan_fwd",
    "cm_scan_bwd",
    "cm_reduce_dbc", "cm_reduce_rows", "cm_conv_fwd", "cm_conv_num_part", "cm_conv_bwd", "cm_conv_update",
    "cm_fbank_logmel", "cm_fbank_floor", "cm_abi_sizeof", "cm_reduce_multi", "cm_layernorm_num_part", "cm_layernorm_num_part2", "cm_layernorm_fwd",
    "cm_layernorm_bwd", "cm_scan_fwd_workspace_bytes", "cm_dwconv_num_part", "cm_dwconv_fwd", "cm_dwconv_bwd_weight", "cm_colsum_num_part", "cm_colsum",
    "cm_ssm_step", "cm_add_ln_fwd", "cm_add_ln_bwd", "cm_gelu_dropout_fwd", "cm_gelu_dropout_bwd", "cm_tsmm_num_part", "cm_tsmm",
    "cm_ln_act_num_part", "cm_ln_act_fwd", "cm_ln_act_bwd", "cm_optim_num_part", "cm_sumsq_partial", "cm_adamw_step",
    "cm_fbank_wav_supported", "cm_fbank_wav_logmel", "cm_ctc_workspace_floats", "cm_ctc_loss",
    "cm_stem_supported", "cm_stem_num_part", "cm_stem_fwd", "cm_stem_bwd",
    "cm_reduce_batch", "cm_glu_fwd", "cm_glu_bwd",
    "cm_add_ln_dbsum_supported", "cm_add_ln_num_part", "cm_act_colsum_supported", "cm_act_num_part", "cm_gelu_dropout_fwd_v2", "cm_gelu_dropout_bwd_v2",
)
CM_REDUCE_MAX_JOBS = 8
CM_REDUCE_BATCH_MAX = 64

_DTYPES = {torch.float32: CM_F32, torch.bfloat16: CM_BF16, torch.float16: CM_F16}


class Tensor3(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sb", C.c_int64), ("sd", C.c_int64), ("sl", C.c_int64)]


class ScanDir(C.Structure):
    _fields_ = [
        ("reverse", C.c_int32), ("bc_const", C.c_int32),
        ("u", Tensor3), ("delta", Tensor3), ("Bm", Tensor3), ("Cm", Tensor3),
        ("A", C.c_void_p), ("A_sd", C.c_int64), ("A_sn", C.c_int64),
        ("Dskip", C.c_void_p), ("delta_bias", C.c_void_p),
        ("ckpt", C.c_void_p), ("ckpt_sb", C.c_int64), ("ckpt_sd", C.c_int64),
        ("last_state", C.c_void_p), ("ls_sb", C.c_int64), ("ls_sd", C.c_int64), ("ls_sn", C.c_int64),
    ]


class ScanFwdArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("dim", C.c_int32), ("seqlen", C.c_int32), ("dstate", C.c_int32),
        ("ndir", C.c_int32), ("dtype", C.c_int32), ("flags", C.c_uint32), ("out_scale", C.c_float),
        ("lanes_per_channel", C.c_int32), ("reserved", C.c_int32),
        ("dir", ScanDir * 2),
        ("z", Tensor3), ("out", Tensor3), ("out_pre", Tensor3),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class ScanBwdDir(C.Structure):
    _fields_ = [
        ("inp", ScanDir), ("du", Tensor3), ("ddelta", Tensor3),
        ("dBC_part", C.c_void_p), ("dA_part", C.c_void_p), ("dD_part", C.c_void_p), ("dbias_part", C.c_void_p),
    ]


class ScanBwdArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("dim", C.c_int32), ("seqlen", C.c_int32), ("dstate", C.c_int32),
        ("ndir", C.c_int32), ("dtype", C.c_int32), ("flags", C.c_uint32), ("out_scale", C.c_float),
        ("lanes_per_channel", C.c_int32), ("reserved", C.c_int32),
        ("dir", ScanBwdDir * 2),
        ("z", Tensor3), ("out_pre", Tensor3), ("dout", Tensor3), ("dz", Tensor3),
    ]


class ConvDir(C.Structure):
    _fields_ = [
        ("anticausal", C.c_int32), ("reserved", C.c_int32),
        ("weight", C.c_void_p), ("bias", C.c_void_p), ("out", Tensor3),
        ("dweight_part", C.c_void_p), ("dbias_part", C.c_void_p),
    ]


class ConvArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("dim", C.c_int32), ("seqlen", C.c_int32), ("width", C.c_int32),
        ("ndir", C.c_int32), ("dtype", C.c_int32), ("flags", C.c_uint32), ("reserved", C.c_int32),
        ("x", Tensor3), ("dx", Tensor3),
        ("dir", ConvDir * 2),
    ]


class FbankArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("frames", C.c_int32), ("nbins", C.c_int32), ("nmels", C.c_int32),
        ("stft", C.c_void_p), ("s_b", C.c_int64), ("s_f", C.c_int64), ("s_t", C.c_int64),
        ("fbank", C.c_void_p), ("out", C.c_void_p), ("utt_max", C.c_void_p),
        ("amin", C.c_float), ("multiplier", C.c_float), ("db_offset", C.c_float), ("top_db", C.c_float),
    ]


class FbankWavArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("n_samples", C.c_int32), ("frames", C.c_int32), ("nmels", C.c_int32),
        ("n_fft", C.c_int32), ("hop", C.c_int32),
        ("wav", C.c_void_p), ("wav_sb", C.c_int64),
        ("window", C.c_void_p), ("fbank", C.c_void_p), ("band", C.c_void_p), ("out", C.c_void_p), ("utt_max", C.c_void_p),
        ("amin", C.c_float), ("multiplier", C.c_float), ("db_offset", C.c_float), ("top_db", C.c_float),
    ]


class CtcArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("max_time", C.c_int32), ("classes", C.c_int32), ("max_target", C.c_int32),
        ("blank", C.c_int32), ("ws_states", C.c_int32),
        ("log_probs", C.c_void_p), ("lp_sb", C.c_int64), ("lp_st", C.c_int64),
        ("targets", C.c_void_p), ("tg_sb", C.c_int64),
        ("input_lengths", C.c_void_p), ("target_lengths", C.c_void_p),
        ("nll", C.c_void_p), ("grad", C.c_void_p), ("g_sb", C.c_int64), ("g_st", C.c_int64),
        ("workspace", C.c_void_p),
    ]


class StemArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("frames", C.c_int32), ("feats", C.c_int32), ("channels", C.c_int32),
        ("in_dtype", C.c_int32), ("out_dtype", C.c_int32), ("eps", C.c_float), ("slope", C.c_float),
        ("inp", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("y", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p), ("dy", C.c_void_p),
        ("dgamma_part", C.c_void_p), ("dbeta_part", C.c_void_p), ("dweight_part", C.c_void_p), ("dbias_part", C.c_void_p),
    ]


class ActArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("y", C.c_void_p), ("dy", C.c_void_p), ("dx", C.c_void_p), ("mask", C.c_void_p),
        ("seed", C.c_void_p), ("key", C.c_void_p), ("call_id", C.c_uint32), ("dtype", C.c_int32), ("p_drop", C.c_float),
        ("cols", C.c_int32), ("n", C.c_int64), ("colsum_part", C.c_void_p), ("keep_bits", C.c_void_p),
    ]


class ReduceJob(C.Structure):
    _fields_ = [("part", C.c_void_p), ("out", C.c_void_p), ("rows", C.c_int64), ("cols", C.c_int64)]


class ReduceJob2(C.Structure):
    _fields_ = [("part", C.c_void_p), ("out", C.c_void_p), ("rows", C.c_int64), ("cols", C.c_int64), ("stride", C.c_int64)]


class LayerNormArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int64), ("cols", C.c_int32), ("x_dtype", C.c_int32), ("y_dtype", C.c_int32), ("eps", C.c_float),
        ("x", C.c_void_p), ("x_stride", C.c_int64), ("y", C.c_void_p), ("y_stride", C.c_int64),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p),
        ("dy", C.c_void_p), ("dy_stride", C.c_int64), ("dx", C.c_void_p), ("dx_stride", C.c_int64),
        ("dgamma_part", C.c_void_p), ("dbeta_part", C.c_void_p), ("act", C.c_int32), ("n_part", C.c_int32),
    ]


class DwConvArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("dim", C.c_int32), ("seqlen", C.c_int32), ("ksize", C.c_int32),
        ("pad_left", C.c_int32), ("dtype", C.c_int32), ("flip", C.c_int32), ("reserved", C.c_int32),
        ("x", Tensor3), ("y", Tensor3), ("weight", C.c_void_p), ("bias", C.c_void_p), ("dy", Tensor3),
        ("dweight_part", C.c_void_p), ("dbias_part", C.c_void_p),
    ]


_lib = None


class SsmStepArgs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("dim", C.c_int32), ("dstate", C.c_int32), ("dtype", C.c_int32),
                ("state_dtype", C.c_int32), ("flags", C.c_uint32),
                ("state", C.c_void_p), ("x", C.c_void_p), ("dt", C.c_void_p), ("z", C.c_void_p), ("Bm", C.c_void_p),
                ("Cm", C.c_void_p), ("out", C.c_void_p),
                ("x_sb", C.c_int64), ("dt_sb", C.c_int64), ("z_sb", C.c_int64), ("b_sb", C.c_int64), ("c_sb", C.c_int64),
                ("out_sb", C.c_int64),
                ("A", C.c_void_p), ("Dskip", C.c_void_p), ("dt_bias", C.c_void_p)]


class AddLnArgs(C.Structure):
    _fields_ = [("rows", C.c_int64), ("cols", C.c_int32), ("a_dtype", C.c_int32), ("b_dtype", C.c_int32),
                ("y_dtype", C.c_int32), ("eps", C.c_float), ("alpha", C.c_float), ("p_drop", C.c_float),
                ("call_id", C.c_uint32), ("reserved", C.c_int32), ("seed", C.c_void_p),
                ("a", C.c_void_p), ("a_stride", C.c_int64), ("b", C.c_void_p), ("b_stride", C.c_int64),
                ("s", C.c_void_p), ("s_stride", C.c_int64), ("y", C.c_void_p), ("y_stride", C.c_int64),
                ("mask", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("mean", C.c_void_p),
                ("rstd", C.c_void_p),
                ("dy", C.c_void_p), ("dy_stride", C.c_int64), ("ds", C.c_void_p), ("ds_stride", C.c_int64),
                ("da", C.c_void_p), ("da_stride", C.c_int64), ("db", C.c_void_p), ("db_stride", C.c_int64),
                ("dgamma_part", C.c_void_p), ("dbeta_part", C.c_void_p), ("key", C.c_void_p), ("dbsum_part", C.c_void_p)]


class LnActArgs(C.Structure):
    _fields_ = [("rows", C.c_int64), ("cols", C.c_int32), ("dtype", C.c_int32), ("eps", C.c_float), ("slope", C.c_float),
                ("act", C.c_int32), ("pre_bias_n", C.c_int32), ("pre_bias", C.c_void_p),
                ("x", C.c_void_p), ("y", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("mean", C.c_void_p),
                ("rstd", C.c_void_p), ("dy", C.c_void_p), ("dx", C.c_void_p), ("dgamma_part", C.c_void_p),
                ("dbeta_part", C.c_void_p)]


class AdamWArgs(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("p_bf16", C.c_void_p),
                ("sumsq_part", C.c_void_p), ("norm_out", C.c_void_p), ("n", C.c_int64), ("n_part", C.c_int32),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("bias_corr1", C.c_float), ("bias_corr2", C.c_float), ("max_grad_norm", C.c_float), ("grad_scale", C.c_float)]


ABI_STRUCTS = (Tensor3, ScanDir, ScanFwdArgs, ScanBwdDir, ScanBwdArgs, ConvDir, ConvArgs, FbankArgs, ReduceJob,
               LayerNormArgs, DwConvArgs, SsmStepArgs, AddLnArgs, LnActArgs, AdamWArgs, FbankWavArgs, CtcArgs,
               StemArgs, ActArgs, ReduceJob2)

def lib():
    """The loaded shared library; raises (never falls back) when it is absent or stale."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "mamba_asr_b200: %s is missing - build it with `python -m mamba_asr_b200.build` "
                "(there is no CPU / Triton fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name in EXPORTS:
            getattr(L, name).restype = C.c_int
        L.cm_version.argtypes = [C.POINTER(C.c_int32)]
        L.cm_scan_num_ckpt.argtypes = [C.c_int32, C.c_int32]
        L.cm_scan_slab_channels.argtypes = [C.c_int32]
        L.cm_scan_pick_lanes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.cm_scan_pick_lanes_bwd.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.cm_scan_bwd_slab_channels.argtypes = [C.POINTER(ScanBwdArgs)]
        L.cm_scan_fwd.argtypes = [C.POINTER(ScanFwdArgs), C.c_void_p]
        L.cm_scan_bwd.argtypes = [C.POINTER(ScanBwdArgs), C.c_void_p]
        L.cm_reduce_dbc.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, Tensor3, Tensor3,
                                    C.c_void_p]
        L.cm_reduce_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.cm_conv_fwd.argtypes = [C.POINTER(ConvArgs), C.c_void_p]
        L.cm_conv_bwd.argtypes = [C.POINTER(ConvArgs), C.c_void_p]
        L.cm_conv_num_part.argtypes = [C.c_int32, C.c_int32]
        L.cm_conv_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_uint32, C.c_void_p]
        L.cm_fbank_logmel.argtypes = [C.POINTER(FbankArgs), C.c_void_p]
        L.cm_fbank_floor.argtypes = [C.POINTER(FbankArgs), C.c_void_p]
        L.cm_fbank_wav_supported.argtypes = [C.c_int32]
        L.cm_fbank_wav_logmel.argtypes = [C.POINTER(FbankWavArgs), C.c_void_p]
        L.cm_ctc_workspace_floats.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.cm_ctc_workspace_floats.restype = C.c_int64
        L.cm_ctc_loss.argtypes = [C.POINTER(CtcArgs), C.c_void_p]
        L.cm_stem_supported.argtypes = [C.c_int32, C.c_int32]
        L.cm_stem_num_part.argtypes = [C.c_int32, C.c_int32]
        L.cm_stem_fwd.argtypes = [C.POINTER(StemArgs), C.c_void_p]
        L.cm_stem_bwd.argtypes = [C.POINTER(StemArgs), C.c_void_p]
        L.cm_add_ln_dbsum_supported.argtypes = [C.c_int32, C.c_int64]
        L.cm_add_ln_num_part.argtypes = [C.c_int64, C.c_int32]
        L.cm_act_colsum_supported.argtypes = [C.c_int64, C.c_int32]
        L.cm_act_num_part.argtypes = [C.c_int64]
        L.cm_gelu_dropout_fwd_v2.argtypes = [C.POINTER(ActArgs), C.c_void_p]
        L.cm_gelu_dropout_bwd_v2.argtypes = [C.POINTER(ActArgs), C.c_void_p]
        L.cm_abi_sizeof.argtypes = [C.c_int32]
        L.cm_reduce_multi.argtypes = [C.POINTER(ReduceJob), C.c_int32, C.c_void_p]
        L.cm_scan_fwd_workspace_bytes.argtypes = [C.POINTER(ScanFwdArgs)]
        L.cm_scan_fwd_workspace_bytes.restype = C.c_int64
        L.cm_layernorm_num_part.argtypes = [C.c_int64]
        L.cm_layernorm_num_part2.argtypes = [C.c_int64, C.c_int32]
        L.cm_layernorm_fwd.argtypes = [C.POINTER(LayerNormArgs), C.c_void_p]
        L.cm_layernorm_bwd.argtypes = [C.POINTER(LayerNormArgs), C.c_void_p]
        L.cm_colsum_num_part.argtypes = [C.c_int64]
        L.cm_colsum.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
        L.cm_dwconv_num_part.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.cm_dwconv_fwd.argtypes = [C.POINTER(DwConvArgs), C.c_void_p]
        L.cm_dwconv_bwd_weight.argtypes = [C.POINTER(DwConvArgs), C.c_void_p]
        L.cm_ssm_step.argtypes = [C.POINTER(SsmStepArgs), C.c_void_p]
        L.cm_add_ln_fwd.argtypes = [C.POINTER(AddLnArgs), C.c_void_p]
        L.cm_add_ln_bwd.argtypes = [C.POINTER(AddLnArgs), C.c_void_p]
        L.cm_gelu_dropout_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p,
                                          C.c_uint32, C.c_void_p]
        L.cm_gelu_dropout_bwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float,
                                          C.c_void_p]
        L.cm_tsmm_num_part.argtypes = [C.c_int64, C.c_int32]
        L.cm_tsmm.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                              C.c_void_p]
        L.cm_ln_act_num_part.argtypes = [C.c_int64, C.c_int32]
        L.cm_ln_act_fwd.argtypes = [C.POINTER(LnActArgs), C.c_void_p]
        L.cm_ln_act_bwd.argtypes = [C.POINTER(LnActArgs), C.c_void_p]
        L.cm_reduce_batch.argtypes = [C.POINTER(ReduceJob2), C.c_int32, C.c_void_p]
        L.cm_glu_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]
        L.cm_glu_bwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64,
                                 C.c_int32, C.c_void_p]
        L.cm_optim_num_part.argtypes = [C.c_int64]
        L.cm_sumsq_partial.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.cm_adamw_step.argtypes = [C.POINTER(AdamWArgs), C.c_void_p]
        if L.cm_version(None) != CM_ABI_VERSION:
            raise RuntimeError("mamba_asr_b200: %s has a different ABI version; rebuild it" % LIB_PATH)
        for i, st in enumerate(ABI_STRUCTS):
            if L.cm_abi_sizeof(i) != C.sizeof(st):
                raise RuntimeError("mamba_asr_b200: ctypes layout of %s (%d B) differs from the library's (%d B)"
                                   % (st.__name__, C.sizeof(st), L.cm_abi_sizeof(i)))
        _lib = L
    return _lib


def dtype_code(dt):
    try:
        return _DTYPES[dt]
    except KeyError:
        raise TypeError("mamba_asr_b200 kernels take float32, bfloat16 or float16 activations, got %s" % dt)


def t3(t, order="bdl"):
    """cm_tensor3 of a 3-D torch tensor whose dims are a permutation `order` of (b, d, l); None -> absent."""
    if t is None:
        return Tensor3(None, 0, 0, 0)
    st = t.stride()
    return Tensor3(t.data_ptr(), st[order.index("b")], st[order.index("d")], st[order.index("l")])


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def check(code, what):
    if code == 0:
        return
    if code == CM_ERR_BAD_ARG:
        raise ValueError("%s: bad argument" % what)
    if code == CM_ERR_UNSUPPORTED:
        raise NotImplementedError("%s: outside the implemented envelope (dstate <= 16, conv width 2..4, batch <= 65535)"
                                  % what)
    raise RuntimeError("%s: CUDA error %d" % (what, code))
