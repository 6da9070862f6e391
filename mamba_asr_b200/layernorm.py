"""LayerNorm over the last dimension on the sm_100a kernels (cm_layernorm_fwd / cm_layernorm_bwd).

SURVEY.md section 8(f) rank 2: the ConMamba layer wraps every sub-block in a LayerNorm (reference
modules/Conmamba.py:595-621, 638-649).  Under bf16 autocast torch runs each of them as cast -> fp32 kernel -> cast, plus
a separate gamma/beta-gradient kernel in backward - measured 7 ms of a 29.6 ms ConMamba-small step on B200
(gpurun_out/step_profile_cfg2.log).  The kernel reads the residual stream in its own dtype, keeps statistics in fp32
and writes the dtype the consumer wants (the autocast dtype when autocast is on - exactly what the following Linear
would cast to), one launch forward, one launch + one deterministic reduction backward.

``FusedLayerNorm`` subclasses ``nn.LayerNorm`` so parameters, ``state_dict`` keys and ``extra_repr`` are unchanged.
There is no CPU path: the CPU reference arm (oracle/cpu_encoder.py) swaps the class back to ``nn.LayerNorm``.
"""
import os

import torch
import torch.nn as nn

from . import kernels as K


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype, gelu=False):
        Cn = x.shape[-1]
        x2 = x.reshape(-1, Cn)
        if x2.stride(1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < Cn):
            x2 = x2.contiguous()
        y, mean, rstd = K.layernorm_forward(x2, weight, bias, eps, out_dtype, gelu=gelu)
        ctx.gelu = gelu
        ctx.save_for_backward(x2, weight, mean, rstd, *([bias] if (gelu and bias is not None) else []))
        ctx.x_shape = x.shape
        ctx.has_bias = bias is not None
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, mean, rstd = ctx.saved_tensors[:4]
        beta = ctx.saved_tensors[4] if len(ctx.saved_tensors) > 4 else None
        Cn = x2.shape[1]
        need_w = weight is not None and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        dx, dg, db = K.layernorm_backward(x2, dy.reshape(-1, Cn), weight, mean, rstd, need_wgrad=need_w, defer=True,
                                          gelu=ctx.gelu, bias=beta)
        dgw = K.grad_cast(dg, weight.dtype) if (need_w and ctx.needs_input_grad[1]) else None
        dbw = K.grad_cast(db, weight.dtype) if (need_w and ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx.view(ctx.x_shape), dgw, dbw, None, None, None


def layer_norm_gelu_supported(x):
    """True when ``layer_norm(..., gelu=True)`` can take x (see kernels.layernorm_gelu_supported) - checked on the 2-D view the
    kernel sees."""
    if not x.is_cuda or x.dim() < 2 or os.environ.get("CM_NO_LN_GELU_EPILOGUE") is not None:
        return False
    Cn = x.shape[-1]
    return x.is_contiguous() and Cn % 2 == 0 and Cn <= 1024 and x.data_ptr() % 8 == 0


def layer_norm(x, weight, bias, eps=1e-5, out_dtype=None, gelu=False):
    """Functional form.  ``out_dtype`` None: the autocast dtype if CUDA autocast is enabled, else x.dtype.
    ``gelu=True``: gelu(LayerNorm(x)) (exact erf GELU, as nn.GELU()) as an epilogue of the same kernels - the LayerNorm ->
    GELU pair after the depthwise conv of the convolution module (reference modules/Conmamba.py:292-301)."""
    if not x.is_cuda:
        raise RuntimeError("mamba_asr_b200.layer_norm runs on the sm_100a kernel only (no CPU fallback)")
    if out_dtype is None:
        out_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
    if weight is not None and weight.dtype != torch.float32:
        weight = weight.float()
    if bias is not None and bias.dtype != torch.float32:
        bias = bias.float()
    return _LayerNormFn.apply(x, weight, bias, eps, out_dtype, bool(gelu))


class FusedLayerNorm(nn.LayerNorm):
    """``nn.LayerNorm(normalized_shape=int)`` evaluated by the sm_100a kernel.  Default output dtype: the autocast dtype
    under autocast (the norm feeds a Linear / conv that would cast to it anyway), else the input's.  ``keep_dtype=True``
    reproduces torch's own autocast rule instead - fp32 under autocast - for the norms whose output IS the residual
    stream (``norm2`` of a layer, the encoder's final norm)."""

    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True, bias=True, keep_dtype=False, **kw):
        super().__init__(normalized_shape, eps=eps, elementwise_affine=elementwise_affine, bias=bias, **kw)
        if len(self.normalized_shape) != 1 or self.normalized_shape[0] > 1024:
            raise NotImplementedError("FusedLayerNorm normalises one trailing dimension of at most 1024 channels")
        self.keep_dtype = keep_dtype

    def forward(self, x, gelu=False):
        out_dtype = None
        if self.keep_dtype:
            out_dtype = torch.float32 if torch.is_autocast_enabled("cuda") else x.dtype
        return layer_norm(x, self.weight, self.bias, self.eps, out_dtype=out_dtype, gelu=gelu)


# ---------------------------------------------------------------------------------------------------------------------
# LayerNorm over several trailing dimensions + activation in one pass (cm_ln_act_fwd / cm_ln_act_bwd): the conv blocks of
# the CNN front-end (reference hparams/CTC/conmamba_large.yaml:187-199; LeakyReLU, conv bias folded in) and the
# LayerNorm -> activation after the depthwise conv of the convolution module (reference modules/Conmamba.py:292-301)
class _LayerNormActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, pre_bias, eps, slope, act, n_norm):
        x2 = x.reshape(-1, n_norm)
        w, b = weight.reshape(-1), bias.reshape(-1)
        y, mean, rstd = K.ln_act_forward(x2, w, b, eps, slope, act, pre_bias)
        ctx.save_for_backward(x2, w, b, pre_bias, mean, rstd)
        ctx.slope, ctx.act = slope, act
        ctx.w_shape = weight.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, w, b, pre_bias, mean, rstd = ctx.saved_tensors
        need_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx, dg, db = K.ln_act_backward(x2, dy.reshape(-1, x2.shape[1]), w, b, mean, rstd, ctx.slope, ctx.act, pre_bias,
                                       need_wgrad=need_w, defer=True)
        dpb = None
        if pre_bias is not None and ctx.needs_input_grad[3]:
            cs = K.colsum(dx)                                     # (n_norm,) fp32, fixed-order sums
            if cs is None:
                cs = dx.float().sum(0)
            dpb = cs.view(-1, pre_bias.numel()).sum(0)
        return (dx.view(dy.shape), dg.view(ctx.w_shape) if ctx.needs_input_grad[1] else None,
                db.view(ctx.w_shape) if ctx.needs_input_grad[2] else None, dpb, None, None, None, None)


def layer_norm_act(x, norm, act="leaky_relu", negative_slope=0.01, pre_bias=None):
    """``act(norm(x + pre_bias))`` for an affine ``nn.LayerNorm`` over any number of trailing dimensions as ONE sm_100a
    kernel forward and one backward; ``act`` is "leaky_relu" or "gelu" (exact).  ``pre_bias`` (n,) is added along the
    last dimension before the statistics (n = the last dimension: the bias of the conv that produced x).  The result
    keeps x's dtype (under bf16 autocast torch computes LayerNorm and the activation in fp32 and the consumer casts back
    to bf16: the same single rounding).  No CPU path."""
    if not x.is_cuda:
        raise RuntimeError("mamba_asr_b200.layer_norm_act runs on the sm_100a kernel only (no CPU fallback)")
    n_norm = 1
    for d in norm.normalized_shape:
        n_norm *= d
    if tuple(x.shape[-len(norm.normalized_shape):]) != tuple(norm.normalized_shape):
        raise ValueError("layer_norm_act: trailing dimensions %s do not match normalized_shape %s"
                         % (tuple(x.shape), tuple(norm.normalized_shape)))
    if norm.weight is None or norm.bias is None:
        raise NotImplementedError("layer_norm_act needs an affine LayerNorm with bias")
    if pre_bias is not None and pre_bias.numel() != x.shape[-1]:
        raise ValueError("layer_norm_act: pre_bias must have the length of the last dimension")
    xc = x if x.is_contiguous() else x.contiguous()
    if not K.ln_act_supported(xc.reshape(-1, n_norm)) or (pre_bias is not None and pre_bias.numel() % 4 != 0):
        raise NotImplementedError("layer_norm_act: rows of %d elements are outside the kernel envelope "
                                  "(multiple of 4, <= %d)" % (n_norm, K.LN_ACT_MAX_COLS))
    with torch.autocast("cuda", enabled=False):
        return _LayerNormActFn.apply(xc, norm.weight.float(), norm.bias.float(),
                                     None if pre_bias is None else pre_bias.float(), norm.eps, float(negative_slope), act,
                                     n_norm)


def layer_norm_leaky_relu(x, norm, negative_slope=0.01, pre_bias=None):
    """``leaky_relu(norm(x + pre_bias))``: see ``layer_norm_act``."""
    return layer_norm_act(x, norm, "leaky_relu", negative_slope, pre_bias)


# ---------------------------------------------------------------------------------------------------------------------
# first block of the CNN front-end: Conv2d(1 -> C, 3 x 3, stride 2, padding 1) + LayerNorm([F', C]) + LeakyReLU as one kernel
# each way (cm_stem_fwd / cm_stem_bwd; reference hparams/CTC/conmamba_large.yaml:187-199).  Only the features, mean and rstd
# are saved: the conv output is recomputed in backward and never written.
class _StemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, weight, bias, gamma, beta, eps, slope, out_dtype):
        y, mean, rstd = K.stem_forward(feats, weight, bias, gamma, beta, eps, slope, out_dtype)
        ctx.save_for_backward(feats, weight, bias, gamma, beta, mean, rstd)
        ctx.slope = slope
        return y

    @staticmethod
    def backward(ctx, dy):
        feats, weight, bias, gamma, beta, mean, rstd = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("conv_ln_act_stem: no gradient with respect to the features (they are the network input)")
        dw, dcb, dg, db = K.stem_backward(feats, dy, weight, bias, gamma, beta, mean, rstd, ctx.slope, defer=True)
        return (None, dw if ctx.needs_input_grad[1] else None, dcb if bias is not None and ctx.needs_input_grad[2] else None,
                dg if ctx.needs_input_grad[3] else None, db if ctx.needs_input_grad[4] else None, None, None, None)


def conv_ln_act_stem_supported(feats, conv, norm):
    """True when ``conv_ln_act_stem`` can run this block: CUDA features (B, T, F) without gradient, a 1-input-channel 3 x 3
    stride-2 padding-1 zero-padded conv, an affine LayerNorm over (F', C) inside the kernel envelope."""
    if not isinstance(conv, torch.nn.Conv2d) or conv.in_channels != 1 or conv.kernel_size != (3, 3) or conv.stride != (2, 2):
        return False
    if conv.padding != (1, 1) or conv.dilation != (1, 1) or conv.groups != 1 or conv.padding_mode != "zeros":
        return False
    if norm.weight is None or norm.bias is None or feats.dim() != 3 or feats.requires_grad:
        return False
    if tuple(norm.normalized_shape) != ((feats.shape[2] - 1) // 2 + 1, conv.out_channels):
        return False
    return K.stem_supported(feats, conv.weight)


def conv_ln_act_stem(feats, conv, norm, negative_slope=0.01):
    """``leaky_relu(norm(conv(feats[:, None]).permute(0, 2, 3, 1)))`` -> (B, T', F', C) in the autocast dtype (fp32 outside
    autocast), computed in fp32 from the features as given.  No CPU path."""
    if not feats.is_cuda:
        raise RuntimeError("mamba_asr_b200.conv_ln_act_stem runs on the sm_100a kernel only (no CPU fallback)")
    if not conv_ln_act_stem_supported(feats, conv, norm):
        raise NotImplementedError("conv_ln_act_stem: block outside the kernel envelope")
    out_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else feats.dtype
    with torch.autocast("cuda", enabled=False):
        return _StemFn.apply(feats.contiguous(), conv.weight.float(), None if conv.bias is None else conv.bias.float(),
                             norm.weight.float(), norm.bias.float(), norm.eps, float(negative_slope), out_dtype)


# ---------------------------------------------------------------------------------------------------------------------
# residual add + dropout + LayerNorm in one pass (cm_add_ln_fwd / cm_add_ln_bwd)
class DropoutSeed:
    """Device-resident seed of the fused dropout masks.  Every call site gets its own ``call_id`` (a host counter); the
    int64 device scalar is advanced once per step by ``advance()`` - an in-place add that a captured CUDA graph replays,
    so each replay draws fresh masks although the call ids are baked into the graph."""

    _state = {}
    _calls = 0

    @classmethod
    def tensor(cls, device):
        key = (device.type, device.index)
        t = cls._state.get(key)
        if t is None:
            t = torch.full((1,), torch.initial_seed() & 0x7fffffffffffffff, dtype=torch.int64, device=device)
            cls._state[key] = t
        return t

    @classmethod
    def advance(cls, device):
        cls.tensor(device).add_(0x9E3779B97F4A7C15 & 0x7fffffffffffffff)

    @classmethod
    def next_call_id(cls):
        cls._calls += 1
        return cls._calls


class _AddDropoutLayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, weight, bias, eps, alpha, p_drop, out_dtype, b_bias):
        # b_bias: the bias parameter of the Linear that produced b (or None).  It does not enter the forward; its gradient -
        # the column sums of db - comes out of the backward kernel instead of one more pass over db in that Linear's backward.
        Cn = a.shape[-1]
        a2 = a.reshape(-1, Cn)
        b2 = None if b is None else b.reshape(-1, Cn)
        seed = DropoutSeed.tensor(a.device) if p_drop > 0.0 else None
        s, y, mean, rstd, saved = K.add_ln_forward(a2, b2, weight, bias, eps, alpha, p_drop, seed,
                                                   DropoutSeed.next_call_id(), out_dtype)
        ctx.save_for_backward(s, weight, mean, rstd, saved)
        ctx.shape = a.shape
        ctx.alpha, ctx.p_drop = alpha, p_drop
        ctx.b_dtype = None if b is None else b.dtype
        ctx.has_bias = bias is not None
        ctx.b_bias_dtype = None if b_bias is None else b_bias.dtype
        return s.view(a.shape), y.view(a.shape)

    @staticmethod
    def backward(ctx, ds, dy):
        s, weight, mean, rstd, saved = ctx.saved_tensors
        Cn = s.shape[1]
        if dy is None:
            dy = torch.zeros(ctx.shape, dtype=s.dtype, device=s.device)
        need_w = weight is not None and (ctx.needs_input_grad[2] or ctx.needs_input_grad[3])
        need_bb = ctx.b_bias_dtype is not None and ctx.needs_input_grad[8] and ctx.b_dtype is not None
        da, db, dg, dbt, dbs = K.add_ln_backward(s, dy.reshape(-1, Cn), None if ds is None else ds.reshape(-1, Cn), weight,
                                                 mean, rstd, saved, ctx.alpha, ctx.p_drop, ctx.b_dtype or s.dtype,
                                                 need_db=ctx.b_dtype is not None, need_wgrad=need_w, need_dbsum=True,
                                                 defer=True)
        dbb = None
        if need_bb:
            if dbs is None:                                      # geometry outside the quad kernels: the separate pass
                dbs = K.colsum(db, defer=True)
                if dbs is None:
                    dbs = db.float().sum(0)
            dbb = K.grad_cast(dbs, ctx.b_bias_dtype)
        return (da.view(ctx.shape), None if db is None else db.view(ctx.shape),
                K.grad_cast(dg, weight.dtype) if (need_w and ctx.needs_input_grad[2]) else None,
                K.grad_cast(dbt, weight.dtype) if (need_w and ctx.has_bias and ctx.needs_input_grad[3]) else None,
                None, None, None, None, dbb)


def add_ln_kernel_ok(a, b, norm):
    """True when ``add_dropout_layer_norm(a, b, norm, ...)`` runs on the fused kernel (cm_add_ln_*) in the current autocast
    state - the condition under which ``b_bias`` may be passed."""
    if norm.keep_dtype:
        out_dtype = torch.float32 if torch.is_autocast_enabled("cuda") else a.dtype
    elif torch.is_autocast_enabled("cuda"):
        out_dtype = torch.get_autocast_dtype("cuda")
    else:
        out_dtype = a.dtype
    Cn = a.shape[-1]
    return (b is not None and a.is_cuda and a.is_contiguous() and b.is_contiguous() and b.shape == a.shape
            and K.add_ln_supported(a.reshape(-1, Cn), b.reshape(-1, Cn), out_dtype))


def add_dropout_layer_norm(a, b, norm, alpha=1.0, p_drop=0.0, training=True, b_bias=None):
    """(s, y) with  s = a + alpha * dropout(b, p_drop)  and  y = norm(s)  for a ``FusedLayerNorm`` ``norm`` - one kernel
    forward, one backward (cm_add_ln_*) when the combination is implemented, else the separate ops.  ``b`` may be None
    (s = a).  ``b_bias``: the bias parameter of the Linear that produced b, when that Linear was evaluated with
    ``linear(..., bias_grad=False)``: its gradient (the column sums of db) then comes out of this op's backward kernel
    (only valid when ``add_ln_kernel_ok``)."""
    p = float(p_drop) if training else 0.0
    out_dtype = None
    if norm.keep_dtype:
        out_dtype = torch.float32 if torch.is_autocast_enabled("cuda") else a.dtype
    elif torch.is_autocast_enabled("cuda"):
        out_dtype = torch.get_autocast_dtype("cuda")
    else:
        out_dtype = a.dtype
    Cn = a.shape[-1]
    ok = (b is not None and a.is_cuda and a.is_contiguous() and b.is_contiguous() and b.shape == a.shape
          and K.add_ln_supported(a.reshape(-1, Cn), b.reshape(-1, Cn), out_dtype))
    if b_bias is not None and not (torch.is_grad_enabled() and b_bias.requires_grad):
        b_bias = None
    if not ok:
        if b is None:
            s = a
        else:
            if b_bias is not None:                               # the promised bias gradient, by the separate pass
                b = _BiasGradRoute.apply(b, b_bias)
            s = a + alpha * (torch.nn.functional.dropout(b, p, training=True) if p > 0 else b)
        return s, norm(s)
    w = norm.weight if norm.weight is None or norm.weight.dtype == torch.float32 else norm.weight.float()
    bb = norm.bias if norm.bias is None or norm.bias.dtype == torch.float32 else norm.bias.float()
    with torch.autocast("cuda", enabled=False):
        return _AddDropoutLayerNormFn.apply(a, b, w, bb, norm.eps, float(alpha), p, out_dtype, b_bias)


class _BiasGradRoute(torch.autograd.Function):
    """Identity on ``y``; the gradient of ``bias`` is the column sum of the gradient of ``y`` (for a Linear evaluated with
    ``bias_grad=False`` whose consumer cannot form that sum itself)."""

    @staticmethod
    def forward(ctx, y, bias):
        ctx.bias_dtype = bias.dtype
        return y.view_as(y)

    @staticmethod
    def backward(ctx, dy):
        d2 = dy.reshape(-1, dy.shape[-1])
        cs = K.colsum(d2 if d2.stride(-1) == 1 else d2.contiguous(), defer=True) if d2.is_cuda else None
        if cs is None:
            cs = d2.float().sum(0)
        return dy, K.grad_cast(cs, ctx.bias_dtype)


# ---------------------------------------------------------------------------------------------------------------------
# GELU + dropout in one pass (cm_gelu_dropout_fwd / cm_gelu_dropout_bwd)
class _GeluDropoutFn(torch.autograd.Function):
    """``bias`` is the bias of the Linear that produced ``x`` (or None): it does not enter the forward, but its gradient - the
    column sums of dx - is formed inside the backward kernel instead of by one more pass over dx in the Linear's backward."""

    @staticmethod
    def forward(ctx, x, p_drop, bias):
        seed = DropoutSeed.tensor(x.device) if p_drop > 0.0 else None
        y, saved = K.gelu_dropout_forward(x, p_drop, seed, DropoutSeed.next_call_id())
        ctx.save_for_backward(x, saved)
        ctx.p_drop = p_drop
        ctx.bias_dtype = None if bias is None else bias.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        x, saved = ctx.saved_tensors
        if ctx.bias_dtype is None or not ctx.needs_input_grad[2]:
            return K.gelu_dropout_backward(x, dy, saved, ctx.p_drop), None, None
        cols = x.shape[-1]
        dx, cs = K.gelu_dropout_backward(x, dy, saved, ctx.p_drop, colsum_cols=cols, defer=True)
        if cs is None:
            cs = K.colsum(dx.reshape(-1, cols), defer=True)
            if cs is None:
                cs = dx.reshape(-1, cols).float().sum(0)
        return dx, None, K.grad_cast(cs, ctx.bias_dtype)


class _GluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h):
        ctx.save_for_backward(h)
        return K.glu_forward(h)

    @staticmethod
    def backward(ctx, dy):
        (h,) = ctx.saved_tensors
        return K.glu_backward(h, dy)


def glu(h):
    """F.glu(h, dim=-1) on the sm_100a kernels (cm_glu_fwd / cm_glu_bwd: 16-byte accesses at the HBM rate) for contiguous CUDA
    tensors the kernels take; torch's op otherwise (and under CM_NO_GLU_KERNEL=1, the A/B switch)."""
    if os.environ.get("CM_NO_GLU_KERNEL") is None and K.glu_supported(h):
        return _GluFn.apply(h)
    return torch.nn.functional.glu(h, dim=-1)


def gelu_dropout(x, p_drop=0.0, training=True, bias_for_grad=None):
    """dropout(gelu(x)) (exact erf GELU, as ``nn.GELU()``) - one sm_100a kernel forward and one backward on CUDA tensors
    the kernel supports, the two torch ops otherwise.  ``bias_for_grad``: the bias parameter of the Linear that produced x,
    when that Linear was evaluated with ``linear(..., bias_grad=False)``: its gradient then comes out of this op's backward."""
    p = float(p_drop) if training else 0.0
    if not K.gelu_dropout_supported(x):
        if bias_for_grad is not None:
            raise NotImplementedError("gelu_dropout: bias_for_grad needs the kernel path (check gelu_dropout_supported first)")
        y = torch.nn.functional.gelu(x)
        return torch.nn.functional.dropout(y, p, training=True) if p > 0 else y
    if bias_for_grad is not None and not (torch.is_grad_enabled() and bias_for_grad.requires_grad):
        bias_for_grad = None
    return _GeluDropoutFn.apply(x, p, bias_for_grad)
