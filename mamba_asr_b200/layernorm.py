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
import torch
import torch.nn as nn

from . import kernels as K


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        Cn = x.shape[-1]
        x2 = x.reshape(-1, Cn)
        if x2.stride(1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < Cn):
            x2 = x2.contiguous()
        y, mean, rstd = K.layernorm_forward(x2, weight, bias, eps, out_dtype)
        ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.x_shape = x.shape
        ctx.has_bias = bias is not None
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, mean, rstd = ctx.saved_tensors
        Cn = x2.shape[1]
        need_w = weight is not None and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        dx, dg, db = K.layernorm_backward(x2, dy.reshape(-1, Cn), weight, mean, rstd, need_wgrad=need_w)
        dgw = dg.to(weight.dtype) if (need_w and ctx.needs_input_grad[1]) else None
        dbw = db.to(weight.dtype) if (need_w and ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx.view(ctx.x_shape), dgw, dbw, None, None


def layer_norm(x, weight, bias, eps=1e-5, out_dtype=None):
    """Functional form.  ``out_dtype`` None: the autocast dtype if CUDA autocast is enabled, else x.dtype."""
    if not x.is_cuda:
        raise RuntimeError("mamba_asr_b200.layer_norm runs on the sm_100a kernel only (no CPU fallback)")
    if out_dtype is None:
        out_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
    if weight is not None and weight.dtype != torch.float32:
        weight = weight.float()
    if bias is not None and bias.dtype != torch.float32:
        bias = bias.float()
    return _LayerNormFn.apply(x, weight, bias, eps, out_dtype)


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

    def forward(self, x):
        out_dtype = None
        if self.keep_dtype:
            out_dtype = torch.float32 if torch.is_autocast_enabled("cuda") else x.dtype
        return layer_norm(x, self.weight, self.bias, self.eps, out_dtype=out_dtype)
