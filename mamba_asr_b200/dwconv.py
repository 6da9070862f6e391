"""Depthwise conv1d over time on the sm_100a kernels (cm_dwconv_fwd / cm_dwconv_bwd_weight), channel-last.

SURVEY.md section 8(f) rank 2: the kernel_size = 31 depthwise convolution of the ConMamba convolution module (reference
modules/Conmamba.py:281-290).  ``depthwise_conv1d(x, weight, bias, pad_left)`` takes (B, L, C) activations - the layout
the LayerNorm before it and the Linear after it already use - so the module's two transposes disappear with the torch
conv_depthwise2d kernels (measured 14.2 of 88 ms of kernel time in the ConMamba-large step on B200,
gpurun_out/step_profile_large.log).  No CPU path.
"""
import torch

from . import kernels as K


class _DwConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, pad_left):
        w2 = weight.reshape(weight.shape[0], weight.shape[-1])
        w2 = w2.float().contiguous() if (w2.dtype != torch.float32 or not w2.is_contiguous()) else w2
        b2 = None if bias is None else (bias if bias.dtype == torch.float32 else bias.float())
        y = K.dwconv_forward(x, w2, b2, pad_left)
        ctx.save_for_backward(x, w2)
        ctx.pad_left = pad_left
        ctx.w_shape, ctx.w_dtype = weight.shape, weight.dtype
        ctx.has_bias = bias is not None
        ctx.b_dtype = None if bias is None else bias.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w2 = ctx.saved_tensors
        Kk = w2.shape[1]
        if dy.stride(-1) != 1:
            dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = K.dwconv_forward(dy, w2, None, Kk - 1 - ctx.pad_left, flip=True)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw2, db2 = K.dwconv_backward_weight(x, dy, Kk, ctx.pad_left, need_bias=ctx.has_bias, defer=True)
            if ctx.needs_input_grad[1]:
                dw = K.grad_cast(dw2.view(ctx.w_shape), ctx.w_dtype)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = K.grad_cast(db2, ctx.b_dtype)
        return dx, dw, db, None


def depthwise_conv1d(x, weight, bias=None, pad_left=None):
    """y[b,l,c] = bias[c] + sum_k weight[c,(0,)k] * x[b, l - pad_left + k, c].  x: (B, L, C) CUDA tensor with unit channel
    stride; weight: (C, K) or nn.Conv1d's (C, 1, K); pad_left defaults to (K-1)//2 ("same")."""
    if not x.is_cuda:
        raise RuntimeError("mamba_asr_b200.depthwise_conv1d runs on the sm_100a kernel only (no CPU fallback)")
    Kk = weight.shape[-1]
    if pad_left is None:
        pad_left = (Kk - 1) // 2
    if x.stride(-1) != 1:
        x = x.contiguous()
    return _DwConvFn.apply(x, weight, bias, int(pad_left))
