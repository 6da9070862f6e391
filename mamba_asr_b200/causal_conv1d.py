"""Drop-in for the ``causal_conv1d`` pip package as the reference uses it
(``from causal_conv1d import causal_conv1d_fn, causal_conv1d_update``, modules/mamba/bimamba.py:19;
call sites bimamba.py:282-287, :335-341 and selective_scan_interface.py:14).

Backed by cm_conv_fwd / cm_conv_bwd / cm_conv_update; CUDA tensors only, no fallback.
"""
import torch

from . import kernels as K


class CausalConv1dFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, weight, bias=None, seq_idx=None, activation=None):
        if activation not in [None, "silu", "swish"]:
            raise NotImplementedError("activation must be None, silu, or swish")
        if seq_idx is not None:
            raise NotImplementedError("seq_idx (packed variable-length batches) is not used by the ConMamba path")
        ctx.silu = activation in ["silu", "swish"]
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight, bias)
        return K.conv_forward(x, [dict(weight=weight, bias=bias, anticausal=False)], silu=ctx.silu)[0]

    @staticmethod
    def backward(ctx, dout):
        x, weight, bias = ctx.saved_tensors
        dx, dws, dbs = K.conv_backward(x, [dict(weight=weight, bias=bias, anticausal=False)], [dout], silu=ctx.silu)
        return dx, dws[0].to(weight.dtype), (dbs[0].to(bias.dtype) if ctx.has_bias else None), None, None


def causal_conv1d_fn(x, weight, bias=None, seq_idx=None, activation=None):
    """x: (batch, dim, seqlen); weight: (dim, width); bias: (dim,); activation: None | "silu" | "swish".
    out: (batch, dim, seqlen)."""
    return CausalConv1dFn.apply(x, weight, bias, seq_idx, activation)


def causal_conv1d_update(x, conv_state, weight, bias=None, activation=None):
    """x: (batch, dim); conv_state: (batch, dim, width), rolled in place; returns (batch, dim)."""
    if activation not in [None, "silu", "swish"]:
        raise NotImplementedError("activation must be None, silu, or swish")
    return K.conv_update(x, conv_state, weight, bias, silu=activation in ["silu", "swish"])
