"""``nn.Linear`` whose bias gradient is reduced by the sm_100a column-sum kernel (cm_colsum).

The GEMMs stay cuBLAS (forward, dx, dW - exactly what torch's own Linear backward launches); only ``db = dy.sum(rows)``
changes: torch's generic reduce_kernel ran at 1-1.5 TB/s for these (rows ~ 10^4, cols 256-1024) matrices on B200 and
cost 5.3 of 69 ms of the ConMamba-large step.  ``BiasGradLinear`` subclasses ``nn.Linear`` (same parameters, same
``state_dict``); on CPU tensors it is plain ``F.linear`` (the CPU reference arm never sees the kernel).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kernels as K


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.bias_dtype = bias.dtype
        return F.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.matmul(dy, weight)
        if ctx.needs_input_grad[1]:
            dw = torch.mm(dy2.t(), x.reshape(-1, x.shape[-1]))
        if ctx.needs_input_grad[2]:
            s = K.colsum(dy2 if dy2.stride(-1) == 1 else dy2.contiguous())
            db = (s if s is not None else dy2.sum(0)).to(ctx.bias_dtype)
        return dx, dw, db


def linear(x, weight, bias=None):
    """F.linear with the bias gradient on the sm_100a kernel (CUDA tensors with a bias that needs grad); F.linear otherwise."""
    if bias is None or not x.is_cuda or not (torch.is_grad_enabled() and bias.requires_grad):
        return F.linear(x, weight, bias)
    if torch.is_autocast_enabled("cuda"):          # the casts autocast would insert, visible to autograd
        dt = torch.get_autocast_dtype("cuda")
        x, weight, bias = x.to(dt), weight.to(dt), bias.to(dt)
        with torch.autocast("cuda", enabled=False):
            return _LinearFn.apply(x, weight, bias)
    return _LinearFn.apply(x, weight, bias)


class BiasGradLinear(nn.Linear):
    def forward(self, x):
        return linear(x, self.weight, self.bias)
