"""Linear layers of the shell around the hot path: cuBLAS GEMMs with the non-GEMM work on sm_100a kernels or removed.

* bias gradient: ``db = dy.sum(rows)`` on the streaming column-sum kernel (cm_colsum) instead of torch's generic
  reduce_kernel (1-1.5 TB/s for rows ~ 10^4, cols 256-1024 on B200; 5.3 of 69 ms of the ConMamba-large step);
* low-precision parameter copies: under bf16 autocast every use of a weight costs a cast kernel forward and a cast of
  its gradient backward (750 launches = 10 % of the ConMamba-small step on B200).  ``ParamCache`` keeps bf16 copies of
  all parameters of a model in one flat buffer, refreshed by ONE multi-tensor copy per step, and the weight-gradient
  GEMMs write fp32 directly (``torch.mm(..., out_dtype=float32)``), so no per-parameter cast kernel remains.

The GEMMs stay cuBLAS (forward, dx, dW - what torch's own Linear backward launches).  ``BiasGradLinear`` subclasses
``nn.Linear`` (same parameters, same ``state_dict``); on CPU tensors it is plain ``F.linear`` (the CPU reference arm never
sees the kernels).
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kernels as K


class ParamCache:
    """bf16 (autocast dtype) copies of a module tree's fp32 parameters in one flat buffer.

    ``refresh()`` is one ``torch._foreach_copy_`` (a multi-tensor kernel) - call it once per step before the forward
    (the model shells in ``encoder.py`` do, inside the captured graph).  ``get(p)`` returns the copy of parameter ``p`` or
    None.  Only consulted under CUDA autocast; gradients always flow to the fp32 parameters."""

    def __init__(self, module, dtype=torch.bfloat16):
        self.params = [p for p in module.parameters() if p.is_cuda and p.dtype == torch.float32]
        self.dtype = dtype
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + 7) // 8 * 8                      # 16-byte aligned views (TMA / vector loads downstream)
        dev = self.params[0].device if self.params else None
        self.flat = torch.zeros(n, dtype=dtype, device=dev)
        self.views = [self.flat[o:o + p.numel()].view(p.shape) for o, p in zip(offs, self.params)]
        self.index = {id(p): i for i, p in enumerate(self.params)}
        self.versions = [-1] * len(self.params)                # parameter versions the copies were taken at
        self.refresh()

    def refresh(self):
        with torch.no_grad():
            torch._foreach_copy_(self.views, self.params)
        self.versions = [p._version for p in self.params]

    def get(self, p, dtype):
        """The copy of ``p``, or None when there is none or when ``p`` has been written (optimizer step, load_state_dict,
        in-place init) since the last ``refresh()`` - a stale copy is never handed out; the caller then casts ``p`` itself.
        (A captured CUDA graph replays the refresh kernel itself, so the host-side version check only guards eager use.)"""
        if dtype != self.dtype:
            return None
        i = self.index.get(id(p))
        if i is None or self.versions[i] != p._version:
            return None
        return self.views[i]


_ACTIVE = None


def set_param_cache(cache):
    """Install (or clear, with None) the cache ``linear`` consults."""
    global _ACTIVE
    _ACTIVE = cache


def invalidate_param_cache():
    """Mark every copy of the installed cache stale (called by code that writes parameters through raw pointers, e.g. the
    flat-buffer optimizer kernel, which does not bump the tensors' version counters)."""
    if _ACTIVE is not None:
        _ACTIVE.versions = [-1] * len(_ACTIVE.versions)


def cached_param(p, dtype):
    """The cached low-precision copy of ``p`` (no autograd link) or None."""
    return _ACTIVE.get(p, dtype) if (_ACTIVE is not None and p is not None) else None


def _wgrad_nsplit(rows, M, N):
    """Row blocks for the split-K weight-gradient GEMM (measured on B200, tools/prof_wgrad.py, profiles/r01_wgrad_variants_
    session4.txt): as ONE GEMM with K = batch * L in the tens of thousands cuBLAS runs these shapes at 110-440 TFLOP/s
    (30.7 us for 12032 x 1024 x 144); a bmm over 4-16 row blocks plus the fixed-order sum of the fp32 partial products takes
    19.5 us, and is the more accurate evaluation."""
    if os.environ.get("CM_NO_WGRAD_SPLIT") is not None:
        return 1
    if rows >= 24000 and M * N <= (1 << 17):
        want = 16
    elif rows * max(M, N) <= (1 << 24) and M * N >= (1 << 17):
        want = 4
    else:
        want = 8
    while want > 1 and rows % want != 0:
        want //= 2
    return want if rows >= 4096 else 1


def _wgrad_f32(dy2, x2):
    """dy2^T @ x2 with fp32 output for 16-bit operands (rows, M) and (rows, N): deterministic split-K (see above)."""
    rows = dy2.shape[0]
    ns = _wgrad_nsplit(rows, dy2.shape[1], x2.shape[1])
    if ns <= 1:
        return torch.mm(dy2.t(), x2, out_dtype=torch.float32)          # fp32 straight out of the GEMM: no cast kernel
    d3 = dy2.reshape(ns, rows // ns, dy2.shape[1])
    x3 = x2.reshape(ns, rows // ns, x2.shape[1])
    return K.sum_leading(torch.bmm(d3.transpose(1, 2), x3, out_dtype=torch.float32), defer=True)   # queued under deferral


class _LinearFn(torch.autograd.Function):
    """y = x @ w^T + b with explicit low-precision operands: ``w_lp`` / ``b_lp`` are the forward operands (cached copies
    or ``weight`` / ``bias`` themselves), gradients are returned for ``weight`` / ``bias`` in THEIR dtype."""

    @staticmethod
    def forward(ctx, x, weight, bias, w_lp, b_lp):
        ctx.save_for_backward(x, w_lp)
        ctx.w_dtype = weight.dtype
        ctx.b_dtype = None if bias is None else bias.dtype
        return F.linear(x, w_lp, b_lp)

    @staticmethod
    def backward(ctx, dy):
        x, w_lp = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.matmul(dy, w_lp)
        if ctx.needs_input_grad[1]:
            x2 = x.reshape(-1, x.shape[-1])
            if ctx.w_dtype == torch.float32 and dy2.dtype != torch.float32:
                dw = _wgrad_f32(dy2, x2)
            else:
                dw = torch.mm(dy2.t(), x2).to(ctx.w_dtype)
        if ctx.b_dtype is not None and ctx.needs_input_grad[2]:
            s = K.colsum(dy2 if dy2.stride(-1) == 1 else dy2.contiguous(), defer=True)
            db = K.grad_cast(s if s is not None else dy2.sum(0), ctx.b_dtype)
        return dx, dw, db, None, None


def linear(x, weight, bias=None, bias_grad=True):
    """F.linear for CUDA training: bias gradient on cm_colsum, parameter casts from the ``ParamCache`` when one is
    installed, fp32 weight gradients straight from the GEMM.  Plain F.linear on CPU or without grad.
    ``bias_grad=False``: the caller obtains the bias gradient elsewhere (the consumer of the output sums its own input
    gradient over the rows, e.g. ``gelu_dropout(..., bias_for_grad=bias)``): no column-sum pass here."""
    if not x.is_cuda or not torch.is_grad_enabled() or not (weight.requires_grad or (bias is not None and bias.requires_grad)):
        return F.linear(x, weight, bias)
    if bias is not None and not bias_grad:
        b_src, bias = bias, bias.detach()
    else:
        b_src = bias
    if torch.is_autocast_enabled("cuda"):          # the casts autocast would insert
        dt = torch.get_autocast_dtype("cuda")
        w_lp = cached_param(weight, dt)
        b_lp = cached_param(b_src, dt)
        if w_lp is None:
            w_lp = weight.detach().to(dt)
        if bias is not None and b_lp is None:
            b_lp = bias.detach().to(dt)
        with torch.autocast("cuda", enabled=False):
            return _LinearFn.apply(x.to(dt), weight, bias, w_lp, b_lp)
    if bias is None:
        return F.linear(x, weight, bias)
    return _LinearFn.apply(x, weight, bias, weight.detach(), bias.detach())


class BiasGradLinear(nn.Linear):
    def forward(self, x):
        return linear(x, self.weight, self.bias)
