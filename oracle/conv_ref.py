"""Depthwise causal conv1d + SiLU oracle (TEST INFRASTRUCTURE - see oracle/__init__.py).

The reference's CUDA op (causal-conv1d==1.1.3.post1, requirement.txt:7) is not in its tree; the
reference's own torch fallback is the oracle:

    modules/mamba/bimamba.py:83-91    conv1d = nn.Conv1d(D, D, kernel_size=W, groups=D, padding=W-1, bias=...)
    modules/mamba/bimamba.py:278-279  x = self.act(self.conv1d(x)[..., :seqlen])

i.e.  out[b,d,l] = act(bias[d] + sum_k w[d,k] * x[b,d,l-(W-1)+k]),  zeros outside [0, L).
``anticausal=True`` is the same op in the backward direction's original index space
(out[b,d,l] uses x[b,d,l+(W-1)-k]) and equals flip -> conv -> flip.
"""
import torch
import torch.nn.functional as F


def causal_conv1d_oracle(x, weight, bias=None, activation=None, anticausal=False,
                         compute_dtype=torch.float32):
    """x: (B, D, L); weight: (D, W); bias: (D,) or None; activation in {None, "silu", "swish"}."""
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu or swish")
    out_dtype = x.dtype
    D, W = weight.shape
    L = x.shape[-1]
    xx = x.to(compute_dtype)
    if anticausal:
        xx = xx.flip(-1)
    y = F.conv1d(xx, weight.to(compute_dtype)[:, None, :],
                 None if bias is None else bias.to(compute_dtype), padding=W - 1, groups=D)[..., :L]
    if activation is not None:
        y = F.silu(y)
    if anticausal:
        y = y.flip(-1)
    return y.to(out_dtype)


def causal_conv1d_update_oracle(x, conv_state, weight, bias=None, activation=None):
    """Single-token update (causal_conv1d_update; call site modules/mamba/bimamba.py:335-341).

    x: (B, D); conv_state: (B, D, W) rolled in place; returns (B, D).
    """
    conv_state.copy_(torch.roll(conv_state, shifts=-1, dims=-1))
    conv_state[:, :, -1] = x
    y = torch.sum(conv_state.float() * weight.float(), dim=-1)
    if bias is not None:
        y = y + bias.float()
    if activation is not None:
        y = F.silu(y)
    return y.to(x.dtype)
