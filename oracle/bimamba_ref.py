"""BiMamba-v2 block oracle (TEST INFRASTRUCTURE - see oracle/__init__.py).

Composes the scan and conv oracles exactly as the reference's v2 fast path does
(modules/mamba/bimamba.py:192-253 calling MambaInnerFnNoOutProj.forward,
modules/mamba/selective_scan_interface.py:164-229), *with the flips*:

    xz    = W_in @ h^T                                  bimamba.py:192-196
    out   = inner(xz,          conv1d,   x_proj,   dt_proj,   A,   D)      :223-235
    out_b = inner(xz.flip(-1), conv1d_b, x_proj_b, dt_proj_b, A_b, D_b)    :236-248
    y     = 0.5*out + 0.5*out_b.flip(-1)   (or the plain sum)               :250-253
    out_proj(y^T)

``inner`` = conv+SiLU -> x_proj GEMM -> dt_proj GEMM -> B,C slices -> scan with softplus, D, z.
Weights are cast to the activation dtype when ``autocast_dtype`` is given, mirroring
selective_scan_interface.py:174-176.
"""
import torch
import torch.nn.functional as F

from .conv_ref import causal_conv1d_oracle
from .scan_ref import selective_scan_oracle


def mamba_inner_oracle(xz, conv_w, conv_b, x_proj_w, dt_proj_w, A, D, dt_bias,
                       compute_dtype=torch.float32):
    """xz: (B, 2*Dn, L) -> out_z (B, Dn, L).  conv_w: (Dn, 1, W) as stored in nn.Conv1d."""
    Bt, twoD, L = xz.shape
    Dn = twoD // 2
    R = dt_proj_w.shape[1]
    N = A.shape[1]
    x, z = xz[:, :Dn], xz[:, Dn:]
    u = causal_conv1d_oracle(x, conv_w[:, 0, :], conv_b, activation="silu", compute_dtype=compute_dtype)
    # selective_scan_interface.py:186  x_dbl = F.linear(rearrange(conv1d_out, 'b d l -> (b l) d'), x_proj_w)
    x_dbl = F.linear(u.transpose(1, 2).reshape(Bt * L, Dn), x_proj_w.to(u.dtype))
    # :187  delta = rearrange(dt_proj_w @ x_dbl[:, :R].t(), "d (b l) -> b d l")
    delta = (dt_proj_w.to(u.dtype) @ x_dbl[:, :R].t()).reshape(Dn, Bt, L).transpose(0, 1)
    Bm = x_dbl[:, R:R + N].reshape(Bt, L, N).transpose(1, 2).contiguous()      # :193-201
    Cm = x_dbl[:, R + N:].reshape(Bt, L, N).transpose(1, 2).contiguous()       # :203-212
    return selective_scan_oracle(u, delta, A, Bm, Cm, D, z=z, delta_bias=dt_bias, delta_softplus=True,
                                 compute_dtype=compute_dtype)


def bimamba_v2_oracle(hidden, p, if_devide_out=True, compute_dtype=torch.float32):
    """hidden: (B, L, d_model).  ``p``: dict of tensors keyed like the module's state_dict
    (in_proj.weight, conv1d.weight, conv1d.bias, x_proj.weight, dt_proj.weight, dt_proj.bias,
    A_log, D, conv1d_b.*, x_proj_b.weight, dt_proj_b.*, A_b_log, D_b, out_proj.weight)."""
    Bt, L, d = hidden.shape
    act = hidden.dtype
    W_in = p["in_proj.weight"].to(act)
    xz = (W_in @ hidden.reshape(Bt * L, d).t()).reshape(-1, Bt, L).transpose(0, 1)
    A = -torch.exp(p["A_log"].float())
    A_b = -torch.exp(p["A_b_log"].float())
    out = mamba_inner_oracle(xz, p["conv1d.weight"], p["conv1d.bias"], p["x_proj.weight"],
                             p["dt_proj.weight"], A, p["D"].float(), p["dt_proj.bias"].float(),
                             compute_dtype)
    out_b = mamba_inner_oracle(xz.flip(-1), p["conv1d_b.weight"], p["conv1d_b.bias"], p["x_proj_b.weight"],
                               p["dt_proj_b.weight"], A_b, p["D_b"].float(), p["dt_proj_b.bias"].float(),
                               compute_dtype)
    if if_devide_out:
        y = 0.5 * out + 0.5 * out_b.flip(-1)
    else:
        y = out + out_b.flip(-1)
    return F.linear(y.transpose(1, 2), p["out_proj.weight"].to(act), None)
