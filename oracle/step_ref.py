"""CPU restatement of the reference's incremental-decoding path (TEST INFRASTRUCTURE - see oracle/__init__.py).

  ``ssm_step_oracle``     the torch fallback for ``selective_state_update``     modules/mamba/bimamba.py:350-357
  ``mamba_step_oracle``   ``Mamba.step``                                          modules/mamba/bimamba.py:320-365
  ``mamba_prefill_oracle`` ``Mamba.forward`` with ``inference_params`` at offset 0 modules/mamba/bimamba.py:176-186, 274-316

Pinned by ``tests/golden/mamba_step.npz`` (frozen from the reference's own ``Mamba`` by ``make_golden.py --only-step``).
"""
import torch
import torch.nn.functional as F

from .conv_ref import causal_conv1d_update_oracle
from .scan_ref import selective_scan_oracle


def ssm_step_oracle(ssm_state, x, dt, A, Bm, Cm, D=None, z=None, dt_bias=None, dt_softplus=False):
    """bimamba.py:350-357.  ssm_state (B, D, N) is updated in place; returns y (B, D)."""
    dtype = x.dtype
    if dt_bias is not None:
        dt = dt + dt_bias.to(dtype=dt.dtype)
    if dt_softplus:
        dt = F.softplus(dt)
    dA = torch.exp(torch.einsum("bd,dn->bdn", dt, A))
    dB = torch.einsum("bd,bn->bdn", dt, Bm)
    ssm_state.copy_(ssm_state * dA + x.unsqueeze(-1) * dB)
    y = torch.einsum("bdn,bn->bd", ssm_state.to(dtype), Cm)
    if D is not None:
        y = y + D.to(dtype) * x
    if z is not None:
        y = y * F.silu(z)
    return y


def mamba_step_oracle(hidden, conv_state, ssm_state, p, d_state=16):
    """bimamba.py:320-365 with the forward-direction parameters ``p`` (a state_dict); hidden (B, 1, d_model)."""
    xz = F.linear(hidden.squeeze(1), p["in_proj.weight"], p.get("in_proj.bias"))
    x, z = xz.chunk(2, dim=-1)
    x = causal_conv1d_update_oracle(x, conv_state, p["conv1d.weight"][:, 0, :], p.get("conv1d.bias"), "silu")
    x_db = F.linear(x, p["x_proj.weight"])
    R = p["dt_proj.weight"].shape[1]
    dt, Bm, Cm = torch.split(x_db, [R, d_state, d_state], dim=-1)
    dt = F.linear(dt, p["dt_proj.weight"])
    A = -torch.exp(p["A_log"].float())
    y = ssm_step_oracle(ssm_state, x, dt, A, Bm, Cm, p["D"], z=z, dt_bias=p["dt_proj.bias"], dt_softplus=True)
    return F.linear(y, p["out_proj.weight"], p.get("out_proj.bias")).unsqueeze(1)


def mamba_prefill_oracle(hidden, conv_state, ssm_state, p, d_state=16):
    """bimamba.py:187-196, 274-316: the slow path of ``forward`` that fills the cache; hidden (B, L, d_model)."""
    Bt, L, _ = hidden.shape
    xz = F.linear(hidden, p["in_proj.weight"], p.get("in_proj.bias")).transpose(1, 2)      # (B, 2D, L)
    x, z = xz.chunk(2, dim=1)
    W = p["conv1d.weight"].shape[-1]
    conv_state.copy_(F.pad(x, (W - L, 0)))
    Dn = x.shape[1]
    x = F.silu(F.conv1d(x, p["conv1d.weight"], p.get("conv1d.bias"), padding=W - 1, groups=Dn)[..., :L])
    x_dbl = F.linear(x.transpose(1, 2).reshape(Bt * L, Dn), p["x_proj.weight"])
    R = p["dt_proj.weight"].shape[1]
    dt, Bm, Cm = torch.split(x_dbl, [R, d_state, d_state], dim=-1)
    dt = (p["dt_proj.weight"] @ dt.t()).reshape(Dn, Bt, L).transpose(0, 1)
    Bm = Bm.reshape(Bt, L, d_state).transpose(1, 2).contiguous()
    Cm = Cm.reshape(Bt, L, d_state).transpose(1, 2).contiguous()
    y, last = selective_scan_oracle(x, dt, -torch.exp(p["A_log"].float()), Bm, Cm, p["D"].float(), z=z,
                                    delta_bias=p["dt_proj.bias"].float(), delta_softplus=True, return_last_state=True)
    ssm_state.copy_(last)
    return F.linear(y.transpose(1, 2), p["out_proj.weight"], p.get("out_proj.bias"))
