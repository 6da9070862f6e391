#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

It imports the reference's ``modules/mamba/selective_scan_interface.py`` and ``modules/mamba/bimamba.py``
with empty stand-in modules for the three CUDA extensions the reference imports at top level
(``causal_conv1d``, ``causal_conv1d_cuda``, ``selective_scan_cuda`` - selective_scan_interface.py:14-16),
then freezes

  scan_*.npz      inputs, ``selective_scan_ref`` outputs (selective_scan_interface.py:91-157) and, for the
                  fp32 cases, torch-autograd gradients through it for a fixed cotangent
  conv_*.npz      the reference's torch conv fallback  act(conv1d(x)[..., :L])  (bimamba.py:83-91, 278-279)
                  with autograd gradients
  bimamba_v2.npz  the v2 composition written with the reference's own pieces and flips (bimamba.py:223-253)
  mamba_step.npz  prefill + single-token ``step`` outputs and cache states of the reference ``Mamba``
                  (bimamba.py:176-186, 320-414)
  mamba_state_dict.json   parameter names / shapes / dtypes / tagged attrs of the reference
                  ``Mamba(d_model, bimamba_type="v2")`` constructor (bimamba.py:40-174)

The GPU box has no /root/reference; tests only read the committed fixtures.
"""
import json
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ("causal_conv1d", "causal_conv1d_cuda", "selective_scan_cuda"):
        m = types.ModuleType(name)
        if name == "causal_conv1d":
            m.causal_conv1d_fn = None
            m.causal_conv1d_update = None
        sys.modules[name] = m
    sys.path.insert(0, REF)
    import modules.mamba.selective_scan_interface as ssi
    import modules.mamba.bimamba as bim
    return ssi, bim


def t2n(t):
    t = t.detach()
    if t.dtype == torch.bfloat16:
        return t.float().numpy()      # bf16 values are exactly representable in fp32
    return t.numpy()


def scan_case(ssi, name, seed, Bt, Dm, L, N, dtype, bc_dim, has_D, has_z, has_bias, softplus, a_init,
              want_grad):
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    u = rn(Bt, Dm, L).to(dtype)
    delta = 0.5 * rn(Bt, Dm, L)
    if not softplus:
        delta = delta.abs()           # a raw step size must be positive for a decaying recurrence
    delta = delta.to(dtype)
    if a_init == "s4d":
        A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(Dm, 1)
    else:  # xavier-normal A_log as TransformerASR._init_params leaves it (TransformerASR.py:1051-1054)
        A = -torch.exp(rn(Dm, N) * math.sqrt(2.0 / (Dm + N)))
    if bc_dim == 2:
        Bm, Cm = rn(Dm, N), rn(Dm, N)
    elif bc_dim == 3:
        Bm, Cm = rn(Bt, N, L).to(dtype), rn(Bt, N, L).to(dtype)
    else:
        Bm, Cm = rn(Bt, 1, N, L).to(dtype), rn(Bt, 1, N, L).to(dtype)
    Dp = rn(Dm) if has_D else None
    z = rn(Bt, Dm, L).to(dtype) if has_z else None
    if has_bias:
        dt = torch.exp(torch.rand(Dm, generator=g) * (math.log(0.1) - math.log(1e-3)) + math.log(1e-3))
        bias = dt + torch.log(-torch.expm1(-dt))          # bimamba.py:111-118
    else:
        bias = None
    ins = dict(u=u, delta=delta, A=A, B=Bm, C=Cm, D=Dp, z=z, delta_bias=bias)
    leaf = {k: (v.clone().requires_grad_(True) if (want_grad and v is not None) else v)
            for k, v in ins.items()}
    out, last = ssi.selective_scan_ref(leaf["u"], leaf["delta"], leaf["A"], leaf["B"], leaf["C"], leaf["D"],
                                       leaf["z"], leaf["delta_bias"], delta_softplus=softplus,
                                       return_last_state=True)
    rec = {"in_" + k: t2n(v) for k, v in ins.items() if v is not None}
    rec["out"] = t2n(out)
    rec["last_state"] = t2n(last)
    rec["meta"] = np.array(json.dumps(dict(dtype=str(dtype), softplus=softplus, seed=seed)))
    if want_grad:
        cot = torch.randn(out.shape, generator=g)
        rec["cotangent"] = t2n(cot)
        (out.float() * cot).sum().backward()
        for k, v in leaf.items():
            if v is not None:
                rec["grad_" + k] = t2n(v.grad)
    np.savez_compressed(os.path.join(HERE, f"scan_{name}.npz"), **rec)
    print("wrote scan_%s  out|max|=%.4g" % (name, float(out.float().abs().max())))


def conv_case(name, seed, Bt, Dm, L, W, has_bias):
    g = torch.Generator().manual_seed(seed)
    conv = nn.Conv1d(Dm, Dm, bias=has_bias, kernel_size=W, groups=Dm, padding=W - 1)   # bimamba.py:83-91
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.5)
        if has_bias:
            conv.bias.copy_(torch.randn(Dm, generator=g) * 0.5)
    x = torch.randn(Bt, Dm, L, generator=g).requires_grad_(True)
    act = nn.SiLU()
    y = act(conv(x)[..., :L])                                                            # bimamba.py:278-279
    cot = torch.randn(y.shape, generator=g)
    (y * cot).sum().backward()
    rec = dict(in_x=t2n(x), in_weight=t2n(conv.weight[:, 0, :]), out=t2n(y), cotangent=t2n(cot),
               grad_x=t2n(x.grad), grad_weight=t2n(conv.weight.grad[:, 0, :]))
    if has_bias:
        rec["in_bias"] = t2n(conv.bias)
        rec["grad_bias"] = t2n(conv.bias.grad)
    np.savez_compressed(os.path.join(HERE, f"conv_{name}.npz"), **rec)
    print("wrote conv_%s" % name)


def bimamba_case(ssi, bim, seed=11, Bt=2, L=37, d_model=32):
    torch.manual_seed(seed)
    m = bim.Mamba(d_model=d_model, d_state=16, d_conv=4, expand=2, bimamba_type="v2")
    # TransformerASR._init_params (TransformerASR.py:1051-1054): xavier_normal_ on every >=2-D parameter
    for p in m.parameters():
        if p.dim() > 1:
            nn.init.xavier_normal_(p)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    meta = {k: dict(shape=list(v.shape), dtype=str(v.dtype)) for k, v in sd.items()}
    tags = {n: sorted(a for a in ("_no_weight_decay", "_no_reinit") if hasattr(p, a))
            for n, p in m.named_parameters()}
    with open(os.path.join(HERE, "mamba_state_dict.json"), "w") as f:
        json.dump(dict(d_model=d_model, params=meta, tags=tags, d_inner=m.d_inner, dt_rank=m.dt_rank), f,
                  indent=1, sort_keys=True)

    hidden = torch.randn(Bt, L, d_model)
    Dn, R, N = m.d_inner, m.dt_rank, m.d_state

    def inner(xz, conv, x_proj, dt_proj, A, Dp):
        # the body of MambaInnerFnNoOutProj.forward (selective_scan_interface.py:177-218) with the reference's
        # CPU-runnable pieces: its torch conv fallback and its selective_scan_ref
        x, z = xz.chunk(2, dim=1)
        u = m.act(conv(x)[..., :L])
        x_dbl = F.linear(u.transpose(1, 2).reshape(Bt * L, Dn), x_proj.weight)
        delta = (dt_proj.weight @ x_dbl[:, :R].t()).reshape(Dn, Bt, L).transpose(0, 1)
        Bm = x_dbl[:, R:R + N].reshape(Bt, L, N).transpose(1, 2).unsqueeze(1).contiguous()
        Cm = x_dbl[:, -N:].reshape(Bt, L, N).transpose(1, 2).unsqueeze(1).contiguous()
        return ssi.selective_scan_ref(u, delta, A, Bm, Cm, Dp.float(), z=z, delta_bias=dt_proj.bias.float(),
                                      delta_softplus=True)

    with torch.no_grad():
        xz = (m.in_proj.weight @ hidden.reshape(Bt * L, d_model).t()).reshape(2 * Dn, Bt, L).transpose(0, 1)
        A = -torch.exp(m.A_log.float())
        A_b = -torch.exp(m.A_b_log.float())
        out = inner(xz, m.conv1d, m.x_proj, m.dt_proj, A, m.D)
        out_b = inner(xz.flip([-1]), m.conv1d_b, m.x_proj_b, m.dt_proj_b, A_b, m.D_b)
        y = F.linear((0.5 * out + 0.5 * out_b.flip([-1])).transpose(1, 2), m.out_proj.weight, None)
        y_sum = F.linear((out + out_b.flip([-1])).transpose(1, 2), m.out_proj.weight, None)
    rec = {"p_" + k: t2n(v) for k, v in sd.items()}
    rec.update(hidden=t2n(hidden), out=t2n(y), out_nodivide=t2n(y_sum))
    np.savez_compressed(os.path.join(HERE, "bimamba_v2.npz"), **rec)
    print("wrote bimamba_v2, mamba_state_dict.json")


def step_case(ssi, bim, seed=21, Bt=3, L=9, T=5, d_model=32):
    """Incremental decoding through the reference's own ``Mamba`` (bimamba.py:176-186, 320-414): a prefill of L tokens
    with ``inference_params`` (the slow path, bimamba.py:274-316, with ``selective_scan_fn`` bound to the reference's
    ``selective_scan_ref`` because the CUDA extension is absent), then T single-token ``step`` calls on the torch
    fallbacks of bimamba.py:328-333 and :350-357.  Also T steps from zero states (no prefill)."""
    torch.manual_seed(seed)
    m = bim.Mamba(d_model=d_model, d_state=16, d_conv=4, expand=2, bimamba_type="v2", layer_idx=0)
    for p in m.parameters():
        if p.dim() > 1:
            nn.init.xavier_normal_(p)
    with torch.no_grad():
        m.D.copy_(torch.randn(m.D.shape))
        m.A_log.copy_(m.A_log + 0.3 * torch.randn(m.A_log.shape))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    bim.selective_scan_fn = ssi.selective_scan_ref            # same signature (selective_scan_interface.py:82,91)
    assert bim.causal_conv1d_fn is None and bim.causal_conv1d_update is None and bim.selective_state_update is None

    class InferenceParams:                                    # mamba_ssm.utils.generation.InferenceParams fields used
        def __init__(self):
            self.seqlen_offset = 0
            self.key_value_memory_dict = {}

    prompt = torch.randn(Bt, L, d_model)
    tokens = torch.randn(Bt, T, d_model)
    rec = {"p_" + k: t2n(v) for k, v in sd.items()}
    with torch.no_grad():
        ip = InferenceParams()
        out_prefill = m(prompt, inference_params=ip)
        conv0, ssm0 = [t.clone() for t in ip.key_value_memory_dict[0]]
        outs = []
        for t in range(T):
            ip.seqlen_offset = L + t
            outs.append(m(tokens[:, t:t + 1], inference_params=ip))
        conv1, ssm1 = ip.key_value_memory_dict[0]
        rec.update(prompt=t2n(prompt), tokens=t2n(tokens), out_prefill=t2n(out_prefill),
                   conv_after_prefill=t2n(conv0), ssm_after_prefill=t2n(ssm0),
                   out_steps=t2n(torch.cat(outs, dim=1)), conv_final=t2n(conv1), ssm_final=t2n(ssm1))
        # steps from a zero cache
        cs, ss = m.allocate_inference_cache(Bt, 0)
        outs = []
        for t in range(T):
            o, cs, ss = m.step(tokens[:, t:t + 1], cs, ss)
            outs.append(o)
        rec.update(out_steps_zero=t2n(torch.cat(outs, dim=1)), conv_zero_final=t2n(cs), ssm_zero_final=t2n(ss))
    np.savez_compressed(os.path.join(HERE, "mamba_step.npz"), **rec)
    print("wrote mamba_step")


def main():
    torch.set_num_threads(1)          # fixed reduction order
    ssi, bim = import_reference()
    if "--only-step" in sys.argv:
        step_case(ssi, bim)
        return
    f32, bf16 = torch.float32, torch.bfloat16
    scan_case(ssi, "f32_full", 1, 2, 32, 67, 16, f32, 3, True, True, True, True, "xavier", True)
    scan_case(ssi, "f32_s4d", 2, 2, 32, 131, 16, f32, 3, True, True, True, True, "s4d", True)
    scan_case(ssi, "f32_plain4d", 3, 1, 32, 40, 16, f32, 4, False, False, False, False, "xavier", True)
    scan_case(ssi, "f32_constBC", 4, 2, 32, 33, 16, f32, 2, True, False, True, True, "s4d", False)
    scan_case(ssi, "bf16_full", 5, 2, 64, 96, 16, bf16, 4, True, True, True, True, "xavier", False)
    scan_case(ssi, "f32_L1", 6, 1, 32, 1, 16, f32, 3, True, True, True, True, "s4d", True)
    conv_case("w4_bias", 7, 2, 32, 67, 4, True)
    conv_case("w4_nobias", 8, 1, 32, 5, 4, False)
    conv_case("w2_short", 9, 2, 32, 3, 2, True)
    bimamba_case(ssi, bim)
    step_case(ssi, bim)


if __name__ == "__main__":
    main()
