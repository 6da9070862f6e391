"""Shared helpers for the parity tests."""
import json
import math
import os

import numpy as np
import torch

# BASELINE.json north_star tolerances (the reference ships no tests of its own):
#   fp32: rtol 1e-4 ; bf16: rtol 2e-2 ; lengths / masks bit-exact.
# An element-wise rtol needs a floor for values near zero; the floor is taken relative to the tensor's own scale.
RTOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2, torch.float16: 4e-3}


def assert_close(a, ref, dtype=torch.float32, floor="rms", what="", rtol_mul=1.0):
    """|a - ref| <= rtol * (|ref| + scale(ref)) + tiny, scale = rms(ref) (elementwise maps) or max|ref| (sums).
    `tiny` (1e-8 at fp32) only matters for tensors that are identically zero in the reference."""
    a = a.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert a.shape == ref.shape, (what, a.shape, ref.shape)
    rtol = RTOL[dtype] * rtol_mul
    scale = math.sqrt(float((ref ** 2).mean())) if floor == "rms" else float(ref.abs().max())
    err = (a - ref).abs()
    tol = rtol * (ref.abs() + scale) + rtol * 1e-4
    bad = err > tol
    if bool(bad.any()):
        i = int(torch.argmax(err - tol))
        raise AssertionError("%s: %d / %d elements out of tolerance (rtol %.1e, floor %s=%.3e); worst |err| %.3e at flat "
                             "index %d (ref %.6e got %.6e)" % (what, int(bad.sum()), bad.numel(), rtol, floor, scale,
                                                             float(err.flatten()[i]), i, float(ref.flatten()[i]),
                                                             float(a.flatten()[i])))


def load_scan_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"scan_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    dt = torch.bfloat16 if "bfloat16" in meta["dtype"] else torch.float32
    ins = {}
    for k in ("u", "delta", "A", "B", "C", "D", "z", "delta_bias"):
        if "in_" + k in z:
            t = torch.from_numpy(z["in_" + k])
            if k in ("u", "delta", "z") or (k in ("B", "C") and t.dim() >= 3):
                t = t.to(dt)
            ins[k] = t
        else:
            ins[k] = None
    return z, meta, ins, dt


def make_scan_inputs(Bt, D, L, N, dtype, seed=0, a_init="xavier", device="cpu"):
    """Synthetic scan inputs as SURVEY.md section 8d prescribes (dt_bias as bimamba.py:111-118)."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    u = rn(Bt, D, L).to(dtype)
    delta = (0.5 * rn(Bt, D, L)).to(dtype)
    if a_init == "s4d":
        A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(D, 1)
    else:
        A = -torch.exp(rn(D, N) * math.sqrt(2.0 / (D + N)))
    Bm, Cm = rn(Bt, N, L).to(dtype), rn(Bt, N, L).to(dtype)
    Dp = torch.ones(D) + 0.1 * rn(D)
    z = rn(Bt, D, L).to(dtype)
    dt = torch.exp(torch.rand(D, generator=g) * (math.log(0.1) - math.log(1e-3)) + math.log(1e-3))
    bias = dt + torch.log(-torch.expm1(-dt))
    d = dict(u=u, delta=delta, A=A, B=Bm, C=Cm, D=Dp, z=z, delta_bias=bias)
    return {k: v.to(device) for k, v in d.items()}


def channel_last(t):
    """Same logical (B, D, L) tensor, (B, L, D) memory."""
    return t.transpose(1, 2).contiguous().transpose(1, 2)
