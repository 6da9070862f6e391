#!/usr/bin/env python
"""GPU diagnostics (run on the B200 box): prints max errors of every kernel against the CPU oracle for a sweep of
shapes / layouts / dtypes.  Not a test - it never fails on numerics, it informs tolerances and debugging."""
import math
import os
import sys
import time
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_asr_b200 import kernels as K                                     # noqa: E402
from mamba_asr_b200.selective_scan_interface import selective_scan_fn       # noqa: E402
from mamba_asr_b200.causal_conv1d import causal_conv1d_fn                   # noqa: E402
from oracle.scan_ref import selective_scan_oracle                           # noqa: E402
from oracle.conv_ref import causal_conv1d_oracle                            # noqa: E402

dev = "cuda"


def errs(a, ref):
    a = a.detach().double().cpu()
    ref = ref.detach().double().cpu()
    d = (a - ref).abs()
    return "max_abs %.3e  max|ref| %.3e  rel_to_max %.3e  rms_rel %.3e" % (
        d.max(), ref.abs().max(), d.max() / (ref.abs().max() + 1e-30),
        math.sqrt((d ** 2).mean()) / (math.sqrt((ref ** 2).mean()) + 1e-30))


def make_scan(Bt, D, L, N, dtype, seed=0, a_init="xavier"):
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    u = rn(Bt, D, L).to(dtype)
    delta = (0.5 * rn(Bt, D, L)).to(dtype)
    A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(D, 1) if a_init == "s4d" else \
        -torch.exp(rn(D, N) * math.sqrt(2.0 / (D + N)))
    Bm, Cm = rn(Bt, N, L).to(dtype), rn(Bt, N, L).to(dtype)
    Dp = rn(D)
    z = rn(Bt, D, L).to(dtype)
    dt = torch.exp(torch.rand(D, generator=g) * (math.log(0.1) - math.log(1e-3)) + math.log(1e-3))
    bias = dt + torch.log(-torch.expm1(-dt))
    return dict(u=u, delta=delta, A=A, B=Bm, C=Cm, D=Dp, z=z, delta_bias=bias)


def chlast(t):
    return t.transpose(1, 2).contiguous().transpose(1, 2)


def run_scan(tag, ins, layout, lanes, grad=True):
    cpu = {k: v.clone() for k, v in ins.items()}
    cu = {k: v.to(dev) for k, v in ins.items()}
    if layout == "cl":
        for k in ("u", "delta", "z", "B", "C"):
            cu[k] = chlast(cu[k])
    os.environ["CM_SCAN_LANES"] = str(lanes)
    leaf_c = {k: v.requires_grad_(grad) for k, v in cpu.items()}
    leaf_g = {k: v.requires_grad_(grad) for k, v in cu.items()}
    ref = selective_scan_oracle(leaf_c["u"], leaf_c["delta"], leaf_c["A"], leaf_c["B"], leaf_c["C"], leaf_c["D"],
                                leaf_c["z"], leaf_c["delta_bias"], True)
    ref64 = selective_scan_oracle(cpu["u"], cpu["delta"], cpu["A"], cpu["B"], cpu["C"], cpu["D"], cpu["z"],
                                  cpu["delta_bias"], True, compute_dtype=torch.float64).detach()
    d = dict(u=leaf_g["u"], delta=leaf_g["delta"], A=leaf_g["A"], B=leaf_g["B"], C=leaf_g["C"], D=leaf_g["D"],
             delta_bias=leaf_g["delta_bias"], reverse=False)
    out = selective_scan_fn(leaf_g["u"], leaf_g["delta"], leaf_g["A"], leaf_g["B"], leaf_g["C"], leaf_g["D"],
                            leaf_g["z"], leaf_g["delta_bias"], True)
    print("[scan %s %s lanes=%d] fwd vs ref32: %s" % (tag, layout, lanes, errs(out, ref)))
    print("    ours vs fp64: %s" % errs(out.float(), ref64.float()))
    print("    ref32 vs fp64: %s" % errs(ref.float(), ref64.float()))
    if grad:
        gc = torch.Generator().manual_seed(99)
        cot = torch.randn(ref.shape, generator=gc)
        (ref.float() * cot).sum().backward()
        (out.float() * cot.to(dev)).sum().backward()
        for k in ("u", "delta", "A", "B", "C", "D", "z", "delta_bias"):
            print("    grad %-10s %s" % (k, errs(leaf_g[k].grad, leaf_c[k].grad)))


def run_conv(tag, Bt, D, L, W, dtype, layout):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(Bt, D, L, generator=g).to(dtype)
    w = (torch.randn(D, W, generator=g) * 0.5)
    b = torch.randn(D, generator=g) * 0.5
    xc, wc, bc = x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    xg = x.to(dev)
    if layout == "cl":
        xg = chlast(xg)
    xg = xg.requires_grad_()
    wg, bg = w.to(dev).requires_grad_(), b.to(dev).requires_grad_()
    ref = causal_conv1d_oracle(xc, wc, bc, "silu")
    out = causal_conv1d_fn(xg, wg, bg, None, "silu")
    cot = torch.randn(ref.shape, generator=g)
    (ref.float() * cot).sum().backward()
    (out.float() * cot.to(dev)).sum().backward()
    print("[conv %s %s] fwd %s" % (tag, layout, errs(out, ref)))
    print("    dx %s\n    dw %s\n    db %s" % (errs(xg.grad, xc.grad), errs(wg.grad, wc.grad), errs(bg.grad, bc.grad)))


def guarded(f, *a, **k):
    try:
        f(*a, **k)
    except Exception:
        traceback.print_exc()
    torch.cuda.synchronize()


if __name__ == "__main__":
    torch.manual_seed(0)
    print(torch.cuda.get_device_name(0), torch.__version__)
    for layout in ("tc", "cl"):
        for lanes in (1, 2, 4):
            guarded(run_scan, "f32 B2 D64 L67", make_scan(2, 64, 67, 16, torch.float32), layout, lanes)
    guarded(run_scan, "f32 s4d B2 D40 L131", make_scan(2, 40, 131, 16, torch.float32, a_init="s4d"), "cl", 1)
    guarded(run_scan, "f32 B4 D96 L501", make_scan(4, 96, 501, 16, torch.float32, seed=3), "cl", 1)
    guarded(run_scan, "bf16 B4 D96 L501", make_scan(4, 96, 501, 16, torch.bfloat16, seed=3), "cl", 1)
    guarded(run_scan, "f16 B2 D64 L200", make_scan(2, 64, 200, 16, torch.float16, seed=4), "cl", 2)
    guarded(run_scan, "f32 L1", make_scan(2, 32, 1, 16, torch.float32, seed=5), "tc", 1)
    guarded(run_scan, "f32 N8", make_scan(2, 32, 50, 8, torch.float32, seed=6), "cl", 1)
    for layout in ("tc", "cl"):
        guarded(run_conv, "f32 W4", 2, 64, 67, 4, torch.float32, layout)
        guarded(run_conv, "f32 W3 D40", 2, 40, 130, 3, torch.float32, layout)
        guarded(run_conv, "bf16 W4", 4, 288, 376, 4, torch.bfloat16, layout)
    print("diag done")
