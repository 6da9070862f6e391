#!/usr/bin/env python
"""One small launch of every hand-synchronised kernel at ragged shapes, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize.py      (one tool per gpurun call)
    compute-sanitizer --tool racecheck python tools/sanitize.py

Covers the kernels that synchronise through mbarrier full/empty pairs, named barriers and __syncwarp (scan_fwd_sp,
scan_bwd_sp), the TMA kernels (scan_fwd_lc, scan_bwd_lc, scan_fwd_wg, scan_bwd_wg: async-proxy writes into shared memory, setmaxnreg
roles, tensor-core reductions), the one-kernel Fbank front-end, the conv kernels and the
fused LayerNorm + activation kernels, at L in {1, 9, 17, 131} and D in {288, 128}."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_asr_b200 import kernels as K  # noqa: E402


def scan_case(Bt, D, L, dt, env):
    for k, v in env.items():
        os.environ[k] = v
    g = torch.Generator(device="cuda").manual_seed(L)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    cl = lambda: rn(Bt, L, D).to(dt).transpose(1, 2)
    z = cl()
    dirs = []
    for rev in (False, True):
        xd = rn(Bt, L, 48).to(dt)
        dirs.append(dict(u=cl(), delta=(0.5 * rn(Bt, L, D)).to(dt).transpose(1, 2), A=-torch.exp(0.3 * rn(D, 16)),
                         B=xd[..., :16].transpose(1, 2), C=xd[..., 16:32].transpose(1, 2), D=torch.ones(D, device="cuda"),
                         delta_bias=torch.full((D,), -4.0, device="cuda"), reverse=rev))
    res = K.scan_forward(dirs, z=z, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    K.scan_backward(dirs, res["ckpt"], cl(), z=z, out_pre=res["out_pre"], out_scale=0.5, delta_softplus=True)
    K.scan_forward(dirs[:1], z=z, delta_softplus=True, need_last_state=True)
    torch.cuda.synchronize()
    for k in env:
        os.environ.pop(k, None)


def main():
    dt = torch.bfloat16
    n = 0
    for L in (1, 9, 17, 131):
        for D, env in ((288, {}), (128, {}), (128, {"CM_SCAN_LC": "1", "CM_SCAN_LC_BWD": "1"}),
                       (288, {"CM_SCAN_WG": "1", "CM_SCAN_WG_FWD": "1"}), (128, {"CM_SCAN_WG": "1", "CM_SCAN_WG_FWD": "1"})):
            scan_case(2, D, L, dt, env)
            n += 1
        x = torch.randn(2, L, 288, device="cuda").to(dt).transpose(1, 2)
        cd = [dict(weight=torch.randn(288, 4, device="cuda"), bias=torch.randn(288, device="cuda"), anticausal=r) for r in (False, True)]
        us = K.conv_forward(x, cd, silu=True)
        K.conv_backward(x, cd, [torch.ones_like(u) for u in us], silu=True)
        rows = 2 * L
        xm = torch.randn(rows, 640, device="cuda").to(dt)
        w, b = torch.ones(640, device="cuda"), torch.zeros(640, device="cuda")
        y, mean, rstd = K.ln_act_forward(xm, w, b, 1e-5)
        K.ln_act_backward(xm, torch.ones_like(y), w, b, mean, rstd)
        from mamba_asr_b200 import Fbank
        for n_fft, win in ((400, 25), (512, 25)):
            Fbank(n_fft=n_fft, n_mels=80, win_length=win).cuda()(torch.randn(2, 160 * L + 7, device="cuda"))
        torch.cuda.synchronize()
    print("sanitize.py: %d scan cases + conv + ln_act at L in (1, 9, 17, 131) completed" % n)


if __name__ == "__main__":
    main()
