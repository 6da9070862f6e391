#!/usr/bin/env python
"""Kernel-level microbenchmark (run on the B200 box): times cm_scan_fwd / cm_scan_bwd / cm_conv_fwd / cm_conv_bwd
standalone at the BASELINE.json shapes with CUDA events and prints achieved algorithmic GB/s vs the measured HBM peak.

    python tools/prof_kernels.py [--cfg 2|3|4|5] [--iters 20] [--lanes 0] [--only scan_fwd,...] [--dtype bf16|f32]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_asr_b200 import kernels as K  # noqa: E402

CFG = {  # (Bt, D, L, R)
    "1": (8, 288, 251, 9), "2": (32, 288, 376, 9), "3": (64, 512, 501, 16), "4": (64, 1024, 501, 32),
    "5": (4, 512, 7501, 16), "5_1k": (4, 512, 1024, 16), "5_2k": (4, 512, 2048, 16), "5_4k": (4, 512, 4096, 16),
    "5_8k": (4, 512, 8192, 16), "5_16k": (4, 512, 16384, 16), "5_30k": (4, 512, 30001, 16),
}


def hbm_peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()                       # > L2 capacity: evicts the previous iteration's lines
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def timek(fn, name, iters, flush):
    """kernel-only time: event pair around the C-ABI launch (kernels._call), L2 flushed before every launch"""
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        K.start_timing()
        fn()
        ts += K.stop_timing()[name]
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="2,3")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--lanes", default="0")
    ap.add_argument("--only", default="scan_fwd,scan_bwd,conv_fwd,conv_bwd,scan_fwd_infer")
    ap.add_argument("--dtype", default="bf16")
    args = ap.parse_args()
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    s = 2 if dt == torch.bfloat16 else 4
    peak = hbm_peak()
    dev = "cuda"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    only = set(args.only.split(","))
    N = 16
    for cfg in args.cfg.split(","):
        Bt, D, L, R = CFG[cfg]
        pos = Bt * D * L
        g = torch.Generator(device=dev).manual_seed(0)
        rn = lambda *sh: torch.randn(*sh, device=dev, generator=g)
        cl = lambda: rn(Bt, L, D).to(dt).transpose(1, 2)
        z = cl()
        P = 2 * N + (R + 7) // 8 * 8                  # the module's aligned x_dbl layout: [B | C | dt | pad]
        dirs = []
        for rev in (False, True):
            xdbl = rn(Bt, L, P).to(dt)
            dirs.append(dict(u=cl(), delta=(0.5 * rn(Bt, L, D)).to(dt).transpose(1, 2), A=-torch.exp(0.3 * rn(D, N)),
                             B=xdbl[..., :N].transpose(1, 2), C=xdbl[..., N:2 * N].transpose(1, 2),
                             D=torch.ones(D, device=dev), delta_bias=torch.full((D,), -4.0, device=dev), reverse=rev))
        for lanes in [int(x) for x in args.lanes.split(",")]:
            tag = "cfg%s B%d D%d L%d %s lanes=%d" % (cfg, Bt, D, L, args.dtype, lanes)
            if "scan_fwd" in only:
                f = lambda: K.scan_forward(dirs, z=z, out_scale=0.5, delta_softplus=True, need_ckpt=True,
                                           need_out_pre=True, lanes=lanes)
                best, med = timek(f, "cm_scan_fwd", args.iters, flush)
                byts = (6 + 4.0 * N / D) * s * pos
                print("%-40s scan_fwd(train)  best %.3f ms med %.3f ms  alg %.1f GB/s (%.1f%% of %.0f)  %.1f ps/pos"
                      % (tag, best, med, byts / best / 1e6, 100 * byts / best / 1e6 / peak, peak, best * 1e9 / pos))
            if "scan_fwd_infer" in only:
                f = lambda: K.scan_forward(dirs, z=z, out_scale=0.5, delta_softplus=True, lanes=lanes)
                best, med = timek(f, "cm_scan_fwd", args.iters, flush)
                byts = (6 + 4.0 * N / D) * s * pos
                print("%-40s scan_fwd(infer)  best %.3f ms med %.3f ms  alg %.1f GB/s (%.1f%%)"
                      % (tag, best, med, byts / best / 1e6, 100 * byts / best / 1e6 / peak))
            if "scan_bwd" in only:
                res = K.scan_forward(dirs, z=z, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
                dout = cl()
                f = lambda: K.scan_backward(dirs, res["ckpt"], dout, z=z, out_pre=res["out_pre"], out_scale=0.5,
                                            delta_softplus=True, lanes=lanes)
                K.start_timing()
                for _ in range(5):
                    f()
                kt = K.stop_timing()
                best = min(kt["cm_scan_bwd"])
                byts = (11 + 8.0 * N / D) * s * pos
                print("%-40s scan_bwd kernel  best %.3f ms              alg %.1f GB/s (%.1f%%)  %.1f ps/pos ; reducers %s"
                      % (tag, best, byts / best / 1e6, 100 * byts / best / 1e6 / peak, best * 1e9 / pos,
                         {k: round(min(v), 4) for k, v in kt.items() if k != "cm_scan_bwd"}))
        tag = "cfg%s B%d D%d L%d %s" % (cfg, Bt, D, L, args.dtype)
        if "ln_act" in only:
            # wide-row LayerNorm + LeakyReLU at the two conv blocks of the CNN front-end of this config
            T = {2: 1501, 3: 2001, 4: 2001}.get(int(cfg), 4 * L)
            T1 = (T - 1) // 2 + 1
            for rows_, cols_ in ((Bt * T1, 40 * 64), (Bt * L, 20 * 32)):
                xx = rn(rows_, cols_).to(dt)
                w, bb = torch.ones(cols_, device=dev), torch.zeros(cols_, device=dev)
                best, med = timek(lambda: K.ln_act_forward(xx, w, bb, 1e-5, 0.01), "cm_ln_act_fwd", args.iters, flush)
                byts = rows_ * cols_ * 2 * s
                print("%-40s ln_act_fwd %dx%d  best %.4f ms med %.4f  alg %.1f GB/s (%.1f%%)"
                      % (tag, rows_, cols_, best, med, byts / best / 1e6, 100 * byts / best / 1e6 / peak))
                y, mean, rstd = K.ln_act_forward(xx, w, bb, 1e-5, 0.01)
                dy = rn(rows_, cols_).to(dt)
                best, med = timek(lambda: K.ln_act_backward(xx, dy, w, bb, mean, rstd, 0.01), "cm_ln_act_bwd", args.iters, flush)
                byts = rows_ * cols_ * 3 * s
                print("%-40s ln_act_bwd %dx%d  best %.4f ms med %.4f  alg %.1f GB/s (%.1f%%)"
                      % (tag, rows_, cols_, best, med, byts / best / 1e6, 100 * byts / best / 1e6 / peak))
        if "aux" in only:
            # LayerNorm / depthwise conv / column sums at the layer's (rows, d_model) shape
            d_model = D // 2
            rows = Bt * L
            for xdt, ydt in ((torch.float32, dt), (dt, dt)):
                xs_, ys_ = (4 if xdt == torch.float32 else 2), (4 if ydt == torch.float32 else 2)
                xx = rn(rows, d_model).to(xdt)
                w, bb = torch.ones(d_model, device=dev), torch.zeros(d_model, device=dev)
                best, med = timek(lambda: K.layernorm_forward(xx, w, bb, 1e-5, ydt), "cm_layernorm_fwd", args.iters, flush)
                byts = rows * d_model * (xs_ + ys_)
                print("%-40s ln_fwd %s->%s  best %.4f ms  alg %.1f GB/s (%.1f%%)" % (tag, xdt, ydt, best, byts / best / 1e6, 100 * byts / best / 1e6 / peak))
                y, mean, rstd = K.layernorm_forward(xx, w, bb, 1e-5, ydt)
                dy = rn(rows, d_model).to(ydt)
                best, med = timek(lambda: K.layernorm_backward(xx, dy, w, mean, rstd), "cm_layernorm_bwd", args.iters, flush)
                byts = rows * d_model * (2 * xs_ + ys_)
                print("%-40s ln_bwd %s  best %.4f ms  alg %.1f GB/s (%.1f%%)" % (tag, xdt, best, byts / best / 1e6, 100 * byts / best / 1e6 / peak))
            # fused residual add + dropout + LayerNorm (fp32 residual, bf16 branch, bf16 out)
            a32, bb16 = rn(rows, d_model), rn(rows, d_model).to(dt)
            seed = torch.zeros(1, dtype=torch.int64, device=dev)
            w, bb = torch.ones(d_model, device=dev), torch.zeros(d_model, device=dev)
            best, med = timek(lambda: K.add_ln_forward(a32, bb16, w, bb, 1e-5, 0.5, 0.1, seed, 1, dt), "cm_add_ln_fwd", args.iters, flush)
            byts = rows * d_model * (4 + s + 4 + s + 1)
            print("%-40s add_ln_fwd      best %.4f ms  alg %.1f GB/s (%.1f%%)" % (tag, best, byts / best / 1e6, 100 * byts / best / 1e6 / peak))
            sv, yv, mean, rstd, mask = K.add_ln_forward(a32, bb16, w, bb, 1e-5, 0.5, 0.1, seed, 1, dt)
            dyv, dsv = rn(rows, d_model).to(dt), rn(rows, d_model)
            best, med = timek(lambda: K.add_ln_backward(sv, dyv, dsv, w, mean, rstd, mask, 0.5, 0.1, dt), "cm_add_ln_bwd", args.iters, flush)
            byts = rows * d_model * (4 + s + 4 + 1 + 4 + s)
            print("%-40s add_ln_bwd      best %.4f ms  alg %.1f GB/s (%.1f%%)" % (tag, best, byts / best / 1e6, 100 * byts / best / 1e6 / peak))
            # single-token decoding at the decoder width of this config (state (B, D, 16) fp32)
            st_ = rn(Bt, D, N)
            xs1, dts1, zs1 = rn(Bt, D).to(dt), rn(Bt, D).to(dt), rn(Bt, D).to(dt)
            Bs1, Cs1 = rn(Bt, N).to(dt), rn(Bt, N).to(dt)
            As1 = -torch.exp(0.3 * rn(D, N))
            Ds1, bs1 = torch.ones(D, device=dev), torch.full((D,), -4.0, device=dev)
            best, med = timek(lambda: K.ssm_step(st_, xs1, dts1, As1, Bs1, Cs1, Ds1, z=zs1, dt_bias=bs1, dt_softplus=True),
                              "cm_ssm_step", args.iters, flush)
            byts = 2 * Bt * D * N * 4
            print("%-40s ssm_step (B=%d, D=%d) best %.4f ms  alg %.1f GB/s (%.1f%%)" % (tag, Bt, D, best, byts / best / 1e6, 100 * byts / best / 1e6 / peak))
            # tall-skinny weight-gradient GEMMs of x_proj / dt_proj: A^T B over batch * L rows
            if dt != torch.float32:
                Rp = (R + 7) // 8 * 8
                ta, tb1, tb2 = rn(rows, D).to(dt), rn(rows, Rp).to(dt), rn(rows, 2 * N + Rp).to(dt)
                for nm, tb in (("dt_proj", tb1), ("x_proj", tb2)):
                    best, med = timek(lambda: K.tsmm(ta, tb), "cm_tsmm", args.iters, flush)
                    byts = rows * (D + tb.shape[1]) * s
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize(); e0.record()
                    for _ in range(20):
                        torch.mm(ta.t(), tb, out_dtype=torch.float32)
                    e1.record(); torch.cuda.synchronize()
                    print("%-40s tsmm %s (%dx%dx%d) best %.4f ms  alg %.1f GB/s (%.1f%%) ; torch.mm warm %.4f ms"
                          % (tag, nm, rows, D, tb.shape[1], best, byts / best / 1e6, 100 * byts / best / 1e6 / peak, e0.elapsed_time(e1) / 20))
            xc = rn(Bt, L, d_model).to(dt)
            wc, bc_ = rn(d_model, 31), rn(d_model)
            best, med = timek(lambda: K.dwconv_forward(xc, wc, bc_, 15), "cm_dwconv_fwd", args.iters, flush)
            byts = 2 * s * rows * d_model
            print("%-40s dwconv_fwd k31  best %.4f ms  alg %.1f GB/s (%.1f%%)" % (tag, best, byts / best / 1e6, 100 * byts / best / 1e6 / peak))
            dyc = rn(Bt, L, d_model).to(dt)
            best, med = timek(lambda: K.dwconv_backward_weight(xc, dyc, 31, 15), "cm_dwconv_bwd_weight", args.iters, flush)
            print("%-40s dwconv_bwdw k31 best %.4f ms  alg %.1f GB/s (%.1f%%)" % (tag, best, byts / best / 1e6, 100 * byts / best / 1e6 / peak))
            dff = rn(rows, 4 * d_model).to(dt)
            best, med = timek(lambda: K.colsum(dff), "cm_colsum", args.iters, flush)
            byts = s * rows * 4 * d_model
            print("%-40s colsum 4d       best %.4f ms  alg %.1f GB/s (%.1f%%)" % (tag, best, byts / best / 1e6, 100 * byts / best / 1e6 / peak))
        x = cl()
        cdirs = [dict(weight=rn(D, 4), bias=rn(D), anticausal=False), dict(weight=rn(D, 4), bias=rn(D), anticausal=True)]
        tag = "cfg%s B%d D%d L%d %s" % (cfg, Bt, D, L, args.dtype)
        if "conv_fwd" in only:
            outs = [K.empty_like_bdl(x), K.empty_like_bdl(x)]
            best, med = timeit(lambda: K.conv_forward(x, cdirs, silu=True, outs=outs), args.iters, flush)
            byts = 3 * s * pos
            print("%-40s conv_fwd         best %.3f ms med %.3f ms  alg %.1f GB/s (%.1f%%)"
                  % (tag, best, med, byts / best / 1e6, 100 * byts / best / 1e6 / peak))
        if "conv_bwd" in only:
            douts = [cl(), cl()]
            K.start_timing()
            for _ in range(5):
                K.conv_backward(x, cdirs, douts, silu=True)
            kt = K.stop_timing()
            best = min(kt["cm_conv_bwd"])
            byts = 4 * s * pos
            print("%-40s conv_bwd kernel  best %.3f ms              alg %.1f GB/s (%.1f%%)"
                  % (tag, best, byts / best / 1e6, 100 * byts / best / 1e6 / peak))


if __name__ == "__main__":
    main()
