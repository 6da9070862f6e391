#!/usr/bin/env python
"""Tensor-pipe utilisation of the projection GEMMs around the hot path (SURVEY.md section 8(d) metric iii) - run on the B200.

The dense projections of a ConMamba layer stay on cuBLAS; this tool times each of them at a BASELINE
configuration's shapes (bf16, CUDA events, L2 flushed before every call) - forward, input gradient and weight gradient - and
prints achieved TFLOP/s against the measured dense bf16 peak (MEASURED_PEAKS.json: burst figure, a GEMM timed alone) next to
the bytes it has to move against the measured HBM peak: whichever fraction is larger is the roof that binds the call.

    python tools/prof_gemms.py [--cfg 3] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = {  # name: (rows = batch * encoder frames, d_model, d_ffn, d_inner, dt_rank, d_state)
    "2": (32 * 376, 144, 1024, 288, 9, 16),
    "3": (64 * 501, 256, 1024, 512, 16, 16),
    "4": (64 * 501, 512, 2048, 1024, 32, 16),
}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), float(p["bf16_tflops"])
    except Exception:
        return 6650.0, 1600.0


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="3")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    hbm, tf = peaks()
    dev = "cuda"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for cfg in args.cfg.split(","):
        rows, d, dff, D, R, N = CFG[cfg]
        Rp = (R + 7) // 8 * 8
        gemms = [  # name, K (in features), N (out features), per-layer count
            ("in_proj (d -> 2D)", d, 2 * D, 1), ("x_proj (D -> 2N+R), per direction", D, 2 * N + Rp, 2),
            ("dt_proj (R -> D), per direction", Rp, D, 2), ("out_proj (D -> d)", D, d, 1),
            ("FFN up (d -> d_ffn), x2 modules", d, dff, 2), ("FFN down (d_ffn -> d), x2 modules", dff, d, 2),
            ("conv-module pointwise (d -> 2d)", d, 2 * d, 1), ("conv-module linear (d -> d)", d, d, 1),
        ]
        print("cfg %s: rows = %d, bf16; peaks: %.0f TFLOP/s dense bf16 (burst), %.0f GB/s HBM" % (cfg, rows, tf, hbm))
        print("%-40s %5s %5s | %-30s | %-30s | %-30s" % ("GEMM", "K", "N", "forward  us TF/s %tc %hbm", "dgrad", "wgrad (fp32 out)"))
        tot_t, tot_f = 0.0, 0.0
        for name, Kf, Nf, cnt in gemms:
            x = torch.randn(rows, Kf, device=dev).bfloat16()
            w = torch.randn(Nf, Kf, device=dev).bfloat16()
            dy = torch.randn(rows, Nf, device=dev).bfloat16()
            flops = 2.0 * rows * Kf * Nf
            cells = []
            for what, fn, byts in (
                    ("fwd", lambda: torch.mm(x, w.t()), 2.0 * (rows * Kf + Nf * Kf + rows * Nf)),
                    ("dgrad", lambda: torch.mm(dy, w), 2.0 * (rows * Nf + Nf * Kf + rows * Kf)),
                    ("wgrad", lambda: torch.mm(dy.t(), x, out_dtype=torch.float32), 2.0 * (rows * Nf + rows * Kf) + 4.0 * Nf * Kf)):
                ms = timeit(fn, args.iters, flush)
                cells.append("%7.1f %6.0f %5.1f%% %5.1f%%" % (ms * 1e3, flops / ms / 1e9, 100 * flops / ms / 1e9 / tf,
                                                            100 * byts / ms / 1e6 / hbm))
                tot_t += cnt * ms
                tot_f += cnt * flops
            print("%-40s %5d %5d | %s | %s | %s" % (name, Kf, Nf, cells[0], cells[1], cells[2]))
        print("per layer (counts applied, single-GEMM weight gradients): %.3f ms, %.1f GFLOP -> %.0f TFLOP/s = %.1f%% of the tensor peak"
              % (tot_t, tot_f / 1e9, tot_f / tot_t / 1e9, 100 * tot_f / tot_t / 1e9 / tf))
        print()


if __name__ == "__main__":
    main()
