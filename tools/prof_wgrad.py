#!/usr/bin/env python
"""Weight-gradient GEMM variants of the Linear layers around the Mamba block (dW = dy^T x, K = batch * L rows) timed on
the B200 box: which cuBLAS formulation should linear._LinearFn.backward use?

    python tools/prof_wgrad.py [--iters 30]
"""
import argparse

import torch


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
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
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    args = ap.parse_args()
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    f32 = torch.float32
    shapes = [("small ffn1", 12032, 1024, 144), ("small ffn2", 12032, 144, 1024), ("small in_proj", 12032, 576, 144),
              ("small out_proj", 12032, 144, 288), ("small pw", 12032, 288, 144), ("small lin", 12032, 144, 144),
              ("large ffn1", 32064, 1024, 256), ("large ffn2", 32064, 256, 1024), ("large in_proj", 32064, 1024, 256),
              ("large out_proj", 32064, 256, 512), ("large pw", 32064, 512, 256), ("large lin", 32064, 256, 256)]
    for name, rows, M, N in shapes:
        dy = torch.randn(rows, M, device=dev).bfloat16()
        x = torch.randn(rows, N, device=dev).bfloat16()
        ref = dy.float().t() @ x.float()
        variants = {
            "mm(dy^T,x) f32out [current]": lambda: torch.mm(dy.t(), x, out_dtype=f32),
            "mm(x^T,dy)^T f32out": lambda: torch.mm(x.t(), dy, out_dtype=f32).t(),
            "mm(dy^T,x) bf16out": lambda: torch.mm(dy.t(), x),
        }
        for ns in (4, 8, 16):
            if rows % ns == 0:
                variants["bmm split %d + sum" % ns] = (lambda ns=ns: torch.bmm(
                    dy.unflatten(0, (ns, rows // ns)).transpose(1, 2), x.unflatten(0, (ns, rows // ns)), out_dtype=f32).sum(0))
        variants["dy^T.contiguous() then mm"] = lambda: torch.mm(dy.t().contiguous(), x, out_dtype=f32)
        out = []
        for vn, fn in variants.items():
            t = timeit(fn, args.iters, flush)
            err = float((fn().float() - ref).abs().max() / ref.abs().max())
            out.append((t, vn, err))
        base = out[0][0]
        print("%-16s rows %5d  dW %4d x %4d  (%.1f GFLOP)" % (name, rows, M, N, 2e-9 * rows * M * N))
        for t, vn, err in out:
            print("    %-32s %7.1f us  x%.2f  relerr %.1e" % (vn, t * 1e3, base / t, err))


if __name__ == "__main__":
    main()
