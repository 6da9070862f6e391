#!/usr/bin/env python
"""Times the fused elementwise kernels of a ConMamba layer alone at the ConMamba-large shape (32064 rows x 256 / 1024 columns):
CUDA events around the launch, L2 flushed before every launch, best of N; algorithmic bytes against the measured HBM peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_asr_b200 import kernels as K  # noqa: E402

dev = "cuda"
rows, d, dff = 64 * 501, 256, 1024
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    peak = 6650.0
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randn(rows, d, device=dev, generator=g)
b = torch.randn(rows, d, device=dev, generator=g).bfloat16()
w, bb = torch.ones(d, device=dev), torch.zeros(d, device=dev)
seed = torch.zeros(1, dtype=torch.int64, device=dev)
x = torch.randn(rows, dff, device=dev, generator=g).bfloat16()
dy = torch.randn(rows, d, device=dev, generator=g).bfloat16()
ds = torch.randn(rows, d, device=dev, generator=g)
dyf = torch.randn(rows, dff, device=dev, generator=g).bfloat16()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
s, y, mean, rstd, mask = K.add_ln_forward(a, b, w, bb, 1e-5, 0.5, 0.1, seed, 0, torch.bfloat16)
yy, m2 = K.gelu_dropout_forward(x, 0.1, seed, 0)
xb = torch.randn(rows, d, device=dev, generator=g).bfloat16()
yl, ml, rl = K.layernorm_forward(xb, w, bb, 1e-5, torch.bfloat16)


def timeit(name, fn, nbytes, iters=12):
    best = 1e9
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%-28s best %.4f ms  alg %6.1f GB/s (%.1f%% of %.0f)" % (name, best, nbytes / best / 1e6, 100 * nbytes / best / 1e6 / peak, peak))


n, nf = rows * d, rows * dff
timeit("add_ln_fwd 32064x256", lambda: K.add_ln_forward(a, b, w, bb, 1e-5, 0.5, 0.1, seed, 1, torch.bfloat16), n * 12)
timeit("add_ln_bwd 32064x256", lambda: K.add_ln_backward(s, dy, ds, w, mean, rstd, mask, 0.5, 0.1, torch.bfloat16), n * 16)
timeit("add_ln_bwd +dbsum", lambda: K.add_ln_backward(s, dy, ds, w, mean, rstd, mask, 0.5, 0.1, torch.bfloat16, need_dbsum=True), n * 16)
_s2, _y2, _m2, _r2, mask_b = K.add_ln_forward(a, b, w, bb, 1e-5, 0.5, 0.1, seed, 0, torch.bfloat16, store_mask=True)
timeit("add_ln_fwd stored mask", lambda: K.add_ln_forward(a, b, w, bb, 1e-5, 0.5, 0.1, seed, 1, torch.bfloat16, store_mask=True), n * 13)
timeit("add_ln_bwd stored mask", lambda: K.add_ln_backward(s, dy, ds, w, mean, rstd, mask_b, 0.5, 0.1, torch.bfloat16), n * 17)
timeit("gelu_dropout_fwd 32064x1024", lambda: K.gelu_dropout_forward(x, 0.1, seed, 1), nf * 5)
timeit("gelu_dropout_bwd 32064x1024", lambda: K.gelu_dropout_backward(x, dyf, m2, 0.1), nf * 6)
timeit("gelu_dropout_bwd +colsum", lambda: K.gelu_dropout_backward(x, dyf, m2, 0.1, colsum_cols=dff), nf * 6)
timeit("layernorm_fwd 32064x256", lambda: K.layernorm_forward(xb, w, bb, 1e-5, torch.bfloat16), n * 4)
timeit("layernorm_bwd 32064x256", lambda: K.layernorm_backward(xb, dy, w, ml, rl), n * 6)
timeit("colsum 32064x1024", lambda: K.colsum(dyf), nf * 2)
timeit("colsum 32064x256", lambda: K.colsum(dy), n * 2)
