#!/usr/bin/env python
"""Launches the fused add+dropout+LayerNorm and GELU+dropout kernels at the ConMamba-large layer shape (32064 rows x 256 /
1024 columns), a few times each, for an `ncu -k regex:"add_ln|gelu_dropout"` capture (run on the B200 box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mamba_asr_b200 import kernels as K  # noqa: E402

dev = "cuda"
rows, d, dff = 64 * 501, 256, 1024
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
for it in range(3):
    flush.zero_()
    s, y, mean, rstd, mask = K.add_ln_forward(a, b, w, bb, 1e-5, 0.5, 0.1, seed, it, torch.bfloat16)
    flush.zero_()
    K.add_ln_backward(s, dy, ds, w, mean, rstd, mask, 0.5, 0.1, torch.bfloat16)
    flush.zero_()
    yy, m2 = K.gelu_dropout_forward(x, 0.1, seed, it)
    flush.zero_()
    K.gelu_dropout_backward(x, dyf, m2, 0.1)
torch.cuda.synchronize()
print("done")
