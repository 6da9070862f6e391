"""CPU: the oracle restatements against the golden fixtures frozen from the reference itself
(tests/golden/make_golden.py).  fp32 scan/conv outputs must be bit-identical."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.bimamba_ref import bimamba_v2_oracle
from oracle.conv_ref import causal_conv1d_oracle
from oracle.fbank_ref import fbank_numpy_dft, fbank_oracle, mel_filterbank_matrix
from oracle.lengths_ref import (abs_lengths, abs_lengths_numpy, encoder_frames, fbank_frames,
                                key_padding_mask)
from oracle.scan_ref import selective_scan_oracle

torch.set_num_threads(1)

SCAN_CASES = ["f32_full", "f32_s4d", "f32_plain4d", "f32_constBC", "bf16_full", "f32_L1"]


def load_scan(golden_dir, name):
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
    return z, meta, ins


@pytest.mark.parametrize("name", SCAN_CASES)
def test_scan_oracle_matches_reference_bitwise(golden_dir, name):
    z, meta, ins = load_scan(golden_dir, name)
    out, last = selective_scan_oracle(ins["u"], ins["delta"], ins["A"], ins["B"], ins["C"], ins["D"], ins["z"],
                                      ins["delta_bias"], delta_softplus=meta["softplus"], return_last_state=True)
    assert np.array_equal(out.float().numpy(), z["out"])
    assert np.array_equal(last.numpy(), z["last_state"])


@pytest.mark.parametrize("name", ["f32_full", "f32_s4d", "f32_plain4d", "f32_L1"])
def test_scan_oracle_autograd_matches_reference(golden_dir, name):
    z, meta, ins = load_scan(golden_dir, name)
    leaf = {k: (v.clone().requires_grad_(True) if v is not None else None) for k, v in ins.items()}
    out = selective_scan_oracle(leaf["u"], leaf["delta"], leaf["A"], leaf["B"], leaf["C"], leaf["D"], leaf["z"],
                                leaf["delta_bias"], delta_softplus=meta["softplus"])
    (out * torch.from_numpy(z["cotangent"])).sum().backward()
    for k, v in leaf.items():
        if v is not None:
            np.testing.assert_allclose(v.grad.numpy(), z["grad_" + k], rtol=1e-6, atol=1e-6)


def test_scan_oracle_reverse_equals_flip(golden_dir):
    _, meta, ins = load_scan(golden_dir, "f32_full")
    f = lambda t: None if t is None else t.flip(-1)
    a = selective_scan_oracle(ins["u"], ins["delta"], ins["A"], ins["B"], ins["C"], ins["D"], ins["z"],
                              ins["delta_bias"], delta_softplus=True, reverse=True)
    b = selective_scan_oracle(f(ins["u"]), f(ins["delta"]), ins["A"], f(ins["B"]), f(ins["C"]), ins["D"],
                              f(ins["z"]), ins["delta_bias"], delta_softplus=True).flip(-1)
    assert torch.equal(a, b)


def test_scan_oracle_fp64_truth_close(golden_dir):
    z, meta, ins = load_scan(golden_dir, "f32_s4d")
    out64 = selective_scan_oracle(ins["u"], ins["delta"], ins["A"], ins["B"], ins["C"], ins["D"], ins["z"],
                                  ins["delta_bias"], delta_softplus=True, compute_dtype=torch.float64)
    np.testing.assert_allclose(out64.numpy(), z["out"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["w4_bias", "w4_nobias", "w2_short"])
def test_conv_oracle_matches_reference(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"conv_{name}.npz"))
    x = torch.from_numpy(z["in_x"]).requires_grad_(True)
    w = torch.from_numpy(z["in_weight"]).requires_grad_(True)
    b = torch.from_numpy(z["in_bias"]).requires_grad_(True) if "in_bias" in z else None
    y = causal_conv1d_oracle(x, w, b, activation="silu")
    assert np.array_equal(y.detach().numpy(), z["out"])
    (y * torch.from_numpy(z["cotangent"])).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), z["grad_x"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(w.grad.numpy(), z["grad_weight"], rtol=1e-5, atol=1e-5)
    if b is not None:
        np.testing.assert_allclose(b.grad.numpy(), z["grad_bias"], rtol=1e-5, atol=1e-5)


def test_conv_oracle_anticausal_equals_flip(golden_dir):
    z = np.load(os.path.join(golden_dir, "conv_w4_bias.npz"))
    x, w, b = (torch.from_numpy(z[k]) for k in ("in_x", "in_weight", "in_bias"))
    a = causal_conv1d_oracle(x, w, b, "silu", anticausal=True)
    ref = causal_conv1d_oracle(x.flip(-1), w, b, "silu").flip(-1)
    assert torch.equal(a, ref)
    # explicit index form of SURVEY 9.1.2: out[l] = silu(b + sum_k w[k] x[l+(W-1)-k])
    L, W = x.shape[-1], w.shape[1]
    xp = torch.nn.functional.pad(x, (0, W - 1))
    acc = b[None, :, None].expand_as(x).clone()
    for k in range(W):
        acc = acc + w[None, :, k, None] * xp[..., (W - 1 - k):(W - 1 - k) + L]
    torch.testing.assert_close(a, torch.nn.functional.silu(acc), rtol=1e-5, atol=1e-6)


def test_bimamba_oracle_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "bimamba_v2.npz"))
    p = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p_")}
    hidden = torch.from_numpy(z["hidden"])
    y = bimamba_v2_oracle(hidden, p, if_devide_out=True)
    np.testing.assert_allclose(y.numpy(), z["out"], rtol=1e-5, atol=1e-6)
    y2 = bimamba_v2_oracle(hidden, p, if_devide_out=False)
    np.testing.assert_allclose(y2.numpy(), z["out_nodivide"], rtol=1e-5, atol=1e-6)


def test_fbank_oracle_vs_independent_dft():
    g = torch.Generator().manual_seed(3402)
    wav = 0.1 * torch.randn(2, 16000, generator=g)
    for n_fft, win in ((512, 32), (400, 25)):
        a = fbank_oracle(wav, n_fft=n_fft, win_length_ms=win)
        assert a.shape == (2, fbank_frames(16000), 80) and a.dtype == torch.float32
        b = fbank_numpy_dft(wav.numpy(), n_fft=n_fft, win_length_ms=win)
        np.testing.assert_allclose(a.numpy(), b, rtol=0, atol=2e-3)     # dB units


def test_fbank_topdb_floor_and_mel_shape():
    fb = mel_filterbank_matrix(512, 80)
    assert fb.shape == (257, 80) and float(fb.min()) == 0.0 and float(fb.max()) <= 1.0
    assert int((fb > 0).sum(0).min()) >= 1
    wav = torch.zeros(1, 4000)
    wav[0, 2000:2100] = 1.0
    out = fbank_oracle(wav)
    assert float(out.max() - out.min()) <= 80.0 + 1e-4


def test_length_integers_bit_exact():
    assert fbank_frames(160000) == 1001 and encoder_frames(1001) == 251
    assert fbank_frames(240000) == 1501 and encoder_frames(1501) == 376
    assert fbank_frames(320000) == 2001 and encoder_frames(2001) == 501
    assert encoder_frames(30001) == 7501
    wl = torch.linspace(0.5, 1.0, 64)
    for L in (251, 376, 501, 7501):
        a = abs_lengths(wl, L)
        assert np.array_equal(a.numpy(), abs_lengths_numpy(wl.numpy(), L))
        m = key_padding_mask(wl, L)
        assert m.shape == (64, L) and m.dtype == torch.bool
        assert np.array_equal((~m).sum(1).numpy(), a.numpy().astype(np.int64))
    # half-to-even: 0.5 * 501 = 250.5 -> 250
    assert float(abs_lengths(torch.tensor([0.5]), 501)) == 250.0
