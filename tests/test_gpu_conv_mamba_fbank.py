"""GPU parity: causal conv1d kernels, the fused BiMamba-v2 block, the Fbank tail and the extension-module shims."""
import json
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cm_testutil import assert_close, channel_last, make_scan_inputs

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------ conv
@pytest.mark.parametrize("name", ["w4_bias", "w4_nobias", "w2_short"])
@pytest.mark.parametrize("layout", ["tc", "cl"])
def test_conv_matches_reference_golden(golden_dir, name, layout):
    from mamba_asr_b200.causal_conv1d import causal_conv1d_fn
    z = np.load(os.path.join(golden_dir, f"conv_{name}.npz"))
    x = torch.from_numpy(z["in_x"]).cuda()
    if layout == "cl":
        x = channel_last(x)
    x.requires_grad_(True)
    w = torch.from_numpy(z["in_weight"]).cuda().requires_grad_(True)
    b = torch.from_numpy(z["in_bias"]).cuda().requires_grad_(True) if "in_bias" in z else None
    y = causal_conv1d_fn(x, w, b, activation="silu")
    assert_close(y, torch.from_numpy(z["out"]), what="conv out")
    (y * torch.from_numpy(z["cotangent"]).cuda()).sum().backward()
    assert_close(x.grad, torch.from_numpy(z["grad_x"]), what="dx")
    assert_close(w.grad, torch.from_numpy(z["grad_weight"]), floor="max", what="dweight")
    if b is not None:
        assert_close(b.grad, torch.from_numpy(z["grad_bias"]), floor="max", what="dbias")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 64, 67, 4), (1, 40, 130, 3), (3, 288, 100, 4), (2, 33, 17, 2), (2, 64, 1, 4)])
@pytest.mark.parametrize("layout", ["tc", "cl"])
def test_conv_bidirectional_fused(dtype, shape, layout):
    """One launch producing causal (conv1d) and anticausal (conv1d_b on the flipped sequence) outputs, fwd + bwd."""
    from mamba_asr_b200 import kernels as K
    from oracle.conv_ref import causal_conv1d_oracle
    Bt, D, L, W = shape
    g = torch.Generator().manual_seed(L)
    x = torch.randn(Bt, D, L, generator=g).to(dtype)
    ws = [torch.randn(D, W, generator=g) * 0.5 for _ in range(2)]
    bs = [torch.randn(D, generator=g) * 0.5 for _ in range(2)]
    xc = x.clone().float().requires_grad_(True)
    wc = [w.clone().requires_grad_(True) for w in ws]
    bc = [b.clone().requires_grad_(True) for b in bs]
    refs = [causal_conv1d_oracle(xc, wc[0], bc[0], "silu"), causal_conv1d_oracle(xc, wc[1], bc[1], "silu", anticausal=True)]
    cots = [torch.randn(Bt, D, L, generator=g) for _ in range(2)]
    (refs[0] * cots[0] + refs[1] * cots[1]).sum().backward()

    xg = x.cuda()
    if layout == "cl":
        xg = channel_last(xg)
    dirs = [dict(weight=ws[0].cuda(), bias=bs[0].cuda(), anticausal=False),
            dict(weight=ws[1].cuda(), bias=bs[1].cuda(), anticausal=True)]
    outs = K.conv_forward(xg, dirs, silu=True)
    for r in range(2):
        assert_close(outs[r].float(), refs[r], dtype, what=f"out[{r}]")
    douts = [c.to(dtype).cuda() for c in cots]
    if layout == "cl":
        douts = [channel_last(t) for t in douts]
    dx, dws, dbs = K.conv_backward(xg, dirs, douts, silu=True)
    assert_close(dx.float(), xc.grad, dtype, what="dx")
    for r in range(2):
        assert_close(dws[r], wc[r].grad, dtype, floor="max", what=f"dw[{r}]")
        assert_close(dbs[r], bc[r].grad, dtype, floor="max", what=f"db[{r}]")


def test_conv_update_matches_oracle():
    from mamba_asr_b200.causal_conv1d import causal_conv1d_update
    from oracle.conv_ref import causal_conv1d_update_oracle
    g = torch.Generator().manual_seed(0)
    Bt, D, W = 3, 64, 4
    state = torch.randn(Bt, D, W, generator=g)
    w, b = torch.randn(D, W, generator=g), torch.randn(D, generator=g)
    sg = state.clone().cuda()
    for _ in range(5):
        x = torch.randn(Bt, D, generator=g)
        ref = causal_conv1d_update_oracle(x, state, w, b, "silu")
        out = causal_conv1d_update(x.cuda(), sg, w.cuda(), b.cuda(), "silu")
        assert_close(out, ref, what="update out")
        assert torch.equal(sg.cpu(), state)


# ------------------------------------------------------------------------------------------------ mamba block
def _load_bimamba(golden_dir):
    z = np.load(os.path.join(golden_dir, "bimamba_v2.npz"))
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p_")}
    return z, sd


def test_bimamba_v2_module_matches_reference_composition(golden_dir):
    from mamba_asr_b200 import Mamba
    z, sd = _load_bimamba(golden_dir)
    meta = json.load(open(os.path.join(golden_dir, "mamba_state_dict.json")))
    m = Mamba(d_model=meta["d_model"], bimamba_type="v2").cuda()
    m.load_state_dict(sd, strict=True)
    hidden = torch.from_numpy(z["hidden"]).cuda()
    out = m(hidden)
    assert_close(out, torch.from_numpy(z["out"]), what="bimamba v2 out", rtol_mul=2.0)
    m.if_devide_out = False
    assert_close(m(hidden), torch.from_numpy(z["out_nodivide"]), what="bimamba v2 out (sum)", rtol_mul=2.0)


def test_bimamba_v2_gradients_match_oracle_autograd(golden_dir):
    from mamba_asr_b200 import Mamba
    from oracle.bimamba_ref import bimamba_v2_oracle
    z, sd = _load_bimamba(golden_dir)
    meta = json.load(open(os.path.join(golden_dir, "mamba_state_dict.json")))
    m = Mamba(d_model=meta["d_model"], bimamba_type="v2").cuda()
    m.load_state_dict(sd, strict=True)
    hidden = torch.from_numpy(z["hidden"])
    hc = hidden.clone().requires_grad_(True)
    pc = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = bimamba_v2_oracle(hc, pc)
    cot = torch.randn(ref.shape, generator=torch.Generator().manual_seed(3))
    (ref * cot).sum().backward()
    hg = hidden.cuda().requires_grad_(True)
    out = m(hg)
    (out * cot.cuda()).sum().backward()
    assert_close(hg.grad, hc.grad, what="d hidden", rtol_mul=2.0)
    for name, p in m.named_parameters():
        assert p.grad is not None, name
        assert_close(p.grad, pc[name].grad, floor="max", what=f"d {name}", rtol_mul=2.0)


def test_bimamba_v2_bf16_autocast_within_tolerance(golden_dir):
    """bf16 path (autocast, as precision: bf16 in the YAML) against the fp32 oracle fed the same parameters."""
    from mamba_asr_b200 import Mamba
    z, sd = _load_bimamba(golden_dir)
    meta = json.load(open(os.path.join(golden_dir, "mamba_state_dict.json")))
    m = Mamba(d_model=meta["d_model"], bimamba_type="v2").cuda()
    m.load_state_dict(sd, strict=True)
    hidden = torch.from_numpy(z["hidden"]).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(hidden)
    assert out.dtype == torch.bfloat16
    assert_close(out.float(), torch.from_numpy(z["out"]), torch.bfloat16, what="bf16 out")


def test_unimamba_matches_oracle():
    from mamba_asr_b200 import UniMamba
    from oracle.bimamba_ref import mamba_inner_oracle
    torch.manual_seed(5)
    m = UniMamba(d_model=48).cuda()
    for p in m.parameters():
        if p.dim() > 1:
            torch.nn.init.xavier_normal_(p)
    hidden = torch.randn(2, 53, 48)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    xz = (sd["in_proj.weight"] @ hidden.reshape(-1, 48).t()).reshape(-1, 2, 53).transpose(0, 1)
    y = mamba_inner_oracle(xz, sd["conv1d.weight"], sd["conv1d.bias"], sd["x_proj.weight"], sd["dt_proj.weight"],
                           -torch.exp(sd["A_log"]), sd["D"], sd["dt_proj.bias"])
    ref = F.linear(y.transpose(1, 2), sd["out_proj.weight"])
    assert_close(m(hidden.cuda()), ref, what="unimamba out", rtol_mul=2.0)


def test_reference_signature_inner_fn_time_contiguous_layout(golden_dir):
    """mamba_inner_fn_no_out_proj with the reference's (B, 2D, L) xz view (strides (L, B*L, 1), bimamba.py:192-196)."""
    from mamba_asr_b200.selective_scan_interface import mamba_inner_fn_no_out_proj
    from oracle.bimamba_ref import mamba_inner_oracle
    z, sd = _load_bimamba(golden_dir)
    hidden = torch.from_numpy(z["hidden"])
    Bt, L, d = hidden.shape
    xz = (sd["in_proj.weight"] @ hidden.reshape(Bt * L, d).t()).reshape(-1, Bt, L).transpose(0, 1)
    A = -torch.exp(sd["A_log"])
    ref = mamba_inner_oracle(xz, sd["conv1d.weight"], sd["conv1d.bias"], sd["x_proj.weight"], sd["dt_proj.weight"], A,
                             sd["D"], sd["dt_proj.bias"])
    xg = (sd["in_proj.weight"].cuda() @ hidden.cuda().reshape(Bt * L, d).t()).reshape(-1, Bt, L).transpose(0, 1)
    out = mamba_inner_fn_no_out_proj(xg, sd["conv1d.weight"].cuda(), sd["conv1d.bias"].cuda(), sd["x_proj.weight"].cuda(),
                                     sd["dt_proj.weight"].cuda(), A.cuda(), None, None, sd["D"].cuda(),
                                     delta_bias=sd["dt_proj.bias"].cuda(), delta_softplus=True)
    assert out.shape == ref.shape
    assert_close(out, ref, what="inner fn out", rtol_mul=2.0)


# ------------------------------------------------------------------------------------------------ extension shims
def test_extension_module_shims_follow_the_pybind_abi():
    """compat/selective_scan_cuda + causal_conv1d_cuda: argument order / return lists of the reference call sites
    (selective_scan_interface.py:42, 67-70, 182, 286)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "compat"))
    import causal_conv1d_cuda
    import selective_scan_cuda
    from oracle.conv_ref import causal_conv1d_oracle
    from oracle.scan_ref import selective_scan_oracle
    ins = make_scan_inputs(2, 64, 70, 16, torch.float32, seed=9)
    c = {k: v.cuda() for k, v in ins.items()}
    B4, C4 = c["B"].unsqueeze(1).contiguous(), c["C"].unsqueeze(1).contiguous()
    out, x, out_z = selective_scan_cuda.fwd(c["u"], c["delta"], c["A"], B4, C4, c["D"], c["z"], c["delta_bias"], True)
    ref_z, last = selective_scan_oracle(ins["u"], ins["delta"], ins["A"], ins["B"], ins["C"], ins["D"], ins["z"],
                                        ins["delta_bias"], True, return_last_state=True)
    ref_pre = selective_scan_oracle(ins["u"], ins["delta"], ins["A"], ins["B"], ins["C"], ins["D"], None,
                                    ins["delta_bias"], True)
    assert_close(out_z, ref_z, what="out_z")
    assert_close(out, ref_pre, what="out")
    assert x.dim() == 4 and x.shape[:2] == (2, 64) and x.shape[-1] == 32
    assert_close(x[:, :, -1, 1::2], last, what="last_state via x[:, :, -1, 1::2]")
    dout = torch.randn_like(out_z)
    dz = torch.empty_like(c["z"])
    ret = selective_scan_cuda.bwd(c["u"], c["delta"], c["A"], B4, C4, c["D"], c["z"], c["delta_bias"], dout, x, out, dz,
                                  True, True)
    assert len(ret) == 9 and ret[7].data_ptr() == dz.data_ptr() and ret[3].shape == B4.shape
    leaf = {k: v.clone().requires_grad_(True) for k, v in ins.items()}
    r = selective_scan_oracle(leaf["u"], leaf["delta"], leaf["A"], leaf["B"], leaf["C"], leaf["D"], leaf["z"],
                              leaf["delta_bias"], True)
    (r * dout.cpu()).sum().backward()
    assert_close(ret[0], leaf["u"].grad, what="du")
    assert_close(ret[1], leaf["delta"].grad, what="ddelta")
    assert_close(ret[2], leaf["A"].grad, floor="max", what="dA")
    assert_close(ret[3][:, 0], leaf["B"].grad, floor="max", what="dB")
    assert_close(ret[7], leaf["z"].grad, what="dz")
    assert_close(ret[8], ref_z, what="recomputed out_z")

    g = torch.Generator().manual_seed(1)
    xx, w, b = torch.randn(2, 64, 50, generator=g), torch.randn(64, 4, generator=g), torch.randn(64, generator=g)
    y = causal_conv1d_cuda.causal_conv1d_fwd(xx.cuda(), w.cuda(), b.cuda(), None, True)
    assert_close(y, causal_conv1d_oracle(xx, w, b, "silu"), what="conv fwd shim")
    dxz = torch.empty(2, 128, 50, device="cuda")
    dx, dw, db = causal_conv1d_cuda.causal_conv1d_bwd(xx.cuda(), w.cuda(), b.cuda(), torch.ones_like(y), None,
                                                      dxz[:, :64], True)
    assert dx.data_ptr() == dxz.data_ptr() and dw.shape == (64, 4) and db.shape == (64,)


# ------------------------------------------------------------------------------------------------ fbank
@pytest.mark.parametrize("route", ["dft_kernel", "cufft"])
@pytest.mark.parametrize("cfg", [dict(n_fft=512, win_length=32), dict(n_fft=400, win_length=25), dict(n_fft=512, win_length=25)])
def test_fbank_matches_oracle(cfg, route, monkeypatch):
    """Both routes of the front-end against the oracle: the one-kernel windowed DFT (cm_fbank_wav_logmel, default) and
    torch.stft + cm_fbank_logmel (CM_FBANK_CUFFT=1); n_fft 512 / win 25 ms is hparams/CTC/conmamba_large.yaml:103-105."""
    if route == "cufft":
        monkeypatch.setenv("CM_FBANK_CUFFT", "1")
    from mamba_asr_b200 import Fbank
    from oracle.fbank_ref import fbank_oracle
    g = torch.Generator().manual_seed(3402)
    wav = 0.1 * torch.randn(3, 16000 * 2 + 37, generator=g)
    wav[1, 9000:] = 0.0                      # a padded utterance: the top_db floor must kick in
    fb = Fbank(sample_rate=16000, n_mels=80, **cfg).cuda()
    out = fb(wav.cuda())
    ref = fbank_oracle(wav, n_fft=cfg["n_fft"], win_length_ms=cfg["win_length"])
    assert out.shape == ref.shape and out.dtype == torch.float32
    # dB values; same STFT algorithm on a different FFT library: compare with the fp32 contract on the dB scale
    assert_close(out, ref, floor="max", what="fbank dB")
    assert float((out[1].max() - out[1].min()).cpu()) <= 80.0 + 1e-3


def test_fbank_dft_kernel_agrees_with_cufft_route_on_ragged_lengths(monkeypatch):
    """The in-kernel DFT against cuFFT on the same samples: lengths that end inside a frame, a single frame, silence, and a
    tone (weak bins 100 dB below the strongest: the fp32 two-stage DFT must not raise their floor)."""
    from mamba_asr_b200 import Fbank
    g = torch.Generator().manual_seed(77)
    for n_fft, win in ((400, 25), (512, 32), (512, 25)):
        fb = Fbank(sample_rate=16000, n_mels=80, n_fft=n_fft, win_length=win).cuda()
        for n in (159, 160, 1601, 16000 + 81, 48000):
            wav = 0.05 * torch.randn(2, n, generator=g)
            wav[1] = torch.sin(2 * 3.14159265 * 440.0 * torch.arange(n) / 16000.0)
            a = fb(wav.cuda())
            monkeypatch.setenv("CM_FBANK_CUFFT", "1")
            b = fb(wav.cuda())
            monkeypatch.delenv("CM_FBANK_CUFFT")
            assert a.shape == b.shape == (2, 1 + n // 160, 80)
            assert_close(a, b, floor="max", what=f"fbank dft vs cufft n_fft={n_fft} n={n}")


def test_fbank_frame_count_is_bit_exact():
    from mamba_asr_b200 import Fbank
    from oracle.lengths_ref import fbank_frames
    fb = Fbank(n_fft=400, n_mels=80).cuda()
    for n in (16000, 159999, 160000, 160001, 240000):
        assert fb(torch.zeros(1, n, device="cuda")).shape[1] == fbank_frames(n)


@pytest.mark.parametrize("shape", [(4, 37, 144), (2, 501, 256), (3, 5, 33), (1, 1, 512), (2, 9, 1024)])
@pytest.mark.parametrize("dtypes", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16),
                                    (torch.float32, torch.bfloat16), (torch.bfloat16, torch.float32)])
def test_layernorm_matches_torch_fp32_reference(shape, dtypes):
    """cm_layernorm_fwd / cm_layernorm_bwd against F.layer_norm evaluated in fp32 on the same (rounded) inputs:
    output, dx, dgamma, dbeta.  Tolerance: fp32 1e-4, bf16 2e-2 (BASELINE.json)."""
    import torch.nn.functional as F
    from mamba_asr_b200.layernorm import layer_norm
    xd, yd = dtypes
    g = torch.Generator().manual_seed(3)
    Cn = shape[-1]
    x = (torch.randn(*shape, generator=g) * 2 + 0.5).to(xd).cuda()
    w = (1 + 0.1 * torch.randn(Cn, generator=g)).cuda()
    b = (0.1 * torch.randn(Cn, generator=g)).cuda()
    cot = torch.randn(*shape, generator=g).to(yd).cuda()
    xr = x.float().clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (Cn,), wr, br, 1e-5)
    (ref * cot.float()).sum().backward()
    xk = x.clone().requires_grad_(True)
    wk, bk = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    out = layer_norm(xk, wk, bk, 1e-5, out_dtype=yd)
    assert out.dtype == yd and out.shape == x.shape
    (out.float() * cot.float()).sum().backward()
    assert_close(out.float(), ref, yd, what="ln out")
    assert xk.grad.dtype == xd
    assert_close(xk.grad.float(), xr.grad, xd, what="ln dx")
    tol_dt = torch.float32 if (xd == torch.float32 and yd == torch.float32) else torch.bfloat16
    assert_close(wk.grad, wr.grad, tol_dt, floor="max", what="ln dgamma")
    assert_close(bk.grad, br.grad, tol_dt, floor="max", what="ln dbeta")


def test_layernorm_module_follows_autocast_and_has_no_cpu_path():
    from mamba_asr_b200.layernorm import FusedLayerNorm
    ln = FusedLayerNorm(144).cuda()
    keep = FusedLayerNorm(144, keep_dtype=True).cuda()
    x = torch.randn(2, 7, 144, device="cuda")
    assert ln(x).dtype == torch.float32
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert ln(x).dtype == torch.bfloat16
        assert keep(x.bfloat16()).dtype == torch.float32      # torch's autocast rule for layer_norm
    assert sorted(ln.state_dict().keys()) == ["bias", "weight"]
    with pytest.raises(RuntimeError):
        FusedLayerNorm(144)(torch.randn(2, 144))


@pytest.mark.parametrize("shape", [(64, 501, 512), (3, 67, 288), (1, 5, 64), (7, 16), (2, 3, 4, 32)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_glu_kernels_match_torch_glu(shape, dtype):
    """cm_glu_fwd / cm_glu_bwd (the gate of the convolution module, reference modules/Conmamba.py:268-279) against F.glu and
    its autograd gradient evaluated in fp32 on the same rounded inputs; unsupported geometries fall back to torch."""
    from mamba_asr_b200 import kernels as K
    from mamba_asr_b200.layernorm import glu
    g = torch.Generator().manual_seed(23)
    h = (2.0 * torch.randn(*shape, generator=g)).to(dtype).cuda()
    cot = torch.randn(*shape[:-1], shape[-1] // 2, generator=g).to(dtype).cuda()
    assert K.glu_supported(h)
    hr = h.float().clone().requires_grad_(True)
    ref = F.glu(hr, dim=-1)
    (ref * cot.float()).sum().backward()
    hk = h.clone().requires_grad_(True)
    l0 = K.LAUNCHES
    out = glu(hk)
    (out.float() * cot.float()).sum().backward()
    assert K.LAUNCHES == l0 + 2                               # one kernel each way
    assert out.shape == ref.shape and out.dtype == dtype
    assert_close(out.float(), ref, dtype, what="glu y")
    assert_close(hk.grad.float(), hr.grad, dtype, what="glu dh")
    odd = torch.randn(4, 5, 24, device="cuda").to(dtype)     # 12 gated channels: not a multiple of 8 -> torch's op
    assert not K.glu_supported(odd)
    assert torch.equal(glu(odd), F.glu(odd, dim=-1))


@pytest.mark.parametrize("ksize,causal", [(31, False), (31, True), (15, False), (7, False), (3, True)])
@pytest.mark.parametrize("shape", [(2, 501, 256), (3, 67, 144), (1, 5, 33), (2, 94, 64)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16, "bf16-ffma"])
def test_depthwise_conv1d_matches_torch_conv1d(ksize, causal, shape, dtype, monkeypatch):
    """cm_dwconv_fwd / cm_dwconv_bwd_weight against the reference's own op (nn.Conv1d(C, C, K, padding, groups=C),
    modules/Conmamba.py:281-290) evaluated in fp32 on the same (rounded) inputs: y, dx, dweight, dbias."""
    from mamba_asr_b200.dwconv import depthwise_conv1d
    # 16-bit tensors take the tensor-pipe forward / backward-data kernel (taps rounded to the 16-bit type) by default;
    # "bf16-ffma" keeps them on the FFMA tile kernel (CM_DWCONV_NO_MMA=1)
    if isinstance(dtype, str):
        monkeypatch.setenv("CM_DWCONV_NO_MMA", "1")
        dtype = torch.bfloat16
    Bt, L, Cn = shape
    g = torch.Generator().manual_seed(9)
    x = torch.randn(Bt, L, Cn, generator=g).to(dtype).cuda()
    w = (torch.randn(Cn, 1, ksize, generator=g) / ksize ** 0.5).cuda()
    b = (0.1 * torch.randn(Cn, generator=g)).cuda()
    cot = torch.randn(Bt, L, Cn, generator=g).to(dtype).cuda()
    pad = ksize - 1 if causal else (ksize - 1) // 2
    xr = x.float().clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv1d(xr.transpose(1, 2), wr, br, padding=pad, groups=Cn)
    ref = (ref[..., :-pad] if causal else ref).transpose(1, 2)
    (ref * cot.float()).sum().backward()
    xk = x.clone().requires_grad_(True)
    wk, bk = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    out = depthwise_conv1d(xk, wk, bk, pad_left=pad)
    assert out.shape == x.shape and out.dtype == dtype
    (out.float() * cot.float()).sum().backward()
    assert_close(out.float(), ref, dtype, what="dwconv y")
    assert_close(xk.grad.float(), xr.grad, dtype, what="dwconv dx")
    assert_close(wk.grad, wr.grad, dtype, floor="max", what="dwconv dweight")
    assert_close(bk.grad, br.grad, dtype, floor="max", what="dwconv dbias")


@pytest.mark.parametrize("causal", [False, True])
def test_convolution_module_kernel_path_equals_reference_op_chain(causal, monkeypatch):
    """ConvolutionModule on the channel-last kernels == the reference's transpose -> Conv1d -> GLU -> Conv1d chain."""
    from mamba_asr_b200.conmamba import ConvolutionModule
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)    # the reference chain's cuDNN convs in true fp32
    torch.manual_seed(4)
    m = ConvolutionModule(64, kernel_size=31, causal=causal).cuda().eval()
    x = torch.randn(2, 77, 64, device="cuda")
    assert m.use_kernel
    y_k = m(x)
    m.use_kernel = False
    for mod in m.modules():                      # the reference chain uses torch's LayerNorm
        if type(mod).__name__ == "FusedLayerNorm":
            mod.__class__ = torch.nn.LayerNorm
    y_t = m(x)
    assert_close(y_k, y_t, what="conv module")


@pytest.mark.parametrize("shape", [(3, 501, 256, 1024), (2, 37, 144, 576), (1, 5, 32, 10), (4, 9, 64, 11),
                                   (32, 376, 144, 1024), (32, 376, 1024, 144), (64, 501, 256, 256), (5, 1002, 64, 48)])
@pytest.mark.parametrize("autocast", [False, True])
def test_bias_grad_linear_matches_nn_linear(shape, autocast):
    """BiasGradLinear (cm_colsum for db, cuBLAS for the rest) against nn.Linear: output and every gradient, fp32 and
    under bf16 autocast; odd output widths take torch's own reduction; the long-row shapes take the split-K weight
    gradient (4, 8, 16 row blocks and a row count that only divides by 2)."""
    from mamba_asr_b200.linear import BiasGradLinear
    Bt, L, cin, cout = shape
    torch.manual_seed(5)
    ref = torch.nn.Linear(cin, cout).cuda()
    mine = BiasGradLinear(cin, cout).cuda()
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(Bt, L, cin, device="cuda")
    cot = torch.randn(Bt, L, cout, device="cuda")
    outs = []
    for m in (ref, mine):
        xi = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            y = m(xi)
        (y.float() * cot).sum().backward()
        outs.append((y.float(), xi.grad, m.weight.grad, m.bias.grad))
    dt = torch.bfloat16 if autocast else torch.float32
    for a, b, nm in zip(outs[1], outs[0], ("y", "dx", "dw", "db")):
        assert a.dtype == b.dtype
        assert_close(a.float(), b.float(), dt, floor="max", what="linear " + nm)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_extension_shims_with_the_reference_tensor_geometry(dtype):
    """The B3 shims driven with EXACTLY the tensor geometry of the reference's MambaInnerFnNoOutProj
    (selective_scan_interface.py:180-187, 218, 249-256, 286): xz produced by `in_proj.weight @ hidden` rearranged
    "d (b l) -> b d l" (strides (L, B*L, 1)), x / z as chunk views of it, delta with the same strided layout, B / C as
    contiguous (b 1 n l), dz_ and dx_ as strided halves of ONE dxz buffer.  Everything is checked against the oracle."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "compat"))
    import causal_conv1d_cuda
    import selective_scan_cuda
    from oracle.conv_ref import causal_conv1d_oracle
    from oracle.scan_ref import selective_scan_oracle
    g = torch.Generator().manual_seed(17)
    Bt, L, dm, D, R, N, W = 3, 53, 48, 96, 3, 16, 4
    hidden = torch.randn(Bt, L, dm, generator=g)
    w_in = torch.randn(2 * D, dm, generator=g) * dm ** -0.5
    conv_w, conv_b = torch.randn(D, W, generator=g) * 0.5, torch.randn(D, generator=g) * 0.1
    w_x = torch.randn(R + 2 * N, D, generator=g) * D ** -0.5
    w_dt = torch.randn(D, R, generator=g) * R ** -0.5
    A = -torch.exp(torch.randn(D, N, generator=g) * 0.2)
    Dk = torch.ones(D)
    dt_bias = torch.log(torch.expm1(torch.full((D,), 0.02)))
    cast = lambda t: t.to(dtype)
    dev = lambda t: t.cuda()

    def run(xz, on_gpu):
        """selective_scan_interface.py:180-218 with the op names of that file; returns the tensors handed around."""
        x, z = xz.chunk(2, dim=1)
        assert x.stride() == (L, Bt * L, 1)
        if on_gpu:
            conv1d_out = causal_conv1d_cuda.causal_conv1d_fwd(x, dev(conv_w), dev(conv_b), None, True)
        else:
            conv1d_out = causal_conv1d_oracle(x, conv_w, conv_b, "silu")
        wx, wdt = (dev(cast(w_x)), dev(cast(w_dt))) if on_gpu else (cast(w_x), cast(w_dt))
        x_dbl = torch.nn.functional.linear(conv1d_out.permute(0, 2, 1).reshape(Bt * L, D), wx)
        delta = (wdt @ x_dbl[:, :R].t()).view(D, Bt, L).permute(1, 0, 2)                # "d (b l) -> b d l": a strided view
        assert delta.stride() == (L, Bt * L, 1)
        Bm = x_dbl[:, R:R + N].view(Bt, L, N).permute(0, 2, 1).unsqueeze(1).contiguous()
        Cm = x_dbl[:, -N:].view(Bt, L, N).permute(0, 2, 1).unsqueeze(1).contiguous()
        return x, z, conv1d_out, delta, Bm, Cm

    xz_c = cast((w_in @ hidden.permute(2, 0, 1).reshape(dm, Bt * L)).view(2 * D, Bt, L).permute(1, 0, 2))
    assert xz_c.stride() == (L, Bt * L, 1)
    xz_g = torch.empty_strided(xz_c.shape, xz_c.stride(), dtype=dtype, device="cuda").copy_(xz_c)
    x, z, conv1d_out, delta, Bm, Cm = run(xz_g, True)
    cx, cz, c_conv, c_delta, cB, cC = run(xz_c, False)
    assert_close(conv1d_out.float(), c_conv.float(), dtype, what="conv1d_out")
    out, inter, out_z = selective_scan_cuda.fwd(conv1d_out, delta, dev(A), Bm, Cm, dev(Dk), z, dev(dt_bias), True)
    # oracle on what the GPU path actually fed its scan (its own conv1d_out / delta / B / C, rounded to dtype)
    leaf = dict(u=conv1d_out.cpu().float(), delta=delta.cpu().float(), B=Bm[:, 0].cpu().float(), C=Cm[:, 0].cpu().float(),
                z=z.cpu().float())
    leaf = {k: v.clone().requires_grad_(True) for k, v in leaf.items()}
    Al, Dl, bl = A.clone().requires_grad_(True), Dk.clone().requires_grad_(True), dt_bias.clone().requires_grad_(True)
    ref = selective_scan_oracle(leaf["u"], leaf["delta"], Al, leaf["B"], leaf["C"], Dl, leaf["z"], bl, True)
    assert_close(out_z.float(), ref.detach(), dtype, what="out_z")
    dout = cast(torch.randn(Bt, D, L, generator=g))
    (ref * dout.float()).sum().backward()
    dxz = torch.empty_like(xz_g)                                                           # :249-250
    dx, dz = dxz.chunk(2, dim=1)
    ret = selective_scan_cuda.bwd(conv1d_out, delta, dev(A), Bm, Cm, dev(Dk), z, dev(dt_bias), dev(dout), inter, out, dz,
                                  True, True)
    dconv1d_out, ddelta, dA, dB, dC, dD, ddb, dz_ret, out_z2 = ret
    assert dz_ret.data_ptr() == dz.data_ptr()
    assert_close(dz.float(), leaf["z"].grad, dtype, what="dz (strided half of dxz)")
    assert_close(dconv1d_out.float(), leaf["u"].grad, dtype, what="dconv1d_out")
    assert_close(ddelta.float(), leaf["delta"].grad, dtype, what="ddelta")
    assert_close(dB[:, 0].float(), leaf["B"].grad, dtype, floor="max", what="dB")
    assert_close(dC[:, 0].float(), leaf["C"].grad, dtype, floor="max", what="dC")
    assert_close(dA, Al.grad, dtype, floor="max", what="dA")
    assert_close(dD, Dl.grad, dtype, floor="max", what="dD")
    assert_close(ddb, bl.grad, dtype, floor="max", what="ddelta_bias")
    assert_close(out_z2.float(), ref.detach(), dtype, what="recomputed out_z")
    # conv backward into the strided dx half (:286), gradient handed over in the "d (b l) -> b d l" geometry of :282-283
    dco = torch.empty_strided((Bt, D, L), (L, Bt * L, 1), dtype=dtype, device="cuda").copy_(dconv1d_out)
    dx_ret, dcw, dcb = causal_conv1d_cuda.causal_conv1d_bwd(x, dev(conv_w), dev(conv_b), dco, None, dx, True)
    assert dx_ret.data_ptr() == dx.data_ptr()
    xl = cx.float().clone().requires_grad_(True)
    wl, bl2 = conv_w.clone().requires_grad_(True), conv_b.clone().requires_grad_(True)
    (causal_conv1d_oracle(xl, wl, bl2, "silu") * dco.cpu().float()).sum().backward()
    assert_close(dx.float(), xl.grad, dtype, what="dx (strided half of dxz)")
    assert_close(dcw, wl.grad, dtype, floor="max", what="dconv1d_weight")
    assert_close(dcb, bl2.grad, dtype, floor="max", what="dconv1d_bias")
