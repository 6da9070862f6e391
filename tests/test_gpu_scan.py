"""GPU parity of the selective-scan kernels against the oracle and the golden fixtures frozen from the reference.
Every call goes Python wrapper -> ctypes -> C ABI (libconmamba_b200.so) -> sm_100a kernel."""
import numpy as np
import pytest
import torch

from cm_testutil import RTOL, assert_close, channel_last, load_scan_golden, make_scan_inputs

pytestmark = pytest.mark.gpu

GRAD_KEYS = ("u", "delta", "A", "B", "C", "D", "z", "delta_bias")
SUM_KEYS = ("A", "B", "C", "D", "delta_bias")     # gradients that are sums over many positions


def _cuda(ins, layout="tc"):
    """layout: "tc" time-contiguous (B, D, L) memory (the reference's); "cl" channel-last memory, B and C separate tensors;
    "xdbl" channel-last with B | C | dt-pad packed in ONE (B, L, 48) row-major buffer the way the module's x_dbl holds them
    (mamba_inner._aligned_proj_weights) - the layout the TMA kernels (scan_fwd_lc.cu / scan_bwd_lc.cu) take."""
    out = {}
    for k, v in ins.items():
        if v is None:
            out[k] = None
            continue
        t = v.cuda()
        if layout in ("cl", "xdbl") and t.dim() == 3:
            t = channel_last(t)
        out[k] = t
    if layout == "xdbl" and out.get("B") is not None and out["B"].dim() == 3 and out["B"].shape[1] == 16:
        Bt, N, L = out["B"].shape
        xd = torch.zeros(Bt, L, 2 * N + 16, dtype=out["B"].dtype, device="cuda")
        xd[..., :N] = out["B"].transpose(1, 2)
        xd[..., N:2 * N] = out["C"].transpose(1, 2)
        out["B"], out["C"] = xd[..., :N].transpose(1, 2), xd[..., N:2 * N].transpose(1, 2)
    return out


def _fn():
    from mamba_asr_b200.selective_scan_interface import selective_scan_fn
    return selective_scan_fn


@pytest.mark.parametrize("name", ["f32_full", "f32_s4d", "f32_plain4d", "f32_constBC", "bf16_full", "f32_L1"])
@pytest.mark.parametrize("layout", ["tc", "cl"])
def test_scan_forward_matches_reference_golden(golden_dir, name, layout):
    z, meta, ins, dt = load_scan_golden(golden_dir, name)
    c = _cuda(ins, layout)
    out, last = _fn()(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["z"], c["delta_bias"],
                      delta_softplus=meta["softplus"], return_last_state=True)
    assert out.dtype == dt and out.shape == tuple(z["out"].shape)
    assert_close(out.float(), torch.from_numpy(z["out"]), dt, what=f"{name} out")
    assert_close(last, torch.from_numpy(z["last_state"]), dt, what=f"{name} last_state")


@pytest.mark.parametrize("name", ["f32_full", "f32_s4d", "f32_plain4d", "f32_L1"])
@pytest.mark.parametrize("layout", ["tc", "cl"])
def test_scan_backward_matches_reference_autograd(golden_dir, name, layout):
    z, meta, ins, dt = load_scan_golden(golden_dir, name)
    c = _cuda(ins, layout)
    leaf = {k: (v.requires_grad_(True) if v is not None else None) for k, v in c.items()}
    out = _fn()(leaf["u"], leaf["delta"], leaf["A"], leaf["B"], leaf["C"], leaf["D"], leaf["z"], leaf["delta_bias"],
                delta_softplus=meta["softplus"])
    (out * torch.from_numpy(z["cotangent"]).cuda()).sum().backward()
    for k in GRAD_KEYS:
        if leaf[k] is None:
            continue
        assert leaf[k].grad.shape == leaf[k].shape
        assert_close(leaf[k].grad, torch.from_numpy(z["grad_" + k]), dt, floor="max" if k in SUM_KEYS else "rms",
                     what=f"{name} grad_{k}")


@pytest.mark.parametrize("lanes", [1, 2, 4])
@pytest.mark.parametrize("shape", [(2, 64, 67), (1, 40, 131), (3, 96, 8), (2, 32, 9), (2, 33, 40)])
def test_scan_lane_splits_and_ragged_shapes(lanes, shape, monkeypatch):
    """1/2/4 lanes per channel; dims that are not multiples of the warp's channel count; L around the 8-step tile."""
    from oracle.scan_ref import selective_scan_oracle
    monkeypatch.setenv("CM_SCAN_LANES", str(lanes))
    Bt, D, L = shape
    ins = make_scan_inputs(Bt, D, L, 16, torch.float32, seed=L)
    c = _cuda(ins, "cl")
    lc = {k: v.clone().requires_grad_(True) for k, v in ins.items()}
    lg = {k: v.requires_grad_(True) for k, v in c.items()}
    ref = selective_scan_oracle(lc["u"], lc["delta"], lc["A"], lc["B"], lc["C"], lc["D"], lc["z"], lc["delta_bias"], True)
    out = _fn()(lg["u"], lg["delta"], lg["A"], lg["B"], lg["C"], lg["D"], lg["z"], lg["delta_bias"], True)
    assert_close(out, ref, what="out")
    cot = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    (ref * cot).sum().backward()
    (out * cot.cuda()).sum().backward()
    for k in GRAD_KEYS:
        assert_close(lg[k].grad, lc[k].grad, floor="max" if k in SUM_KEYS else "rms", what=f"grad_{k}")


@pytest.mark.parametrize("dstate", [4, 8, 16])
def test_scan_small_dstate(dstate):
    from oracle.scan_ref import selective_scan_oracle
    ins = make_scan_inputs(2, 32, 50, dstate, torch.float32, seed=3)
    c = _cuda(ins, "cl")
    ref = selective_scan_oracle(ins["u"], ins["delta"], ins["A"], ins["B"], ins["C"], ins["D"], ins["z"],
                                ins["delta_bias"], True)
    out = _fn()(c["u"], c["delta"], c["A"], c["B"], c["C"], c["D"], c["z"], c["delta_bias"], True)
    assert_close(out, ref, what=f"dstate {dstate}")


def test_scan_rejects_what_it_does_not_implement():
    ins = make_scan_inputs(1, 32, 16, 32, torch.float32)
    c = _cuda(ins)
    with pytest.raises(NotImplementedError):
        _fn()(c["u"], c["delta"], c["A"], c["B"], c["C"])
    cpu = make_scan_inputs(1, 32, 16, 16, torch.float32)
    with pytest.raises(RuntimeError):      # no CPU fallback
        _fn()(cpu["u"], cpu["delta"], cpu["A"], cpu["B"], cpu["C"])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bidirectional_fused_equals_flip_composition(dtype):
    """cm_scan_fwd with ndir=2 == 0.5*out + 0.5*out_b.flip(-1) of bimamba.py:223-253 built from the oracle."""
    from mamba_asr_b200 import kernels as K
    from oracle.scan_ref import selective_scan_oracle
    Bt, D, L, N = 2, 64, 77, 16
    f = make_scan_inputs(Bt, D, L, N, dtype, seed=1)
    bwd = make_scan_inputs(Bt, D, L, N, dtype, seed=2)
    z = f["z"]
    of = selective_scan_oracle(f["u"], f["delta"], f["A"], f["B"], f["C"], f["D"], z, f["delta_bias"], True)
    fl = lambda t: t.flip(-1)
    ob = selective_scan_oracle(fl(bwd["u"]), fl(bwd["delta"]), bwd["A"], fl(bwd["B"]), fl(bwd["C"]), bwd["D"], fl(z),
                               bwd["delta_bias"], True)
    ref = 0.5 * of.float() + 0.5 * fl(ob).float()
    dirs = []
    for src, rev in ((f, False), (bwd, True)):
        c = _cuda(src, "cl")
        dirs.append(dict(u=c["u"], delta=c["delta"], A=c["A"], B=c["B"], C=c["C"], D=c["D"],
                         delta_bias=c["delta_bias"], reverse=rev))
    res = K.scan_forward(dirs, z=channel_last(z.cuda()), out_scale=0.5, delta_softplus=True)
    assert_close(res["out"].float(), ref, dtype, what="bidir out")


def test_bidirectional_backward_matches_autograd_of_flip_composition():
    from mamba_asr_b200 import kernels as K
    from oracle.scan_ref import selective_scan_oracle
    Bt, D, L, N = 2, 64, 45, 16
    f = make_scan_inputs(Bt, D, L, N, torch.float32, seed=11)
    bw = make_scan_inputs(Bt, D, L, N, torch.float32, seed=12)
    lf = {k: v.clone().requires_grad_(True) for k, v in f.items()}
    lb = {k: v.clone().requires_grad_(True) for k, v in bw.items() if k != "z"}
    fl = lambda t: t.flip(-1)
    of = selective_scan_oracle(lf["u"], lf["delta"], lf["A"], lf["B"], lf["C"], lf["D"], lf["z"], lf["delta_bias"], True)
    ob = selective_scan_oracle(fl(lb["u"]), fl(lb["delta"]), lb["A"], fl(lb["B"]), fl(lb["C"]), lb["D"], fl(lf["z"]),
                               lb["delta_bias"], True)
    ref = 0.5 * of + 0.5 * fl(ob)
    cot = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5))
    (ref * cot).sum().backward()

    dirs = []
    for src, rev in ((f, False), (bw, True)):
        c = _cuda(src, "cl")
        dirs.append(dict(u=c["u"], delta=c["delta"], A=c["A"], B=c["B"], C=c["C"], D=c["D"],
                         delta_bias=c["delta_bias"], reverse=rev))
    zc = channel_last(f["z"].cuda())
    res = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    g = K.scan_backward(dirs, res["ckpt"], channel_last(cot.cuda()), z=zc, out_pre=res["out_pre"], out_scale=0.5,
                        delta_softplus=True)
    assert_close(g["dz"], lf["z"].grad, what="dz")
    for r, leaf in enumerate((lf, lb)):
        assert_close(g["du"][r], leaf["u"].grad, what=f"du[{r}]")
        assert_close(g["ddelta"][r], leaf["delta"].grad, what=f"ddelta[{r}]")
        assert_close(g["dB"][r], leaf["B"].grad, floor="max", what=f"dB[{r}]")
        assert_close(g["dC"][r], leaf["C"].grad, floor="max", what=f"dC[{r}]")
        assert_close(g["dA"][r], leaf["A"].grad, floor="max", what=f"dA[{r}]")
        assert_close(g["dD"][r], leaf["D"].grad, floor="max", what=f"dD[{r}]")
        assert_close(g["dbias"][r], leaf["delta_bias"].grad, floor="max", what=f"dbias[{r}]")


def test_scan_backward_is_deterministic():
    ins = make_scan_inputs(4, 96, 200, 16, torch.bfloat16, seed=7)
    c = _cuda(ins, "cl")
    grads = []
    for _ in range(2):
        leaf = {k: v.clone().requires_grad_(True) for k, v in c.items()}
        out = _fn()(leaf["u"], leaf["delta"], leaf["A"], leaf["B"], leaf["C"], leaf["D"], leaf["z"], leaf["delta_bias"], True)
        out.float().square().sum().backward()
        grads.append({k: leaf[k].grad.clone() for k in GRAD_KEYS})
    for k in GRAD_KEYS:
        assert torch.equal(grads[0][k], grads[1][k]), k


def test_full_size_properties_config3_shapes():
    """ConMamba-large shapes (B 64, D 512, L 501, bf16): properties that need no CPU oracle pass.
    (i) reversed scan == flip -> scan -> flip, bit for bit; (ii) fused bidirectional == the two unidirectional
    kernels combined; (iii) the scan is linear in u when D skip and gate are fixed."""
    from mamba_asr_b200 import kernels as K
    Bt, D, L, N = 64, 512, 501, 16
    f = make_scan_inputs(Bt, D, L, N, torch.bfloat16, seed=21, device="cuda")
    b = make_scan_inputs(Bt, D, L, N, torch.bfloat16, seed=22, device="cuda")
    cl = lambda d: {k: (channel_last(v) if v.dim() == 3 else v) for k, v in d.items()}
    f, b = cl(f), cl(b)
    mk = lambda s, rev, **kw: dict(u=kw.get("u", s["u"]), delta=s["delta"], A=s["A"], B=s["B"], C=s["C"], D=s["D"],
                                   delta_bias=s["delta_bias"], reverse=rev)
    # (i)
    r1 = K.scan_forward([mk(b, True)], z=None, delta_softplus=True)["out"]
    fl = lambda t: channel_last(t.flip(-1))
    bf = {k: (fl(v) if v.dim() == 3 else v) for k, v in b.items()}
    r2 = K.scan_forward([mk(bf, False)], z=None, delta_softplus=True)["out"].flip(-1)
    assert torch.equal(r1, r2)
    # (ii)
    of = K.scan_forward([mk(f, False)], z=None, delta_softplus=True)["out"].float()
    fused = K.scan_forward([mk(f, False), mk(b, True)], z=f["z"], out_scale=0.5, delta_softplus=True)["out"].float()
    zf = f["z"].float()
    comp = 0.5 * (of + r1.float()) * (zf * torch.sigmoid(zf))
    # `comp` rounds each direction to bf16 before the add and the gate (as the reference does); the fused kernel
    # rounds once, so the difference is bounded by bf16 eps of the PRE-gate magnitudes -> floor on max|ref|
    assert_close(fused, comp, torch.bfloat16, floor="max", what="fused vs separate")
    # (iii)
    u2 = channel_last(torch.randn_like(f["u"]))
    o1 = K.scan_forward([mk(f, False)], delta_softplus=True)["out"].float()
    o2 = K.scan_forward([mk(f, False, u=u2)], delta_softplus=True)["out"].float()
    o12 = K.scan_forward([mk(f, False, u=(f["u"] + u2))], delta_softplus=True)["out"].float()
    # o1 and o2 are each rounded to bf16 and may cancel in the sum: the floor is the magnitude of the summands
    assert_close(o12, o1 + o2, torch.bfloat16, floor="max", what="linearity in u")


@pytest.mark.parametrize("cpl", [2, 4])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 64, 67), (1, 32, 131), (3, 96, 8), (2, 32, 9), (2, 32, 1), (2, 64, 16), (1, 32, 17),
                                   (2, 32, 33), (1, 160, 300)])
def test_state_parallel_forward_matches_oracle(cpl, dtype, shape, monkeypatch):
    """scan_fwd_sp.cu (lane = state; the default channel-last kernel): uni- and bidirectional outputs, last state and
    checkpoints against the oracle / the lane-per-channel kernel, over tile-boundary and ragged lengths."""
    from mamba_asr_b200 import kernels as K
    from oracle.scan_ref import selective_scan_oracle
    monkeypatch.setenv("CM_SP_CPL", str(cpl))
    Bt, D, L = shape
    N = 16
    f = make_scan_inputs(Bt, D, L, N, dtype, seed=31)
    bw = make_scan_inputs(Bt, D, L, N, dtype, seed=32)
    z = f["z"]
    fl = lambda t: t.flip(-1)
    of, lf = selective_scan_oracle(f["u"], f["delta"], f["A"], f["B"], f["C"], f["D"], z, f["delta_bias"], True,
                                   return_last_state=True)
    ob = selective_scan_oracle(fl(bw["u"]), fl(bw["delta"]), bw["A"], fl(bw["B"]), fl(bw["C"]), bw["D"], fl(z),
                               bw["delta_bias"], True)
    dirs = []
    for src, rev in ((f, False), (bw, True)):
        c = _cuda(src, "cl")
        dirs.append(dict(u=c["u"], delta=c["delta"], A=c["A"], B=c["B"], C=c["C"], D=c["D"],
                         delta_bias=c["delta_bias"], reverse=rev))
    zc = channel_last(z.cuda())
    uni = K.scan_forward(dirs[:1], z=zc, delta_softplus=True, need_last_state=True, need_ckpt=True)
    assert_close(uni["out"].float(), of.float(), dtype, what="sp uni out")
    assert_close(uni["last_state"][0], lf, dtype, what="sp last_state")
    bi = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    assert_close(bi["out"].float(), 0.5 * of.float() + 0.5 * fl(ob).float(), dtype, what="sp bidir out")
    # the lane-per-channel kernels write the same checkpoints (the backward kernels read them)
    monkeypatch.setenv("CM_SCAN_NO_SP", "1")
    uni_o = K.scan_forward(dirs[:1], z=zc, delta_softplus=True, need_last_state=True, need_ckpt=True)
    bi_o = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    nck1, nck2 = K.num_ckpt(L, 1), K.num_ckpt(L, 2)
    assert_close(uni["ckpt"][0][:, :, :nck1], uni_o["ckpt"][0][:, :, :nck1], dtype, floor="max", what="sp ckpt uni")
    for r in range(2):
        assert_close(bi["ckpt"][r][:, :, :nck2], bi_o["ckpt"][r][:, :, :nck2], dtype, floor="max", what="sp ckpt bidir")
    assert_close(bi["out_pre"].float(), bi_o["out_pre"].float(), dtype, floor="max", what="sp out_pre")


def _bidir_backward_case(Bt, D, L, dtype, seed, layout="cl"):
    from mamba_asr_b200 import kernels as K
    N = 16
    f = make_scan_inputs(Bt, D, L, N, dtype, seed=seed)
    bw = make_scan_inputs(Bt, D, L, N, dtype, seed=seed + 1)
    dirs = []
    for src, rev in ((f, False), (bw, True)):
        c = _cuda(src, layout)
        dirs.append(dict(u=c["u"], delta=c["delta"], A=c["A"], B=c["B"], C=c["C"], D=c["D"],
                         delta_bias=c["delta_bias"], reverse=rev))
    zc = channel_last(f["z"].cuda())
    cot = torch.randn((Bt, D, L), generator=torch.Generator().manual_seed(seed + 2)).to(dtype)
    return K, f, bw, dirs, zc, cot


@pytest.mark.parametrize("shape", [(2, 64, 45), (1, 32, 8), (2, 32, 9), (1, 32, 1), (2, 64, 17), (1, 96, 67), (2, 32, 16)])
def test_state_parallel_backward_matches_autograd(shape):
    """scan_bwd_sp.cu (the default channel-last backward): every gradient of the fused bidirectional block and of a
    unidirectional scan against autograd through the oracle, fp32, over ragged lengths and tile boundaries."""
    from oracle.scan_ref import selective_scan_oracle
    Bt, D, L = shape
    K, f, bw, dirs, zc, cot = _bidir_backward_case(Bt, D, L, torch.float32, seed=41)
    lf = {k: v.clone().requires_grad_(True) for k, v in f.items()}
    lb = {k: v.clone().requires_grad_(True) for k, v in bw.items() if k != "z"}
    fl = lambda t: t.flip(-1)
    of = selective_scan_oracle(lf["u"], lf["delta"], lf["A"], lf["B"], lf["C"], lf["D"], lf["z"], lf["delta_bias"], True)
    ob = selective_scan_oracle(fl(lb["u"]), fl(lb["delta"]), lb["A"], fl(lb["B"]), fl(lb["C"]), lb["D"], fl(lf["z"]),
                               lb["delta_bias"], True)
    ((0.5 * of + 0.5 * fl(ob)) * cot).sum().backward()
    res = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    g = K.scan_backward(dirs, res["ckpt"], channel_last(cot.cuda()), z=zc, out_pre=res["out_pre"], out_scale=0.5,
                        delta_softplus=True)
    assert_close(g["dz"], lf["z"].grad, what="dz")
    for r, leaf in enumerate((lf, lb)):
        assert_close(g["du"][r], leaf["u"].grad, what=f"du[{r}]")
        assert_close(g["ddelta"][r], leaf["delta"].grad, what=f"ddelta[{r}]")
        assert_close(g["dB"][r], leaf["B"].grad, floor="max", what=f"dB[{r}]")
        assert_close(g["dC"][r], leaf["C"].grad, floor="max", what=f"dC[{r}]")
        assert_close(g["dA"][r], leaf["A"].grad, floor="max", what=f"dA[{r}]")
        assert_close(g["dD"][r], leaf["D"].grad, floor="max", what=f"dD[{r}]")
        assert_close(g["dbias"][r], leaf["delta_bias"].grad, floor="max", what=f"dbias[{r}]")
    # unidirectional, no gate, no softplus
    lu = {k: v.clone().requires_grad_(True) for k, v in f.items() if k != "z"}
    ou = selective_scan_oracle(lu["u"], lu["delta"], lu["A"], lu["B"], lu["C"], lu["D"], None, lu["delta_bias"], False)
    (ou * cot).sum().backward()
    r1 = K.scan_forward(dirs[:1], delta_softplus=False, need_ckpt=True)
    g1 = K.scan_backward(dirs[:1], r1["ckpt"], channel_last(cot.cuda()), delta_softplus=False)
    assert_close(g1["du"][0], lu["u"].grad, what="uni du")
    assert_close(g1["ddelta"][0], lu["delta"].grad, what="uni ddelta")
    assert_close(g1["dB"][0], lu["B"].grad, floor="max", what="uni dB")
    assert_close(g1["dC"][0], lu["C"].grad, floor="max", what="uni dC")
    if L > 1:   # L == 1: dA is exactly 0 in the reference (h_{-1} = 0); the kernel's a*h_{-1} = h_0 - du*B leaves rounding noise
        assert_close(g1["dA"][0], lu["A"].grad, floor="max", what="uni dA")


def test_state_parallel_backward_bf16_matches_lane_per_channel_kernel(monkeypatch):
    """bf16 I/O: the two backward kernels see the same rounded inputs; they must agree to bf16 rounding of the outputs."""
    K, f, bw, dirs, zc, cot = _bidir_backward_case(3, 96, 203, torch.bfloat16, seed=51)
    res = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    run = lambda: K.scan_backward(dirs, res["ckpt"], channel_last(cot.cuda()), z=zc, out_pre=res["out_pre"], out_scale=0.5,
                                  delta_softplus=True)
    g_sp = run()
    monkeypatch.setenv("CM_SCAN_NO_SP", "1")
    g_old = run()
    assert_close(g_sp["dz"].float(), g_old["dz"].float(), torch.bfloat16, what="dz")
    for r in range(2):
        for key in ("du", "ddelta"):
            assert_close(g_sp[key][r].float(), g_old[key][r].float(), torch.bfloat16, what=f"{key}[{r}]")
        for key in ("dB", "dC", "dA", "dD", "dbias"):
            assert_close(g_sp[key][r].float(), g_old[key][r].float(), torch.bfloat16, floor="max", what=f"{key}[{r}]")


def test_full_size_backward_properties_config3_shapes(monkeypatch):
    """ConMamba-large shapes (B 64, D 512, L 501, bf16), backward: (i) two launches are bit-identical (fixed-order sums,
    no atomics); (ii) the state-parallel kernel and the lane-per-channel kernel agree on every gradient; (iii) the
    gradients are linear in the output cotangent."""
    from mamba_asr_b200 import kernels as K
    Bt, D, L, N = 64, 512, 501, 16
    f = make_scan_inputs(Bt, D, L, N, torch.bfloat16, seed=61, device="cuda")
    b = make_scan_inputs(Bt, D, L, N, torch.bfloat16, seed=62, device="cuda")
    cl = lambda d: {k: (channel_last(v) if v.dim() == 3 else v) for k, v in d.items()}
    f, b = cl(f), cl(b)
    dirs = [dict(u=s["u"], delta=s["delta"], A=s["A"], B=s["B"], C=s["C"], D=s["D"], delta_bias=s["delta_bias"], reverse=rev)
            for s, rev in ((f, False), (b, True))]
    res = K.scan_forward(dirs, z=f["z"], out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    g1 = channel_last(torch.randn(Bt, D, L, device="cuda").bfloat16())
    g2 = channel_last(torch.randn(Bt, D, L, device="cuda").bfloat16())
    run = lambda g: K.scan_backward(dirs, res["ckpt"], g, z=f["z"], out_pre=res["out_pre"], out_scale=0.5, delta_softplus=True)
    a1, a1b, a2, a12 = run(g1), run(g1), run(g2), run(g1 + g2)
    keys_t, keys_p = ("du", "ddelta", "dB", "dC"), ("dA", "dD", "dbias")
    assert torch.equal(a1["dz"], a1b["dz"])
    for r in range(2):
        for k in keys_t + keys_p:
            assert torch.equal(a1[k][r], a1b[k][r]), (k, r)                      # (i)
    for r in range(2):                                                            # (iii)
        for k in keys_t:
            assert_close(a12[k][r].float(), a1[k][r].float() + a2[k][r].float(), torch.bfloat16, floor="max", what=f"lin {k}[{r}]")
        for k in keys_p:
            assert_close(a12[k][r], a1[k][r] + a2[k][r], torch.bfloat16, floor="max", what=f"lin {k}[{r}]")
    monkeypatch.setenv("CM_SCAN_NO_SP", "1")                                      # (ii)
    o1 = run(g1)
    assert_close(a1["dz"].float(), o1["dz"].float(), torch.bfloat16, floor="max", what="dz vs lane-per-channel")
    for r in range(2):
        for k in keys_t:
            assert_close(a1[k][r].float(), o1[k][r].float(), torch.bfloat16, floor="max", what=f"{k}[{r}] vs lane-per-channel")
        for k in keys_p:
            assert_close(a1[k][r], o1[k][r], torch.bfloat16, floor="max", what=f"{k}[{r}] vs lane-per-channel")


def test_long_sequence_forward_matches_lane_per_channel_kernel(monkeypatch):
    """BASELINE config 5 shape class (batch 4, D 512, L 7501, bf16 and fp32): the state-parallel forward against the
    lane-per-channel kernel (both pinned to the oracle at small sizes)."""
    from mamba_asr_b200 import kernels as K
    for dtype in (torch.bfloat16, torch.float32):
        Bt, D, L, N = 4, 512, 7501, 16
        f = make_scan_inputs(Bt, D, L, N, dtype, seed=71, device="cuda")
        b = make_scan_inputs(Bt, D, L, N, dtype, seed=72, device="cuda")
        cl = lambda d: {k: (channel_last(v) if v.dim() == 3 else v) for k, v in d.items()}
        f, b = cl(f), cl(b)
        dirs = [dict(u=s["u"], delta=s["delta"], A=s["A"], B=s["B"], C=s["C"], D=s["D"], delta_bias=s["delta_bias"], reverse=rev)
                for s, rev in ((f, False), (b, True))]
        run = lambda: K.scan_forward(dirs, z=f["z"], out_scale=0.5, delta_softplus=True)["out"].float()
        monkeypatch.delenv("CM_SCAN_NO_SP", raising=False)
        o_sp = run()
        monkeypatch.setenv("CM_SCAN_NO_SP", "1")
        o_old = run()
        assert_close(o_sp, o_old, dtype, floor="max", what=f"long forward {dtype}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,windows", [((2, 32, 1500), 3), ((1, 64, 2049), 5), ((2, 32, 1024), 4)])
def test_time_windowed_forward_equals_whole_sequence_launch(dtype, shape, windows, monkeypatch):
    """Chunk-parallel launch (per-window summaries -> serial combine -> windows from their incoming states) against the
    whole-sequence launch of the same kernel and against the oracle: bidirectional fused, unidirectional + last state."""
    from mamba_asr_b200 import kernels as K
    from oracle.scan_ref import selective_scan_oracle
    Bt, D, L = shape
    N = 16
    f = make_scan_inputs(Bt, D, L, N, dtype, seed=81)
    bw = make_scan_inputs(Bt, D, L, N, dtype, seed=82)
    dirs = []
    for src, rev in ((f, False), (bw, True)):
        c = _cuda(src, "cl")
        dirs.append(dict(u=c["u"], delta=c["delta"], A=c["A"], B=c["B"], C=c["C"], D=c["D"],
                         delta_bias=c["delta_bias"], reverse=rev))
    zc = channel_last(f["z"].cuda())
    monkeypatch.setenv("CM_SCAN_NO_WINDOWS", "1")
    whole_bi = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True)["out"].float()
    whole_uni = K.scan_forward(dirs[1:], z=zc, delta_softplus=True, need_last_state=True)
    monkeypatch.delenv("CM_SCAN_NO_WINDOWS")
    monkeypatch.setenv("CM_SCAN_WINDOWS", str(windows))
    win_bi = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True)["out"].float()
    win_uni = K.scan_forward(dirs[1:], z=zc, delta_softplus=True, need_last_state=True)
    assert_close(win_bi, whole_bi, dtype, floor="max", what="windowed bidir vs whole")
    assert_close(win_uni["out"].float(), whole_uni["out"].float(), dtype, floor="max", what="windowed uni vs whole")
    assert_close(win_uni["last_state"][0], whole_uni["last_state"][0], dtype, floor="max", what="windowed last state")
    fl = lambda t: t.flip(-1)
    of = selective_scan_oracle(f["u"], f["delta"], f["A"], f["B"], f["C"], f["D"], f["z"], f["delta_bias"], True)
    ob = selective_scan_oracle(fl(bw["u"]), fl(bw["delta"]), bw["A"], fl(bw["B"]), fl(bw["C"]), bw["D"], fl(f["z"]),
                               bw["delta_bias"], True)
    # the oracle composition rounds each direction to the I/O dtype before the add (bimamba.py:250-253): for bf16 the
    # floor is the magnitude of the summands
    assert_close(win_bi, 0.5 * of.float() + 0.5 * fl(ob).float(), dtype, floor="rms" if dtype == torch.float32 else "max",
                 what="windowed bidir vs oracle")


# ------------------------------------------------------------------------------------------------ round-2 parity holes
def _oracle_bidir_grads(f, bw, cot, dtype):
    """Oracle forward + autograd of the fused bidirectional block on the SAME (dtype-rounded) inputs, computed in fp32:
    leaves are fp32 copies of the rounded values, so the gradients are the exact adjoint of what the kernel was fed."""
    from oracle.scan_ref import selective_scan_oracle
    lf = {k: v.float().clone().requires_grad_(True) for k, v in f.items()}
    lb = {k: v.float().clone().requires_grad_(True) for k, v in bw.items() if k != "z"}
    fl = lambda t: t.flip(-1)
    of = selective_scan_oracle(lf["u"], lf["delta"], lf["A"], lf["B"], lf["C"], lf["D"], lf["z"], lf["delta_bias"], True)
    ob = selective_scan_oracle(fl(lb["u"]), fl(lb["delta"]), lb["A"], fl(lb["B"]), fl(lb["C"]), lb["D"], fl(lf["z"]),
                               lb["delta_bias"], True)
    ref = 0.5 * of + 0.5 * fl(ob)
    (ref * cot.float()).sum().backward()
    return ref.detach(), lf, lb


def _check_bidir_grads(g, lf, lb, dtype, tag):
    assert_close(g["dz"].float(), lf["z"].grad, dtype, what=f"{tag} dz")
    for r, leaf in enumerate((lf, lb)):
        assert_close(g["du"][r].float(), leaf["u"].grad, dtype, what=f"{tag} du[{r}]")
        assert_close(g["ddelta"][r].float(), leaf["delta"].grad, dtype, what=f"{tag} ddelta[{r}]")
        assert_close(g["dB"][r].float(), leaf["B"].grad, dtype, floor="max", what=f"{tag} dB[{r}]")
        assert_close(g["dC"][r].float(), leaf["C"].grad, dtype, floor="max", what=f"{tag} dC[{r}]")
        assert_close(g["dA"][r], leaf["A"].grad, dtype, floor="max", what=f"{tag} dA[{r}]")
        assert_close(g["dD"][r], leaf["D"].grad, dtype, floor="max", what=f"{tag} dD[{r}]")
        assert_close(g["dbias"][r], leaf["delta_bias"].grad, dtype, floor="max", what=f"{tag} dbias[{r}]")


@pytest.mark.parametrize("kernel", ["default", "tma", "no_sp", "lanes2", "generic"])
@pytest.mark.parametrize("shape", [(2, 64, 45), (1, 96, 203), (2, 32, 9)])
def test_bf16_backward_matches_oracle_autograd(kernel, shape, monkeypatch):
    """bf16 I/O (what BASELINE configs 2-4 train with): every backward kernel - the default, the lane-per-channel
    kernels (1 and 2 lanes) and the generic-stride kernel ((B, D, L) memory) - against autograd through the oracle fed the
    bf16-rounded inputs and computed in fp32; tolerance rtol 2e-2 (BASELINE.json north_star)."""
    Bt, D, L = shape
    K, f, bw, dirs, zc, cot = _bidir_backward_case(Bt, D, L, torch.bfloat16, seed=91,
                                                   layout="xdbl" if kernel == "tma" else "cl")
    if kernel == "no_sp":
        monkeypatch.setenv("CM_SCAN_NO_SP", "1")
    elif kernel == "lanes2":
        monkeypatch.setenv("CM_SCAN_NO_SP", "1")
        monkeypatch.setenv("CM_SCAN_LANES", "2")
    elif kernel == "generic":                                   # time-contiguous memory: only scan_fwd.cu / scan_bwd.cu take it
        dirs = [{k: (v.contiguous() if torch.is_tensor(v) and v.dim() == 3 else v) for k, v in d.items()} for d in dirs]
        zc = zc.contiguous()
    ref, lf, lb = _oracle_bidir_grads(f, bw, cot, torch.bfloat16)
    res = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    assert_close(res["out"].float(), ref, torch.bfloat16, what=f"{kernel} out")
    go = cot.cuda() if kernel == "generic" else channel_last(cot.cuda())
    g = K.scan_backward(dirs, res["ckpt"], go, z=zc, out_pre=res["out_pre"], out_scale=0.5, delta_softplus=True)
    _check_bidir_grads(g, lf, lb, torch.bfloat16, kernel)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("D", [288, 512, 1024])
@pytest.mark.parametrize("L", [17, 67])
def test_benchmark_widths_forward_backward_match_oracle(dtype, D, L):
    """The channel widths of the BASELINE configs (D = 288: configs 1/2, 512: configs 3/5, 1024: config 4) with short
    sequences the CPU oracle finishes in seconds: fused bidirectional forward and every gradient, fp32 and bf16."""
    K, f, bw, dirs, zc, cot = _bidir_backward_case(2, D, L, dtype, seed=100 + D + L, layout="xdbl")
    ref, lf, lb = _oracle_bidir_grads(f, bw, cot, dtype)
    res = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    assert_close(res["out"].float(), ref, dtype, what="out")
    g = K.scan_backward(dirs, res["ckpt"], channel_last(cot.cuda()), z=zc, out_pre=res["out_pre"], out_scale=0.5,
                        delta_softplus=True)
    _check_bidir_grads(g, lf, lb, dtype, "D%d L%d" % (D, L))
    # unidirectional (decoder, config 4)
    from oracle.scan_ref import selective_scan_oracle
    lu = {k: v.float().clone().requires_grad_(True) for k, v in f.items()}
    ou = selective_scan_oracle(lu["u"], lu["delta"], lu["A"], lu["B"], lu["C"], lu["D"], lu["z"], lu["delta_bias"], True)
    (ou * cot.float()).sum().backward()
    r1 = K.scan_forward(dirs[:1], z=zc, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    assert_close(r1["out"].float(), ou.detach(), dtype, what="uni out")
    g1 = K.scan_backward(dirs[:1], r1["ckpt"], channel_last(cot.cuda()), z=zc, out_pre=r1["out_pre"], delta_softplus=True)
    assert_close(g1["du"][0].float(), lu["u"].grad, dtype, what="uni du")
    assert_close(g1["ddelta"][0].float(), lu["delta"].grad, dtype, what="uni ddelta")
    assert_close(g1["dz"].float(), lu["z"].grad, dtype, what="uni dz")
    assert_close(g1["dB"][0].float(), lu["B"].grad, dtype, floor="max", what="uni dB")
    assert_close(g1["dC"][0].float(), lu["C"].grad, dtype, floor="max", what="uni dC")
    assert_close(g1["dA"][0], lu["A"].grad, dtype, floor="max", what="uni dA")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(2, 64, 67), (1, 32, 131), (3, 96, 8), (2, 32, 9), (2, 32, 1), (2, 64, 16), (1, 32, 17),
                                   (2, 32, 33), (1, 160, 300), (5, 32, 15), (1, 64, 32), (3, 32, 48)])
def test_tma_forward_matches_oracle(dtype, shape, monkeypatch):
    """scan_fwd_lc.cu (lane = channel, TMA-staged operands; the default for the module's x_dbl layout): unidirectional
    (gate, last state, checkpoints, no-softplus) and fused bidirectional outputs against the oracle over tile-boundary and
    ragged lengths, odd numbers of channel blocks (inactive warps), and against the state-parallel kernel's checkpoints."""
    from mamba_asr_b200 import kernels as K
    from oracle.scan_ref import selective_scan_oracle
    monkeypatch.setenv("CM_SCAN_LC", "1")
    Bt, D, L = shape
    N = 16
    f = make_scan_inputs(Bt, D, L, N, dtype, seed=131)
    bw = make_scan_inputs(Bt, D, L, N, dtype, seed=132)
    z = f["z"]
    fl = lambda t: t.flip(-1)
    of, lf = selective_scan_oracle(f["u"], f["delta"], f["A"], f["B"], f["C"], f["D"], z, f["delta_bias"], True,
                                   return_last_state=True)
    ob = selective_scan_oracle(fl(bw["u"]), fl(bw["delta"]), bw["A"], fl(bw["B"]), fl(bw["C"]), bw["D"], fl(z),
                               bw["delta_bias"], True)
    o_plain = selective_scan_oracle(f["u"], f["delta"], f["A"], f["B"], f["C"], None, None, None, False)
    dirs = []
    for src, rev in ((f, False), (bw, True)):
        c = _cuda(src, "xdbl")
        dirs.append(dict(u=c["u"], delta=c["delta"], A=c["A"], B=c["B"], C=c["C"], D=c["D"],
                         delta_bias=c["delta_bias"], reverse=rev))
    zc = channel_last(z.cuda())
    uni = K.scan_forward(dirs[:1], z=zc, delta_softplus=True, need_last_state=True, need_ckpt=True)
    assert_close(uni["out"].float(), of.float(), dtype, what="tma uni out")
    assert_close(uni["last_state"][0], lf, dtype, what="tma last_state")
    plain = K.scan_forward([dict(dirs[0], D=None, delta_bias=None)], z=None, delta_softplus=False)
    assert_close(plain["out"].float(), o_plain.float(), dtype, what="tma plain out (no D, no z, no softplus)")
    bi = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    assert_close(bi["out"].float(), 0.5 * of.float() + 0.5 * fl(ob).float(), dtype,
                 floor="rms" if dtype == torch.float32 else "max", what="tma bidir out")
    # same checkpoints / pre-gate sums as the state-parallel kernel (the backward kernels read them)
    monkeypatch.setenv("CM_SCAN_NO_LC", "1")
    uni_o = K.scan_forward(dirs[:1], z=zc, delta_softplus=True, need_last_state=True, need_ckpt=True)
    bi_o = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    nck1, nck2 = K.num_ckpt(L, 1), K.num_ckpt(L, 2)
    assert_close(uni["ckpt"][0][:, :, :nck1], uni_o["ckpt"][0][:, :, :nck1], dtype, floor="max", what="tma ckpt uni")
    for r in range(2):
        assert_close(bi["ckpt"][r][:, :, :nck2], bi_o["ckpt"][r][:, :, :nck2], dtype, floor="max", what="tma ckpt bidir")
    assert_close(bi["out_pre"].float(), bi_o["out_pre"].float(), dtype, floor="max", what="tma out_pre")


def test_tma_forward_full_size_config3_against_state_parallel(monkeypatch):
    """ConMamba-large shape (B 64, D 512, L 501, bf16): the TMA kernel against the state-parallel kernel (both pinned to
    the oracle at small sizes), plus bit-identical repeat launches."""
    from mamba_asr_b200 import kernels as K
    Bt, D, L, N = 64, 512, 501, 16
    g = torch.Generator(device="cuda").manual_seed(5)
    rn = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
    cl = lambda: rn(Bt, L, D).bfloat16().transpose(1, 2)
    z = cl()
    dirs = []
    for rev in (False, True):
        xd = rn(Bt, L, 48).bfloat16()
        dirs.append(dict(u=cl(), delta=(0.5 * rn(Bt, L, D)).bfloat16().transpose(1, 2), A=-torch.exp(0.3 * rn(D, N)),
                         B=xd[..., :N].transpose(1, 2), C=xd[..., N:2 * N].transpose(1, 2), D=torch.ones(D, device="cuda"),
                         delta_bias=torch.full((D,), -4.0, device="cuda"), reverse=rev))
    run = lambda: K.scan_forward(dirs, z=z, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    monkeypatch.setenv("CM_SCAN_LC", "1")
    a, b = run(), run()
    assert torch.equal(a["out"], b["out"]) and torch.equal(a["ckpt"][0], b["ckpt"][0])
    monkeypatch.setenv("CM_SCAN_NO_LC", "1")
    o = run()
    assert_close(a["out"].float(), o["out"].float(), torch.bfloat16, floor="max", what="out")
    assert_close(a["out_pre"].float(), o["out_pre"].float(), torch.bfloat16, floor="max", what="out_pre")
    for r in range(2):
        assert_close(a["ckpt"][r], o["ckpt"][r], torch.bfloat16, floor="max", what="ckpt[%d]" % r)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 128, 45), (1, 128, 8), (2, 128, 9), (1, 128, 1), (2, 256, 17), (1, 128, 67), (2, 128, 16),
                                   (3, 128, 131), (1, 384, 24)])
def test_tma_backward_matches_autograd(dtype, shape, monkeypatch):
    """scan_bwd_lc.cu (lane = channel, TMA-staged; the default backward for the module's x_dbl layout when dim is a multiple
    of 128): every gradient of the fused bidirectional block and of unidirectional scans (with and without gate / softplus)
    against autograd through the oracle (fed the dtype-rounded inputs, computed in fp32), over ragged lengths and tile
    boundaries."""
    from oracle.scan_ref import selective_scan_oracle
    monkeypatch.setenv("CM_SCAN_LC", "1")
    monkeypatch.setenv("CM_SCAN_LC_BWD", "1")
    Bt, D, L = shape
    K, f, bw, dirs, zc, cot = _bidir_backward_case(Bt, D, L, dtype, seed=141, layout="xdbl")
    ref, lf, lb = _oracle_bidir_grads(f, bw, cot, dtype)
    res = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    g = K.scan_backward(dirs, res["ckpt"], channel_last(cot.cuda()), z=zc, out_pre=res["out_pre"], out_scale=0.5,
                        delta_softplus=True)
    _check_bidir_grads(g, lf, lb, dtype, "tma bidir")
    # unidirectional, no gate, no softplus (reverse direction alone: the decoder never uses it, the kernel must still be right)
    for r, src in ((0, f), (1, bw)):
        lu = {k: v.float().clone().requires_grad_(True) for k, v in src.items() if k != "z"}
        fl = (lambda t: t.flip(-1)) if r == 1 else (lambda t: t)
        ou = fl(selective_scan_oracle(fl(lu["u"]), fl(lu["delta"]), lu["A"], fl(lu["B"]), fl(lu["C"]), lu["D"], None,
                                      lu["delta_bias"], False))
        (ou * cot.float()).sum().backward()
        r1 = K.scan_forward(dirs[r:r + 1], delta_softplus=False, need_ckpt=True)
        assert_close(r1["out"].float(), ou.detach(), dtype, what=f"uni[{r}] out")
        g1 = K.scan_backward(dirs[r:r + 1], r1["ckpt"], channel_last(cot.cuda()), delta_softplus=False)
        assert_close(g1["du"][0].float(), lu["u"].grad, dtype, what=f"uni[{r}] du")
        assert_close(g1["ddelta"][0].float(), lu["delta"].grad, dtype, what=f"uni[{r}] ddelta")
        assert_close(g1["dB"][0].float(), lu["B"].grad, dtype, floor="max", what=f"uni[{r}] dB")
        assert_close(g1["dC"][0].float(), lu["C"].grad, dtype, floor="max", what=f"uni[{r}] dC")
        assert_close(g1["dD"][0], lu["D"].grad, dtype, floor="max", what=f"uni[{r}] dD")
        if L > 1:
            assert_close(g1["dA"][0], lu["A"].grad, dtype, floor="max", what=f"uni[{r}] dA")


def test_tma_backward_is_deterministic_and_agrees_with_state_parallel(monkeypatch):
    """Two launches of scan_bwd_lc.cu are bit-identical (fixed-order sums, no atomics); the state-parallel kernel on the same
    tensors agrees; the library reports 128-channel slabs for this layout and 32-channel slabs otherwise."""
    import ctypes as C
    from mamba_asr_b200 import _cabi
    monkeypatch.setenv("CM_SCAN_LC_BWD", "1")
    K, f, bw, dirs, zc, cot = _bidir_backward_case(2, 256, 77, torch.bfloat16, seed=151, layout="xdbl")
    res = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    run = lambda d_: K.scan_backward(d_, res["ckpt"], channel_last(cot.cuda()), z=zc, out_pre=res["out_pre"], out_scale=0.5,
                                     delta_softplus=True)
    a, b = run(dirs), run(dirs)
    for key in ("du", "ddelta", "dB", "dC", "dA", "dD", "dbias"):
        for r in range(2):
            assert torch.equal(a[key][r], b[key][r]), (key, r)
    assert torch.equal(a["dz"], b["dz"])
    monkeypatch.setenv("CM_SCAN_NO_LC", "1")
    o = run(dirs)
    for key in ("du", "ddelta", "dB", "dC"):
        for r in range(2):
            assert_close(a[key][r].float(), o[key][r].float(), torch.bfloat16, floor="max", what=f"{key}[{r}] lc vs sp")
    for key in ("dA", "dD", "dbias"):
        for r in range(2):
            assert_close(a[key][r], o[key][r], torch.bfloat16, floor="max", what=f"{key}[{r}] lc vs sp")
    assert_close(a["dz"].float(), o["dz"].float(), torch.bfloat16, floor="max", what="dz lc vs sp")


def test_tma_backward_full_size_config3(monkeypatch):
    """ConMamba-large shape (B 64, D 512, L 501, bf16), x_dbl layout: the TMA backward against the state-parallel kernel."""
    from mamba_asr_b200 import kernels as K
    Bt, D, L, N = 64, 512, 501, 16
    g = torch.Generator(device="cuda").manual_seed(6)
    rn = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
    cl = lambda: rn(Bt, L, D).bfloat16().transpose(1, 2)
    z = cl()
    dirs = []
    for rev in (False, True):
        xd = rn(Bt, L, 48).bfloat16()
        dirs.append(dict(u=cl(), delta=(0.5 * rn(Bt, L, D)).bfloat16().transpose(1, 2), A=-torch.exp(0.3 * rn(D, N)),
                         B=xd[..., :N].transpose(1, 2), C=xd[..., N:2 * N].transpose(1, 2), D=torch.ones(D, device="cuda"),
                         delta_bias=torch.full((D,), -4.0, device="cuda"), reverse=rev))
    res = K.scan_forward(dirs, z=z, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    go = cl()
    run = lambda: K.scan_backward(dirs, res["ckpt"], go, z=z, out_pre=res["out_pre"], out_scale=0.5, delta_softplus=True)
    monkeypatch.setenv("CM_SCAN_LC_BWD", "1")
    a = run()
    monkeypatch.setenv("CM_SCAN_NO_LC", "1")
    o = run()
    assert_close(a["dz"].float(), o["dz"].float(), torch.bfloat16, floor="max", what="dz")
    for r in range(2):
        for key in ("du", "ddelta", "dB", "dC"):
            assert_close(a[key][r].float(), o[key][r].float(), torch.bfloat16, floor="max", what=f"{key}[{r}]")
        for key in ("dA", "dD", "dbias"):
            assert_close(a[key][r], o[key][r], torch.bfloat16, floor="max", what=f"{key}[{r}]")


# ---- scan_bwd_wg.cu: warpgroup-specialised backward (setmaxnreg roles, TMA operands, tensor-core state sums) ------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 64, 45), (1, 64, 8), (2, 96, 9), (1, 64, 1), (2, 288, 17), (1, 128, 67), (2, 64, 16),
                                   (3, 160, 131), (1, 512, 24), (2, 1024, 19)])
def test_warpgroup_backward_matches_autograd(dtype, shape, monkeypatch):
    """scan_bwd_wg.cu (the default backward for the module's x_dbl layout once the grid fills the GPU; forced here): every
    gradient of the fused bidirectional block and of unidirectional scans (with and without gate / softplus) against
    autograd through the oracle (fed the dtype-rounded inputs, computed in fp32) - ragged lengths, tile boundaries, slabs
    that are half empty (dim = 96, 160, 288) and the benchmark widths 288 / 512 / 1024."""
    from oracle.scan_ref import selective_scan_oracle
    monkeypatch.setenv("CM_SCAN_WG", "1")
    Bt, D, L = shape
    K, f, bw, dirs, zc, cot = _bidir_backward_case(Bt, D, L, dtype, seed=171, layout="xdbl")
    ref, lf, lb = _oracle_bidir_grads(f, bw, cot, dtype)
    res = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    g = K.scan_backward(dirs, res["ckpt"], channel_last(cot.cuda()), z=zc, out_pre=res["out_pre"], out_scale=0.5,
                        delta_softplus=True)
    _check_bidir_grads(g, lf, lb, dtype, "wg bidir")
    for r, src in ((0, f), (1, bw)):
        lu = {k: v.float().clone().requires_grad_(True) for k, v in src.items() if k != "z"}
        fl = (lambda t: t.flip(-1)) if r == 1 else (lambda t: t)
        ou = fl(selective_scan_oracle(fl(lu["u"]), fl(lu["delta"]), lu["A"], fl(lu["B"]), fl(lu["C"]), lu["D"], None,
                                      lu["delta_bias"], False))
        (ou * cot.float()).sum().backward()
        r1 = K.scan_forward(dirs[r:r + 1], delta_softplus=False, need_ckpt=True)
        g1 = K.scan_backward(dirs[r:r + 1], r1["ckpt"], channel_last(cot.cuda()), delta_softplus=False)
        assert_close(g1["du"][0].float(), lu["u"].grad, dtype, what=f"uni[{r}] du")
        assert_close(g1["ddelta"][0].float(), lu["delta"].grad, dtype, what=f"uni[{r}] ddelta")
        assert_close(g1["dB"][0].float(), lu["B"].grad, dtype, floor="max", what=f"uni[{r}] dB")
        assert_close(g1["dC"][0].float(), lu["C"].grad, dtype, floor="max", what=f"uni[{r}] dC")
        assert_close(g1["dD"][0], lu["D"].grad, dtype, floor="max", what=f"uni[{r}] dD")
        if L > 1:
            assert_close(g1["dA"][0], lu["A"].grad, dtype, floor="max", what=f"uni[{r}] dA")


def test_warpgroup_backward_is_deterministic_and_sized_by_the_library(monkeypatch):
    """Two launches are bit-identical (fixed-order sums, no atomics); the state-parallel kernel on the same tensors agrees;
    the library reports 64-channel slabs when it will take the warpgroup kernel and 32-channel slabs otherwise (small grids
    stay on scan_bwd_sp.cu unless forced)."""
    import ctypes as C
    from mamba_asr_b200 import _cabi
    K, f, bw, dirs, zc, cot = _bidir_backward_case(2, 288, 77, torch.bfloat16, seed=181, layout="xdbl")
    res = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    run = lambda: K.scan_backward(dirs, res["ckpt"], channel_last(cot.cuda()), z=zc, out_pre=res["out_pre"], out_scale=0.5,
                                  delta_softplus=True)
    o = run()                                   # 2 x 5 x 2 CTAs: below a wave, the library keeps the state-parallel kernel
    monkeypatch.setenv("CM_SCAN_WG", "1")
    a, b = run(), run()
    for key in ("du", "ddelta", "dB", "dC", "dA", "dD", "dbias"):
        for r in range(2):
            assert torch.equal(a[key][r], b[key][r]), (key, r)
    assert torch.equal(a["dz"], b["dz"])
    for key in ("du", "ddelta", "dB", "dC"):
        for r in range(2):
            assert_close(a[key][r].float(), o[key][r].float(), torch.bfloat16, floor="max", what=f"{key}[{r}] wg vs sp")
    for key in ("dA", "dD", "dbias"):
        for r in range(2):
            assert_close(a[key][r], o[key][r], torch.bfloat16, floor="max", what=f"{key}[{r}] wg vs sp")
    assert_close(a["dz"].float(), o["dz"].float(), torch.bfloat16, floor="max", what="dz wg vs sp")


def test_warpgroup_backward_full_size_config3_is_the_default(monkeypatch):
    """ConMamba-large shape (B 64, D 512, L 501, bf16), x_dbl layout: 1024 CTAs, so the library takes the warpgroup kernel
    without being asked; it agrees with the state-parallel kernel (CM_SCAN_NO_WG=1) and satisfies the size-independent
    property that the gradient is linear in dout."""
    from mamba_asr_b200 import kernels as K
    Bt, D, L, N = 64, 512, 501, 16
    g = torch.Generator(device="cuda").manual_seed(7)
    rn = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
    cl = lambda: rn(Bt, L, D).bfloat16().transpose(1, 2)
    z = cl()
    dirs = []
    for rev in (False, True):
        xd = rn(Bt, L, 64).bfloat16()
        dirs.append(dict(u=cl(), delta=(0.5 * rn(Bt, L, D)).bfloat16().transpose(1, 2), A=-torch.exp(0.3 * rn(D, N)),
                         B=xd[..., :N].transpose(1, 2), C=xd[..., N:2 * N].transpose(1, 2), D=torch.ones(D, device="cuda"),
                         delta_bias=torch.full((D,), -4.0, device="cuda"), reverse=rev))
    res = K.scan_forward(dirs, z=z, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    go = cl()
    run = lambda go_: K.scan_backward(dirs, res["ckpt"], go_, z=z, out_pre=res["out_pre"], out_scale=0.5, delta_softplus=True)
    a = run(go)
    a2 = run((2.0 * go.float()).bfloat16())     # exact in bf16: every gradient doubles (up to the rounding of the outputs)
    monkeypatch.setenv("CM_SCAN_NO_WG", "1")
    o = run(go)
    assert_close(a["dz"].float(), o["dz"].float(), torch.bfloat16, floor="max", what="dz")
    for r in range(2):
        for key in ("du", "ddelta", "dB", "dC"):
            assert_close(a[key][r].float(), o[key][r].float(), torch.bfloat16, floor="max", what=f"{key}[{r}]")
            assert_close(a2[key][r].float(), 2.0 * a[key][r].float(), torch.bfloat16, floor="max", what=f"linear {key}[{r}]")
        for key in ("dA", "dD", "dbias"):
            assert_close(a[key][r], o[key][r], torch.bfloat16, floor="max", what=f"{key}[{r}]")
            assert_close(a2[key][r], 2.0 * a[key][r], torch.bfloat16, floor="max", what=f"linear {key}[{r}]")


# ---- scan_fwd_wg.cu: warpgroup-specialised forward (opt-in: CM_SCAN_WG_FWD=1) ---------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 64, 45), (1, 32, 16), (2, 96, 9), (1, 64, 1), (2, 288, 17), (1, 128, 67), (3, 160, 131),
                                   (1, 512, 33), (2, 1024, 19)])
def test_warpgroup_forward_matches_oracle(dtype, shape, monkeypatch):
    """scan_fwd_wg.cu: fused bidirectional output, pre-gate sum, checkpoints (through the backward pass that consumes them) and
    unidirectional launches with last_state, against the oracle on the dtype-rounded inputs - ragged lengths, tile and
    range boundaries, an odd number of 32-channel blocks in the unidirectional grid (dim = 96, 160, 288)."""
    from oracle.scan_ref import selective_scan_oracle
    monkeypatch.setenv("CM_SCAN_WG_FWD", "1")
    Bt, D, L = shape
    K, f, bw, dirs, zc, cot = _bidir_backward_case(Bt, D, L, dtype, seed=191, layout="xdbl")
    ref, lf, lb = _oracle_bidir_grads(f, bw, cot, dtype)
    res = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    assert_close(res["out"].float(), ref.detach(), dtype, what="wg fwd out")
    g = K.scan_backward(dirs, res["ckpt"], channel_last(cot.cuda()), z=zc, out_pre=res["out_pre"], out_scale=0.5,
                        delta_softplus=True)
    _check_bidir_grads(g, lf, lb, dtype, "wg fwd ckpt -> bwd")
    monkeypatch.delenv("CM_SCAN_WG_FWD")
    o = K.scan_forward(dirs, z=zc, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    monkeypatch.setenv("CM_SCAN_WG_FWD", "1")
    assert_close(res["out_pre"].float(), o["out_pre"].float(), dtype, floor="max", what="out_pre wg vs sp")
    for r in range(2):
        assert_close(res["ckpt"][r], o["ckpt"][r], dtype, floor="max", what="ckpt[%d] wg vs sp" % r)
    for r, src in ((0, f), (1, bw)):
        fl = (lambda t: t.flip(-1)) if r == 1 else (lambda t: t)
        s = {k: v.float() for k, v in src.items() if k != "z"}
        ou, last = selective_scan_oracle(fl(s["u"]), fl(s["delta"]), s["A"], fl(s["B"]), fl(s["C"]), s["D"], None, s["delta_bias"],
                                         False, return_last_state=True)
        r1 = K.scan_forward(dirs[r:r + 1], delta_softplus=False, need_last_state=True)
        assert_close(r1["out"].float(), fl(ou), dtype, what=f"uni[{r}] out")
        assert_close(r1["last_state"][0], last, dtype, floor="max", what=f"uni[{r}] last_state")


def test_warpgroup_forward_full_size_config3(monkeypatch):
    """ConMamba-large shape (B 64, D 512, L 501, bf16): bit-identical across launches and equal to the state-parallel kernel."""
    from mamba_asr_b200 import kernels as K
    Bt, D, L, N = 64, 512, 501, 16
    g = torch.Generator(device="cuda").manual_seed(8)
    rn = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
    cl = lambda: rn(Bt, L, D).bfloat16().transpose(1, 2)
    z = cl()
    dirs = []
    for rev in (False, True):
        xd = rn(Bt, L, 64).bfloat16()
        dirs.append(dict(u=cl(), delta=(0.5 * rn(Bt, L, D)).bfloat16().transpose(1, 2), A=-torch.exp(0.3 * rn(D, N)),
                         B=xd[..., :N].transpose(1, 2), C=xd[..., N:2 * N].transpose(1, 2), D=torch.ones(D, device="cuda"),
                         delta_bias=torch.full((D,), -4.0, device="cuda"), reverse=rev))
    run = lambda: K.scan_forward(dirs, z=z, out_scale=0.5, delta_softplus=True, need_ckpt=True, need_out_pre=True)
    o = run()
    monkeypatch.setenv("CM_SCAN_WG_FWD", "1")
    a, b = run(), run()
    assert torch.equal(a["out"], b["out"]) and torch.equal(a["ckpt"][0], b["ckpt"][0]) and torch.equal(a["ckpt"][1], b["ckpt"][1])
    assert_close(a["out"].float(), o["out"].float(), torch.bfloat16, floor="max", what="out")
    assert_close(a["out_pre"].float(), o["out_pre"].float(), torch.bfloat16, floor="max", what="out_pre")
    for r in range(2):
        assert_close(a["ckpt"][r], o["ckpt"][r], torch.bfloat16, floor="max", what="ckpt[%d]" % r)
