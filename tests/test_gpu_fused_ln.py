"""GPU parity of the fused residual-add + dropout + LayerNorm kernels (cm_add_ln_fwd / cm_add_ln_bwd) against the plain
torch fp32 composition  s = a + alpha * dropout(b) ; y = layer_norm(s)  (reference modules/Conmamba.py:638-649)."""
import pytest
import torch
import torch.nn.functional as F

from cm_testutil import assert_close

pytestmark = pytest.mark.gpu

COMBOS = [(torch.float32, torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16, torch.float32),
          (torch.float32, torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16, torch.bfloat16),
          (torch.bfloat16, torch.bfloat16, torch.float32)]


def _inputs(shape, ta, tb, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(*shape, generator=g).to(ta)
    b = torch.randn(*shape, generator=g).to(tb)
    C = shape[-1]
    w, bias = 1.0 + 0.2 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    cs, cy = torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)
    return a, b, w, bias, cs, cy


@pytest.mark.parametrize("combo", COMBOS)
@pytest.mark.parametrize("shape", [(3, 37, 144), (2, 501, 256), (1, 1, 512), (5, 3, 34), (2, 9, 1024)])
@pytest.mark.parametrize("use_s", [True, False])
def test_add_layer_norm_no_dropout_matches_torch(combo, shape, use_s):
    from mamba_asr_b200.layernorm import FusedLayerNorm, add_dropout_layer_norm
    ta, tb, ty = combo
    a, b, w, bias, cs, cy = _inputs(shape, ta, tb)
    C = shape[-1]
    alpha = 0.5
    # fp32 reference on the rounded inputs; s is rounded to a's dtype before the norm, as the residual stream stores it
    ar, br = a.float().clone().requires_grad_(True), b.float().clone().requires_grad_(True)
    wr, biasr = w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    s_ref = ar + alpha * br
    s_st = s_ref + (s_ref.detach().to(ta).float() - s_ref.detach())          # straight-through rounding
    y_ref = F.layer_norm(s_st, (C,), wr, biasr, 1e-5)
    loss = (y_ref * cy).sum() + ((s_ref * cs).sum() if use_s else 0.0)
    loss.backward()

    norm = FusedLayerNorm(C).cuda()
    norm.keep_dtype = ty == torch.float32
    with torch.no_grad():
        norm.weight.copy_(w)
        norm.bias.copy_(bias)
    ag, bg = a.detach().cuda().requires_grad_(True), b.detach().cuda().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(tb == torch.bfloat16)):
        s, y = add_dropout_layer_norm(ag, bg, norm, alpha=alpha, p_drop=0.0, training=True)
    assert s.dtype == ta and y.dtype == ty, (s.dtype, y.dtype)
    lg = (y.float() * cy.cuda()).sum() + ((s.float() * cs.cuda()).sum() if use_s else 0.0)
    lg.backward()
    assert_close(s.float(), s_ref.detach().to(ta).float(), ta, what="s")
    assert_close(y.float(), y_ref, ty, what="y")
    lo = torch.bfloat16 if torch.bfloat16 in combo else torch.float32
    assert_close(ag.grad.float(), ar.grad, ta if ta == torch.bfloat16 else lo, floor="max", what="da")
    assert_close(bg.grad.float(), br.grad, lo, floor="max", what="db")
    assert_close(norm.weight.grad, wr.grad, lo, floor="max", what="dgamma")
    assert_close(norm.bias.grad, biasr.grad, lo, floor="max", what="dbeta")


@pytest.mark.parametrize("combo", [COMBOS[0], COMBOS[2], COMBOS[3]])
def test_fused_dropout_mask_statistics_and_gradient(combo):
    from mamba_asr_b200 import kernels as K
    from mamba_asr_b200.layernorm import DropoutSeed
    ta, tb, ty = combo
    rows, C, p, alpha = 4096, 256, 0.1, 0.5
    a, b, w, bias, cs, cy = _inputs((rows, C), ta, tb, seed=1)
    a, b, w, bias = a.cuda(), b.cuda(), w.cuda(), bias.cuda()
    seed = DropoutSeed.tensor(a.device)
    s, y, mean, rstd, mask = K.add_ln_forward(a, b, w, bias, 1e-5, alpha, p, seed, 7, ty, store_mask=True)
    s_k, y_k, _, _, key = K.add_ln_forward(a, b, w, bias, 1e-5, alpha, p, seed, 7, ty)     # default: nothing stored but the key
    assert key.dtype == torch.int32 and key.numel() == 1 and torch.equal(s, s_k) and torch.equal(y, y_k)
    keep = mask.float()
    assert set(mask.unique().tolist()) <= {0, 1}
    assert abs(float(keep.mean()) - (1 - p)) < 3e-3                               # 1M samples: sigma = 3e-4
    assert float(keep.mean(0).min()) > 0.85 and float(keep.mean(1).min()) > 0.75    # no dead rows / columns
    s_ref = a.float() + alpha / (1 - p) * keep * b.float()
    assert_close(s.float(), s_ref.to(ta).float(), ta, what="s with dropout")
    y_ref = F.layer_norm(s_ref.to(ta).float(), (C,), w, bias, 1e-5)
    assert_close(y.float(), y_ref, ty, what="y with dropout")
    # same (seed, call id) -> same mask; another call id or an advanced seed -> a different one
    m2 = K.add_ln_forward(a, b, w, bias, 1e-5, alpha, p, seed, 7, ty, store_mask=True)[4]
    m3 = K.add_ln_forward(a, b, w, bias, 1e-5, alpha, p, seed, 8, ty, store_mask=True)[4]
    assert torch.equal(mask, m2) and not torch.equal(mask, m3)
    assert abs(float((mask == m3).float().mean()) - (p * p + (1 - p) ** 2)) < 5e-3  # independent masks
    DropoutSeed.advance(a.device)
    m4 = K.add_ln_forward(a, b, w, bias, 1e-5, alpha, p, seed, 7, ty, store_mask=True)[4]
    assert not torch.equal(mask, m4)
    # backward: db = alpha / (1 - p) * keep * da
    dy = cy.cuda().to(ty)
    ds = cs.cuda().to(ta)
    da, db, dg, dbt = K.add_ln_backward(s, dy, ds, w, mean, rstd, mask, alpha, p, tb)
    assert_close(db.float(), (alpha / (1 - p)) * keep * da.float(), tb, floor="max", what="db vs mask * da")
    # the mask regenerated from the forward's key (captured before the seed advanced) gives the same gradients, and the
    # column sums of db (the bias gradient of the Linear behind b) come out of the same launch
    da_k, db_k, dg_k, dbt_k, dbs = K.add_ln_backward(s, dy, ds, w, mean, rstd, key, alpha, p, tb, need_dbsum=True)
    assert torch.equal(da, da_k) and torch.equal(db, db_k) and torch.equal(dg, dg_k) and torch.equal(dbt, dbt_k)
    assert_close(dbs, db.double().sum(0).float(), torch.float32 if tb == torch.float32 else torch.bfloat16, floor="max",
                 what="column sums of db")
    sr = s.float().requires_grad_(True)
    (F.layer_norm(sr, (C,), w, bias, 1e-5) * dy.float()).sum().backward()
    assert_close(da.float(), sr.grad + ds.float(), ta, floor="max", what="da")


def test_encoder_layer_fused_path_equals_unfused_ops():
    """The layer with the fused add+norm kernels against the same layer on the separate ops (dropout off), fp32 and
    bf16 autocast, forward and backward."""
    from mamba_asr_b200.conmamba import ConmambaEncoderLayer
    torch.manual_seed(0)
    layer = ConmambaEncoderLayer(d_model=64, d_ffn=128, activation=torch.nn.GELU, dropout=0.0,
                                 mamba_config=dict(d_state=16, expand=2, d_conv=4, bidirectional=True)).cuda()
    x = torch.randn(2, 45, 64, device="cuda")
    cot = torch.randn(2, 45, 64, device="cuda")
    for autocast in (False, True):
        outs, grads = [], []
        for fused in (True, False):
            layer.fuse_add_norm = fused
            layer.zero_grad(set_to_none=True)
            xi = x.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                o = layer(xi)
            (o.float() * cot).sum().backward()
            outs.append(o.float())
            grads.append([xi.grad] + [p.grad.clone() for p in layer.parameters()])
        dt = torch.bfloat16 if autocast else torch.float32
        assert_close(outs[0], outs[1], dt, what="layer out (autocast=%s)" % autocast)
        for g0, g1 in zip(grads[0], grads[1]):
            assert_close(g0, g1, dt, floor="max", what="layer grads (autocast=%s)" % autocast, rtol_mul=2.0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(3, 37, 1024), (2, 5, 8), (1, 501, 576)])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_gelu_dropout_matches_torch(dtype, shape, p):
    from mamba_asr_b200 import kernels as K
    from mamba_asr_b200.layernorm import DropoutSeed, gelu_dropout
    g = torch.Generator().manual_seed(5)
    x = (2.0 * torch.randn(*shape, generator=g)).to(dtype)
    cot = torch.randn(*shape, generator=g).to(dtype)
    xg = x.cuda().requires_grad_(True)
    y = gelu_dropout(xg, p, training=True)
    assert y.dtype == dtype
    y.backward(cot.cuda())
    xr = x.float().clone().requires_grad_(True)
    yr = F.gelu(xr)
    if p > 0:
        keep = (y != 0) | (yr.detach().cuda() == 0)          # the mask, read off the output
        frac = float(keep.float().mean())
        if x.numel() > 10000:
            assert abs(frac - (1 - p)) < 0.01, frac
        yr = yr * keep.cpu().float() / (1 - p)
    (yr * cot.float()).sum().backward()
    assert_close(y.float(), yr, dtype, what="gelu_dropout out")
    assert_close(xg.grad.float(), xr.grad, dtype, floor="max", what="gelu_dropout dx")
    # eval mode: no dropout
    assert_close(gelu_dropout(x.cuda(), 0.5, training=False).float(), F.gelu(x.float()), dtype, what="eval")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(12032, 288, 16), (32064, 512, 48), (501, 64, 8), (1, 8, 8), (1000, 1024, 64),
                                   (257, 72, 40), (256, 96, 24)])
def test_tall_skinny_weight_gradient_gemm_matches_fp32_reference(dtype, shape):
    """cm_tsmm (A^T B over batch * L rows; the x_proj / dt_proj weight gradients of
    selective_scan_interface.py:277-283) against an fp64 product of the same 16-bit inputs; strided B as in the module."""
    from mamba_asr_b200 import kernels as K
    rows, M, N = shape
    g = torch.Generator().manual_seed(rows + M + N)
    a = torch.randn(rows, M, generator=g).to(dtype).cuda()
    wide = torch.randn(rows, N + 32, generator=g).to(dtype).cuda()
    b = wide[:, 32:]                                              # leading dimension N + 32, base offset 64 bytes
    assert K.tsmm_supported(a, b)
    out = K.tsmm(a, b)
    ref = (a.double().t() @ b.double()).float()
    assert out.shape == (M, N) and out.dtype == torch.float32
    assert_close(out, ref, torch.float32, floor="max", what="tsmm", rtol_mul=2.0)
    out2 = K.tsmm(a, b)
    assert torch.equal(out, out2)                                  # deterministic


def test_mamba_weight_gradients_same_with_and_without_tsmm(monkeypatch):
    from mamba_asr_b200 import Mamba
    from mamba_asr_b200 import mamba_inner
    monkeypatch.setattr(mamba_inner, "_USE_TSMM", True)
    monkeypatch.setattr(mamba_inner, "_TSMM_MAX_ELEMS", 1 << 40)
    torch.manual_seed(1)
    m = Mamba(d_model=64, bimamba_type="v2").cuda()
    x = torch.randn(3, 70, 64, device="cuda")
    grads = []
    for off in (False, True):
        if off:
            monkeypatch.setenv("CM_NO_TSMM", "1")
        m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x)
        y.float().square().mean().backward()
        grads.append({n: p.grad.clone() for n, p in m.named_parameters()})
    for n in grads[0]:
        assert_close(grads[0][n], grads[1][n], torch.bfloat16, floor="max", what="d " + n)


# ---------------------------------------------------------------------------------------------------------------------
# wide-row LayerNorm + LeakyReLU (cm_ln_act_fwd / cm_ln_act_bwd): the conv blocks of the CNN front-end
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32, torch.float16])
@pytest.mark.parametrize("shape", [(2, 37, 40, 64), (3, 19, 20, 32), (1, 1, 1, 4), (5, 3, 9, 12), (2, 7, 10, 128),
                                   (1, 530, 20, 32), (1, 700, 40, 64)])
@pytest.mark.parametrize("act", ["leaky_relu", "gelu"])
@pytest.mark.parametrize("with_pre_bias", [False, True])
def test_layer_norm_leaky_relu_matches_torch(dtype, shape, act, with_pre_bias):
    """act(LayerNorm([F, C])(x + pre_bias)) and all gradients against the torch fp32 composition on the same rounded
    inputs (rows of 4 .. 2560 elements: every thread-group width of the kernel, ragged last row block, > 1 wave)."""
    from mamba_asr_b200.layernorm import layer_norm_act
    layer_norm_leaky_relu = lambda x_, n_: layer_norm_act(x_, n_, act, 0.01, pbg)
    g = torch.Generator().manual_seed(sum(shape))
    Fd, Cc = shape[-2:]
    x = torch.randn(*shape, generator=g).to(dtype)
    w, b = 1.0 + 0.2 * torch.randn(Fd, Cc, generator=g), 0.1 * torch.randn(Fd, Cc, generator=g)
    cy = torch.randn(*shape, generator=g).to(dtype)

    pb = 0.5 * torch.randn(Cc, generator=g)
    pbr = pb.clone().requires_grad_(True)
    pbg = pb.cuda().requires_grad_(True) if with_pre_bias else None

    xr = x.float().clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    pre = F.layer_norm(xr + pbr if with_pre_bias else xr, (Fd, Cc), wr, br, 1e-5)
    y_ref = F.leaky_relu(pre) if act == "leaky_relu" else F.gelu(pre)
    (y_ref * cy.float()).sum().backward()

    norm = torch.nn.LayerNorm([Fd, Cc]).cuda()
    with torch.no_grad():
        norm.weight.copy_(w)
        norm.bias.copy_(b)
    xg = x.cuda().requires_grad_(True)
    y = layer_norm_leaky_relu(xg, norm)
    assert y.dtype == dtype and y.shape == xg.shape
    (y.float() * cy.cuda().float()).sum().backward()
    assert_close(y.float(), y_ref, dtype, what="y")
    # an element whose pre-activation is within rounding of zero may take the other LeakyReLU branch: not compared
    safe = (pre.detach().abs() > 1e-5).float() if act == "leaky_relu" else torch.ones_like(pre)
    assert float(safe.mean()) > 0.99
    if with_pre_bias:
        assert_close(pbg.grad, pbr.grad, dtype, floor="max", what="d pre_bias", rtol_mul=2.0)
    assert_close(xg.grad.float().cpu() * safe, xr.grad * safe, dtype, floor="max", what="dx")
    assert xg.grad.dtype == dtype
    assert_close(norm.weight.grad, wr.grad, dtype, floor="max", what="dgamma", rtol_mul=2.0)
    assert_close(norm.bias.grad, br.grad, dtype, floor="max", what="dbeta", rtol_mul=2.0)
    # deterministic (fixed-order partial sums, no atomics)
    xg2 = x.cuda().requires_grad_(True)
    g1 = norm.weight.grad.clone()
    norm.zero_grad()
    (layer_norm_leaky_relu(xg2, norm).float() * cy.cuda().float()).sum().backward()
    assert torch.equal(xg2.grad, xg.grad) and torch.equal(norm.weight.grad, g1)


def test_layer_norm_leaky_relu_refuses_what_the_kernel_does_not_cover():
    from mamba_asr_b200.layernorm import layer_norm_leaky_relu
    norm = torch.nn.LayerNorm([3, 2]).cuda()                      # 6 elements: not a multiple of 4
    with pytest.raises(NotImplementedError):
        layer_norm_leaky_relu(torch.randn(2, 3, 2, device="cuda"), norm)
    wide = torch.nn.LayerNorm([41, 64]).cuda()                    # 2624 > 2560
    with pytest.raises(NotImplementedError):
        layer_norm_leaky_relu(torch.randn(2, 41, 64, device="cuda"), wide)
    with pytest.raises(RuntimeError):
        layer_norm_leaky_relu(torch.randn(2, 4, 4), torch.nn.LayerNorm([4, 4]))   # CPU tensor: no fallback


@pytest.mark.parametrize("autocast", [False, True])
def test_conv_front_end_same_with_and_without_the_fused_norm(autocast, monkeypatch):
    """ConvFrontEnd (2 x [conv 3x3 stride 2, LayerNorm([F', C]), LeakyReLU]) on the fused kernel against the same module
    on the two torch ops: output and every parameter gradient."""
    from mamba_asr_b200.encoder import ConvFrontEnd
    # fp32 case in true fp32: a TF32 second conv turns 1e-7 differences of the norm output into 5e-4 ones (input rounding)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    torch.manual_seed(3)
    fe = ConvFrontEnd(80).cuda()
    with torch.no_grad():
        for n in fe.norms:
            n.weight.add_(0.1 * torch.randn_like(n.weight))
            n.bias.add_(0.1 * torch.randn_like(n.bias))
    feats = torch.randn(3, 203, 80, device="cuda")
    cy = torch.randn(3, 51, fe.out_features, device="cuda")
    res = []
    for fused in (True, False):
        fe.use_kernel = fused
        fe.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = fe(feats)
        (out.float() * cy).sum().backward()
        res.append((out.float().detach(), [p.grad.clone() for p in fe.parameters()]))
    dt = torch.bfloat16 if autocast else torch.float32
    assert res[0][0].shape == (3, 51, fe.out_features)
    # autocast: block 1 keeps its conv output in fp32 registers (cm_stem_fwd) where cuDNN rounds it to bf16 before the norm
    assert_close(res[0][0], res[1][0], dt, what="front-end output", rtol_mul=2.0 if autocast else 1.0)
    for (name, _), ga, gb in zip(fe.named_parameters(), res[0][1], res[1][1]):
        if autocast:
            # the fused path adds the conv bias in fp32, cuDNN rounds x + bias to bf16 first: pre-activations within
            # ~1e-2 of zero take the other LeakyReLU branch (a 0.99 * dy step in single terms of the sums), so the
            # bf16 gradients are compared in norm rather than element by element
            # (block 1 on cm_stem_* also keeps the conv output itself in fp32: a few more flips, 0.051 measured)
            rel = float((ga - gb).norm() / gb.norm())
            assert rel < 0.08, (name, rel)
        else:
            assert_close(ga, gb, dt, floor="max", what="d " + name, rtol_mul=4.0)


@pytest.mark.parametrize("autocast", [False, True])
def test_convolution_module_same_with_and_without_the_fused_norm_gelu(autocast, monkeypatch):
    """ConvolutionModule (reference modules/Conmamba.py:182-454) with LayerNorm -> GELU after the depthwise conv as the
    GELU epilogue of cm_layernorm_* (the default) and on cm_ln_act_* (opt-in, CM_FUSE_LN_GELU=1), each against the separate
    cm_layernorm + torch GELU evaluation (CM_NO_LN_GELU_EPILOGUE=1): output and every parameter gradient."""
    from mamba_asr_b200.conmamba import ConvolutionModule
    torch.manual_seed(5)
    m = ConvolutionModule(144, kernel_size=31, activation=torch.nn.GELU, dropout=0.0).cuda()
    x = torch.randn(3, 77, 144, device="cuda")
    cy = torch.randn(3, 77, 144, device="cuda")
    res = []
    for mode in ("ln_act", "epilogue", "separate"):
        monkeypatch.delenv("CM_FUSE_LN_GELU", raising=False)
        monkeypatch.delenv("CM_NO_LN_GELU_EPILOGUE", raising=False)
        if mode == "ln_act":
            monkeypatch.setenv("CM_FUSE_LN_GELU", "1")
        elif mode == "separate":
            monkeypatch.setenv("CM_NO_LN_GELU_EPILOGUE", "1")
        m.zero_grad()
        xg = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = m(xg)
        (out.float() * cy).sum().backward()
        res.append((out.float().detach(), xg.grad.clone(), [p.grad.clone() for p in m.parameters()]))
    dt = torch.bfloat16 if autocast else torch.float32
    for k, what in ((0, "cm_ln_act"), (1, "epilogue")):
        assert_close(res[k][0], res[2][0], dt, what=what + ": conv module output")
        assert_close(res[k][1], res[2][1], dt, floor="max", what=what + ": conv module dx", rtol_mul=2.0)
        for (name, _), ga, gb in zip(m.named_parameters(), res[k][2], res[2][2]):
            assert_close(ga, gb, dt, floor="max", what=what + ": d " + name, rtol_mul=4.0)


@pytest.mark.parametrize("shape", [(3, 67, 256), (2, 501, 144), (5, 64), (1, 1, 16), (4, 33, 1024), (2, 9, 288)])
@pytest.mark.parametrize("mode", ["fp32", "bf16", "autocast"])
def test_layer_norm_gelu_epilogue_matches_torch(shape, mode):
    """cm_layernorm_fwd / _bwd with the GELU epilogue (cm_layernorm_args.act = CM_LN_OUT_GELU; the LayerNorm -> GELU pair of
    the convolution module, reference modules/Conmamba.py:292-301) against F.gelu(F.layer_norm(x)) in fp32 on the same
    rounded input: y, dx, dgamma, dbeta; one kernel each way."""
    from mamba_asr_b200 import kernels as K
    from mamba_asr_b200.layernorm import FusedLayerNorm, layer_norm_gelu_supported
    g = torch.Generator().manual_seed(31)
    Cn = shape[-1]
    dt = torch.float32 if mode != "bf16" else torch.bfloat16
    x = (1.5 * torch.randn(*shape, generator=g) + 0.3).to(dt).cuda()
    cot = torch.randn(*shape, generator=g).cuda()
    norm = FusedLayerNorm(Cn).cuda()
    with torch.no_grad():
        norm.weight.copy_(1.0 + 0.2 * torch.randn(Cn, generator=g))
        norm.bias.copy_(0.3 * torch.randn(Cn, generator=g))
    assert layer_norm_gelu_supported(x)
    xr = x.float().clone().requires_grad_(True)
    wr, br = norm.weight.detach().clone().requires_grad_(True), norm.bias.detach().clone().requires_grad_(True)
    ref = F.gelu(F.layer_norm(xr, (Cn,), wr, br, norm.eps))
    (ref * cot).sum().backward()
    xk = x.clone().requires_grad_(True)
    l0 = K.LAUNCHES
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "autocast"):
        out = norm(xk, gelu=True)
    (out.float() * cot).sum().backward()
    assert K.LAUNCHES - l0 <= 3                                # forward, backward, the dgamma / dbeta reduction
    odt = torch.bfloat16 if mode != "fp32" else torch.float32
    assert out.dtype == odt
    assert_close(out.float(), ref, odt, what="ln+gelu y")
    assert_close(xk.grad.float(), xr.grad, odt, floor="max", what="ln+gelu dx")
    assert_close(norm.weight.grad, wr.grad, odt, floor="max", what="ln+gelu dgamma")
    assert_close(norm.bias.grad, br.grad, odt, floor="max", what="ln+gelu dbeta")


# ---------------------------------------------------------------------------------------------------------------------
# front-end block 1 as one kernel each way (cm_stem_fwd / cm_stem_bwd): Conv2d(1 -> C, 3 x 3, stride 2, padding 1) +
# LayerNorm([F', C]) + LeakyReLU against the three torch ops in fp32 (reference hparams/CTC/conmamba_large.yaml:187-199)
def _stem_reference(feats, conv, norm, slope=0.01):
    x = F.conv2d(feats.float()[:, None], conv.weight, conv.bias, conv.stride, conv.padding)     # (B, C, T', F')
    return F.leaky_relu(norm(x.permute(0, 2, 3, 1)), slope)


@pytest.mark.parametrize("shape", [(3, 203, 80, 64), (2, 50, 80, 64), (1, 1, 80, 64), (2, 37, 81, 32), (4, 64, 40, 128),
                                   (2, 19, 7, 16), (5, 2, 80, 64), (1, 700, 80, 64), (2, 33, 160, 32)])
@pytest.mark.parametrize("in_dtype", [torch.float32, torch.bfloat16])
def test_stem_matches_conv_layernorm_leaky_relu(shape, in_dtype, monkeypatch):
    from mamba_asr_b200.layernorm import conv_ln_act_stem, conv_ln_act_stem_supported
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    Bt, T, Fd, Cn = shape
    torch.manual_seed(Bt * 7 + T + Cn)
    conv = torch.nn.Conv2d(1, Cn, 3, stride=2, padding=1).cuda()
    norm = torch.nn.LayerNorm([(Fd - 1) // 2 + 1, Cn]).cuda()
    with torch.no_grad():
        norm.weight.add_(0.2 * torch.randn_like(norm.weight))
        norm.bias.add_(0.2 * torch.randn_like(norm.bias))
    feats = torch.randn(Bt, T, Fd, device="cuda").to(in_dtype)
    assert conv_ln_act_stem_supported(feats, conv, norm)
    y = conv_ln_act_stem(feats, conv, norm)
    cy = torch.randn(y.shape, device="cuda")
    (y.float() * cy).sum().backward()
    got = [y.float().detach()] + [p.grad.clone() for p in list(conv.parameters()) + list(norm.parameters())]
    conv.zero_grad(); norm.zero_grad()
    ref = _stem_reference(feats, conv, norm)
    (ref * cy.to(ref.dtype)).sum().backward()
    want = [ref.detach()] + [p.grad.clone() for p in list(conv.parameters()) + list(norm.parameters())]
    assert y.dtype == in_dtype and y.shape == ref.shape
    assert_close(got[0], want[0], in_dtype, what="stem output")
    # bf16 case: y is bf16, so autograd hands the kernel dy = cy rounded to bf16 (the fp32 reference sees cy itself)
    for name, a, b in zip(("dweight", "dbias", "dgamma", "dbeta"), got[1:], want[1:]):
        if in_dtype == torch.float32:
            assert_close(a, b, torch.float32, floor="max", what=name, rtol_mul=4.0)
        else:
            assert_close(a, b, torch.bfloat16, floor="max", what=name)


def test_stem_under_autocast_writes_bf16_and_is_deterministic():
    from mamba_asr_b200.layernorm import conv_ln_act_stem
    torch.manual_seed(1)
    conv = torch.nn.Conv2d(1, 64, 3, stride=2, padding=1).cuda()
    norm = torch.nn.LayerNorm([40, 64]).cuda()
    feats = torch.randn(4, 301, 80, device="cuda")
    outs = []
    for _ in range(2):
        conv.zero_grad(); norm.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = conv_ln_act_stem(feats, conv, norm)
        y.float().square().sum().backward()
        outs.append([y.detach().clone()] + [p.grad.clone() for p in list(conv.parameters()) + list(norm.parameters())])
    assert outs[0][0].dtype == torch.bfloat16 and outs[0][0].shape == (4, 151, 40, 64)
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    ref = _stem_reference(feats, conv, norm)
    assert_close(outs[0][0].float(), ref, torch.bfloat16, what="stem output (autocast)")


def test_stem_envelope_and_ab_switch(monkeypatch):
    from mamba_asr_b200.encoder import ConvFrontEnd
    from mamba_asr_b200.layernorm import conv_ln_act_stem, conv_ln_act_stem_supported
    from mamba_asr_b200 import kernels as K
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    feats = torch.randn(2, 21, 80, device="cuda")
    norm = torch.nn.LayerNorm([40, 64]).cuda()
    assert not conv_ln_act_stem_supported(feats, torch.nn.Conv2d(1, 64, 3, stride=1, padding=1).cuda(), norm)
    assert not conv_ln_act_stem_supported(feats, torch.nn.Conv2d(1, 64, 3, stride=2, padding=1, padding_mode="reflect").cuda(), norm)
    assert not conv_ln_act_stem_supported(feats, torch.nn.Conv2d(1, 48, 3, stride=2, padding=1).cuda(),
                                          torch.nn.LayerNorm([40, 48]).cuda())            # 512 % 48 != 0
    assert not conv_ln_act_stem_supported(feats.requires_grad_(True), torch.nn.Conv2d(1, 64, 3, stride=2, padding=1).cuda(), norm)
    with pytest.raises(RuntimeError):
        conv_ln_act_stem(torch.randn(2, 21, 80), torch.nn.Conv2d(1, 64, 3, stride=2, padding=1), torch.nn.LayerNorm([40, 64]))
    # the front-end takes the kernel by itself; CM_NO_FUSE_STEM=1 is the A/B switch back to cuDNN + cm_ln_act
    fe = ConvFrontEnd(80).cuda()
    x = torch.randn(2, 101, 80, device="cuda")
    n0 = K.LAUNCHES
    a = fe(x)
    used = K.LAUNCHES - n0
    monkeypatch.setenv("CM_NO_FUSE_STEM", "1")
    n0 = K.LAUNCHES
    b = fe(x)
    assert K.LAUNCHES - n0 == used            # one cm_stem_fwd instead of one cm_ln_act_fwd
    assert_close(a, b, torch.float32, what="front-end with / without the stem kernel", rtol_mul=10.0)


# ---------------------------------------------------------------------------------------------------------------------
# cm_gelu_dropout_*_v2: the dropout mask regenerated in backward from the forward's key, bias gradient inside the kernel
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(3, 37, 1024), (64, 501, 1024), (5, 7, 2048), (9, 64), (4, 33, 576), (2, 11, 8)])
def test_gelu_dropout_regenerated_mask_and_fused_bias_gradient(dtype, shape, monkeypatch):
    """The three ways backward can learn the dropout mask - the stored byte mask, the forward's key (mask re-hashed, nothing
    stored: CM_DROPOUT_REGEN=1) and the keep bits (the default: one byte per eight elements) - give the same output and the
    same gradient bit for bit; the fused column sums of dx."""
    from mamba_asr_b200 import kernels as K
    from mamba_asr_b200.layernorm import DropoutSeed
    g = torch.Generator().manual_seed(11)
    x = (2.0 * torch.randn(*shape, generator=g)).to(dtype).cuda()
    dy = torch.randn(*shape, generator=g).to(dtype).cuda()
    seed = DropoutSeed.tensor(x.device)
    cols = shape[-1]
    y0, mask = K.gelu_dropout_forward(x, 0.1, seed, 77, store_mask=True)
    y2, bits = K.gelu_dropout_forward(x, 0.1, seed, 77, store_mask=False)
    monkeypatch.setenv("CM_DROPOUT_REGEN", "1")
    y1, key = K.gelu_dropout_forward(x, 0.1, seed, 77, store_mask=False)
    assert mask.dtype == torch.uint8 and key.dtype == torch.int32 and key.numel() == 1
    assert bits.dtype == torch.uint8 and bits.numel() == x.numel() // 8
    assert torch.equal(y0, y1) and torch.equal(y0, y2)
    packed = (mask.reshape(-1, 8) != 0).to(torch.int32) * (2 ** torch.arange(8, device="cuda", dtype=torch.int32))
    assert torch.equal(packed.sum(1).to(torch.uint8), bits)            # bit i of byte v = element 8 v + i kept
    DropoutSeed.advance(x.device)                       # the key was captured at forward time: a later advance is harmless
    dx0 = K.gelu_dropout_backward(x, dy, mask, 0.1)
    dx1, cs = K.gelu_dropout_backward(x, dy, key, 0.1, colsum_cols=cols)
    dx3, cs3 = K.gelu_dropout_backward(x, dy, bits, 0.1, colsum_cols=cols)
    assert torch.equal(dx0, dx1) and torch.equal(dx0, dx3)
    assert (cs is None and cs3 is None) or torch.equal(cs, cs3)
    assert torch.equal((dx1 == 0) | (dy == 0), (mask == 0) | (dy == 0) | (dx1 == 0))
    fused = bool(K.cabi.lib().cm_act_colsum_supported(x.numel(), cols))
    assert fused == (cols in (1024, 2048, 64, 8))
    if fused:
        ref = dx1.reshape(-1, cols).double().sum(0)
        # the kernel sums its fp32 values before the store rounds them: for 16-bit dx the reference (sum of the rounded dx)
        # differs by the rounding of every term
        assert_close(cs, ref.float(), dtype, floor="max", what="fused column sums of dx")
        dx2, cs2 = K.gelu_dropout_backward(x, dy, key, 0.1, colsum_cols=cols)
        assert torch.equal(cs, cs2)                     # fixed summation order
    else:
        assert cs is None


@pytest.mark.parametrize("autocast", [False, True])
def test_feed_forward_bias_gradient_from_the_activation_kernel(autocast, monkeypatch):
    """PositionalwiseFeedForward: the first Linear's bias gradient formed inside cm_gelu_dropout_bwd against the column-sum
    pass in the Linear's own backward (CM_NO_FUSE_BIAS_GRAD=1)."""
    from mamba_asr_b200.conmamba import PositionalwiseFeedForward
    from mamba_asr_b200 import kernels as K
    torch.manual_seed(4)
    ffn = PositionalwiseFeedForward(d_ffn=1024, input_size=256, dropout=0.0, activation=torch.nn.GELU).cuda()
    x = torch.randn(4, 101, 256, device="cuda", requires_grad=True)
    cy = torch.randn(4, 101, 256, device="cuda")
    res, launches = [], []
    for fused in (True, False):
        if not fused:
            monkeypatch.setenv("CM_NO_FUSE_BIAS_GRAD", "1")
        ffn.zero_grad()
        x.grad = None
        n0 = K.LAUNCHES
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            y = ffn(x)
        (y.float() * cy).sum().backward()
        launches.append(K.LAUNCHES - n0)
        res.append([y.detach().float(), x.grad.clone()] + [p.grad.clone() for p in ffn.parameters()])
    assert launches[0] == launches[1] - 1               # one cm_colsum + its reducer less, one reducer more
    dt = torch.bfloat16 if autocast else torch.float32
    for (a, b) in zip(*res):
        assert_close(a, b, dt, floor="max", what="ffn with / without the fused bias gradient")
    assert torch.equal(res[0][0], res[1][0])


@pytest.mark.parametrize("autocast", [False, True])
def test_encoder_layer_bias_gradients_from_the_consumer_kernels(autocast, monkeypatch):
    """ConmambaEncoderLayer in training mode (dropout on): the bias gradients of the four FFN Linears and of the conv module's
    last Linear formed inside cm_gelu_dropout_bwd / cm_add_ln_bwd against the column-sum passes (CM_NO_FUSE_BIAS_GRAD=1).
    Same seed and call ids in both runs, so the dropout masks agree."""
    from mamba_asr_b200.conmamba import ConmambaEncoderLayer
    from mamba_asr_b200.layernorm import DropoutSeed
    from mamba_asr_b200 import kernels as K
    torch.manual_seed(0)
    layer = ConmambaEncoderLayer(d_model=64, d_ffn=256, activation=torch.nn.GELU, dropout=0.1,
                                 mamba_config=dict(d_state=16, expand=2, d_conv=4, bidirectional=True)).cuda().train()
    x = torch.randn(3, 45, 64, device="cuda")
    cot = torch.randn(3, 45, 64, device="cuda")
    res, launches = [], []
    for fused in (True, False):
        if not fused:
            monkeypatch.setenv("CM_NO_FUSE_BIAS_GRAD", "1")
        DropoutSeed._calls = 1000                       # the same call ids -> the same masks in both runs
        layer.zero_grad(set_to_none=True)
        n0 = K.LAUNCHES
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            o = layer(x)
        (o.float() * cot).sum().backward()
        launches.append(K.LAUNCHES - n0)
        res.append((o.float().detach(), {n: p.grad.clone() for n, p in layer.named_parameters()}))
    assert torch.equal(res[0][0], res[1][0])
    assert launches[0] < launches[1]                    # five cm_colsum passes (and their reducers) fewer
    dt = torch.bfloat16 if autocast else torch.float32
    for n in res[0][1]:
        assert_close(res[0][1][n], res[1][1][n], dt, floor="max", what="d " + n)
