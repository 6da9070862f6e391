"""GPU parity of the layer / model shells around the hot path:
the same module tree evaluated (a) on the sm_100a kernels and (b) on the CPU reference path of oracle/cpu_encoder.py
(selective_scan_ref + torch conv composed as bimamba.py:223-253, the reference's transpose -> nn.Conv1d convolution
module, nn.LayerNorm, the Fbank restatement)."""
import copy

import pytest
import torch
import torch.nn.functional as F

from cm_testutil import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _exact_fp32_library_ops():
    """The shells' torch ops (cuDNN front-end convs, cuBLAS GEMMs) default to TF32 for fp32 inputs; the parity check
    needs them at fp32 so that what is compared is the hand-written kernels."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _pair(name, **over):
    from mamba_asr_b200.encoder import CONFIGS, build_model
    from oracle.cpu_encoder import to_cpu_reference
    cfg = CONFIGS[name]
    gpu = build_model(name, dropout=0.0, **over)
    with torch.no_grad():                      # leave the init point where D = 1 and A = -(1..N) exactly
        for n, p in gpu.named_parameters():
            if n.endswith(".D") or n.endswith(".D_b") or "A_log" in n or "A_b_log" in n:
                p.add_(0.2 * torch.randn(p.shape, generator=torch.Generator().manual_seed(len(n))))
    cpu = to_cpu_reference(copy.deepcopy(gpu), cfg["n_fft"], cfg["n_mels"], cfg["win_length"])
    return gpu.cuda(), cpu


def _grads_close(gpu, cpu, skip=()):
    ref = dict(cpu.named_parameters())
    n = 0
    for name, p in gpu.named_parameters():
        rname = name
        for cand in (name, name.replace(".mamba.", ".mamba.inner."), name.replace(".self_mamba.", ".self_mamba.inner."),
                     name.replace(".cross_mamba.", ".cross_mamba.inner.")):
            if cand in ref:
                rname = cand
                break
        g = ref[rname].grad
        assert p.grad is not None and g is not None, name
        assert_close(p.grad, g, floor="max", what="d " + name, rtol_mul=5.0)
        n += 1
    assert n == len(ref)


def test_ctc_encoder_stack_forward_backward_matches_cpu_reference():
    gpu, cpu = _pair("conmamba_small_ctc", d_model=32, d_ffn=64, num_layers=2)
    g = torch.Generator().manual_seed(1)
    wav = 0.1 * torch.randn(3, 12000, generator=g)
    tgt = torch.randint(1, 31, (3, 6), generator=g)

    def loss_of(model, w):
        logp = model(w)
        L = logp.shape[1]
        return logp, F.ctc_loss(logp.transpose(0, 1), tgt, torch.full((3,), L), torch.full((3,), 6), blank=0,
                                reduction="mean")
    lp_c, loss_c = loss_of(cpu, wav)
    loss_c.backward()
    lp_g, loss_g = loss_of(gpu, wav.cuda())
    loss_g.backward()
    assert lp_g.shape == lp_c.shape
    assert_close(lp_g, lp_c, floor="max", what="log-probs", rtol_mul=5.0)
    assert abs(float(loss_g.detach()) - float(loss_c.detach())) <= 1e-4 * abs(float(loss_c.detach()))
    _grads_close(gpu, cpu)


def test_s2s_encoder_decoder_forward_backward_matches_cpu_reference():
    from mamba_asr_b200.encoder import kldiv_loss
    gpu, cpu = _pair("conmambamamba_large_s2s", d_model=32, d_ffn=64, num_layers=1, num_decoder_layers=2,
                     output_neurons=40)
    g = torch.Generator().manual_seed(2)
    wav = 0.1 * torch.randn(2, 9000, generator=g)
    bos = torch.randint(3, 40, (2, 7), generator=g)
    bos[:, 0] = 1
    eos = torch.cat([bos[:, 1:], torch.full((2, 1), 2)], dim=1)
    eos[1, -2:] = 0                                        # padding positions are masked out of the loss

    def loss_of(model, w, b, e):
        p_ctc, p_seq = model(w, b)
        return p_ctc, p_seq, kldiv_loss(p_seq, e, label_smoothing=0.1) + 0.3 * p_ctc[..., 0].mean()
    pc_c, ps_c, loss_c = loss_of(cpu, wav, bos, eos)
    loss_c.backward()
    pc_g, ps_g, loss_g = loss_of(gpu, wav.cuda(), bos.cuda(), eos.cuda())
    loss_g.backward()
    assert_close(pc_g, pc_c, floor="max", what="p_ctc", rtol_mul=5.0)
    assert_close(ps_g, ps_c, floor="max", what="p_seq", rtol_mul=5.0)
    assert abs(float(loss_g.detach()) - float(loss_c.detach())) <= 1e-4 * abs(float(loss_c.detach()))
    _grads_close(gpu, cpu)


def test_config1_small_encoder_forward_bf16_against_cpu_reference():
    """BASELINE.json configs[0]: ConMamba-small CTC encoder forward, batch 8 x 10 s, reference path on CPU (fp32) next
    to the bf16-autocast GPU forward.  Tolerance: bf16 rtol 2e-2 of the log-prob range, accumulated over 12 layers."""
    gpu, cpu = _pair("conmamba_small_ctc")
    gpu.eval(), cpu.eval()
    wav = 0.1 * torch.randn(8, 160000, generator=torch.Generator().manual_seed(7775))
    with torch.no_grad():
        ref = cpu(wav)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = gpu(wav.cuda())
        out32 = gpu(wav.cuda())
    assert out.shape == ref.shape == (8, 251, 31)
    assert_close(out32, ref, floor="max", what="fp32 log-probs", rtol_mul=10.0)
    err = (out.float().cpu() - ref).abs()
    assert float(err.mean()) <= 2e-2 * float(ref.abs().max()), float(err.mean())
    assert float(err.max()) <= 0.15 * float(ref.abs().max()), float(err.max())
    assert float((out.float().cpu().argmax(-1) == ref.argmax(-1)).float().mean()) > 0.9


def test_param_cache_step_equals_per_use_casts():
    """bf16 autocast training step with the flat bf16 parameter cache (one multi-tensor copy per step, fp32 weight
    gradients straight from the GEMMs) against the same step with per-use casts: same loss, gradients within bf16
    rounding of each other; and the cache follows parameter updates."""
    from mamba_asr_b200.encoder import build_model
    from mamba_asr_b200.linear import set_param_cache
    g = torch.Generator().manual_seed(3)
    wav = (0.1 * torch.randn(2, 16000, generator=g)).cuda()
    tgt = torch.randint(1, 31, (2, 5), generator=g)

    def run(model):
        model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logp = model(wav)
        L = logp.shape[1]
        loss = F.ctc_loss(logp.float().transpose(0, 1), tgt, torch.full((2,), L), torch.full((2,), 5), blank=0)
        loss.backward()
        return float(loss.detach()), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    model = build_model("conmamba_small_ctc", d_model=64, d_ffn=128, num_layers=2, dropout=0.0).cuda()
    try:
        set_param_cache(None)
        loss0, g0 = run(model)
        model.enable_param_cache()
        loss1, g1 = run(model)
        assert abs(loss1 - loss0) <= 2e-2 * abs(loss0)
        assert g0.keys() == g1.keys()
        for n in g0:
            assert g1[n].dtype == g0[n].dtype == torch.float32
            assert_close(g1[n], g0[n], torch.bfloat16, floor="max", what="d " + n, rtol_mul=2.0)
        with torch.no_grad():
            for p in model.parameters():
                p.mul_(0.5)
        loss2, _ = run(model)
        set_param_cache(None)
        model._param_cache = None
        loss3, _ = run(model)
        assert abs(loss2 - loss3) <= 2e-2 * abs(loss3) and abs(loss2 - loss1) > 1e-3
    finally:
        set_param_cache(None)


def test_config5_long_form_inference_windowed_scan_equals_whole_sequence(monkeypatch):
    """BASELINE.json configs[4]: ConMamba-large encoder inference on 4 x 300 s (30001 fbank frames, 7501 encoder frames).
    No CPU oracle finishes at this size; the size-independent property is that the chunk-parallel scan launch (time
    windows) and the whole-sequence launch produce the same encoder output."""
    from mamba_asr_b200 import kernels as K
    from mamba_asr_b200.encoder import build_model
    torch.manual_seed(0)
    model = build_model("conmamba_large_ctc", num_layers=3).cuda().eval()
    wav = 0.1 * torch.randn(4, 16000 * 300, device="cuda")
    outs = []
    for no_windows in (False, True):
        if no_windows:
            monkeypatch.setenv("CM_SCAN_NO_WINDOWS", "1")
        n0 = K.LAUNCHES
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(wav)
        torch.cuda.synchronize()
        outs.append((out.float(), K.LAUNCHES - n0))
    assert outs[0][0].shape == (4, 7501, 31) and bool(torch.isfinite(outs[0][0]).all())
    # the two launches order the state arithmetic differently (window summaries + combine vs one pass): bit-identical
    # outputs would mean the windowed path did not run at this shape
    assert not torch.equal(outs[0][0], outs[1][0])
    assert_close(outs[0][0], outs[1][0], torch.bfloat16, floor="max", what="windowed vs whole-sequence log-probs")


def test_graph_replays_draw_fresh_dropout_masks():
    """under CUDA-graph replay the fused dropout kernels must not repeat their masks.  The model forward
    advances the device seed in place (captured, so every replay advances it again)."""
    from mamba_asr_b200.encoder import build_model
    from mamba_asr_b200.graphs import graph_module
    from mamba_asr_b200.layernorm import DropoutSeed, gelu_dropout
    # (i) the kernel-level contract: seed advance + fused GELU/dropout inside one captured graph
    x = torch.randn(64, 256, device="cuda").abs() + 0.5           # gelu(x) != 0 everywhere: zeros are dropped elements
    DropoutSeed.tensor(x.device)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            DropoutSeed.advance(x.device)
            gelu_dropout(x, 0.5, True)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        DropoutSeed.advance(x.device)
        y = gelu_dropout(x, 0.5, True)
    masks = []
    for _ in range(3):
        g.replay()
        masks.append((y == 0).clone())
    assert not torch.equal(masks[0], masks[1]) and not torch.equal(masks[1], masks[2])
    assert 0.4 < masks[0].float().mean().item() < 0.6
    # (ii) the model-level path bench.py uses: torch's own Dropout switched off, only the fused kernels are random
    torch.manual_seed(0)
    m = build_model("conmamba_small_ctc", num_layers=2).cuda().train()
    m.custom_src_module[1].p = 0.0
    wav = 0.1 * torch.randn(2, 8000, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16, cache_enabled=False):
        gm = graph_module(m, (wav,), warmup=3)
        outs = [gm(wav).detach().float().clone() for _ in range(3)]
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
    m.eval()                                                 # no dropout: eager forwards are deterministic
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        a, b = m(wav), m(wav)
    assert torch.equal(a, b)


def test_param_cache_eval_after_update_matches_uncached_model():
    """eval under no_grad + autocast after enable_param_cache() and a parameter update must use the
    CURRENT weights everywhere (x_proj / dt_proj included), i.e. match the same model without a cache."""
    from mamba_asr_b200.encoder import build_model
    from mamba_asr_b200.linear import set_param_cache
    torch.manual_seed(1)
    m = build_model("conmamba_small_ctc", num_layers=2).cuda().eval()
    wav = 0.1 * torch.randn(2, 8000, device="cuda")
    m.enable_param_cache()
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.5)                                      # an "optimizer step" after the cache was built
        with torch.autocast("cuda", dtype=torch.bfloat16):
            cached = m(wav).float()
        set_param_cache(None)
        m._param_cache = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            plain = m(wav).float()
    assert_close(cached, plain, torch.bfloat16, floor="max", what="cached vs uncached eval")
    # direct submodule call without a refresh after another update: the stale copies must not be used
    m.enable_param_cache()
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(0.5)
        feats = m.features(wav)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            e1 = m.encode(feats).float()
        set_param_cache(None)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            e2 = m.encode(feats).float()
    assert_close(e1, e2, torch.bfloat16, floor="max", what="encode() with a stale cache")


def test_flat_adamw_step_matches_torch_adamw_with_clipping():
    """trainer.TrainStep on CUDA (flat buffers, cm_sumsq_partial + cm_adamw_step) against torch.optim.AdamW +
    clip_grad_norm_ + the Noam formula over several steps, with and without active clipping; parameters keep their
    state_dict names / shapes after being re-pointed into the flat buffer."""
    from mamba_asr_b200.trainer import TrainStep, noam_lr
    torch.manual_seed(3)
    def make():
        torch.manual_seed(3)
        return torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.LayerNorm(64), torch.nn.Linear(64, 5)).cuda()
    ours, ref = make(), make()
    keys = list(ours.state_dict().keys())
    hp = dict(lr=2e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=5e-2)
    for max_norm in (0.05, 50.0):                       # clipping active / inactive
        ts = TrainStep(ours, max_grad_norm=max_norm, n_warmup_steps=4, **hp)
        assert list(ours.state_dict().keys()) == keys
        opt = torch.optim.AdamW(ref.parameters(), **hp)
        for step in range(1, 6):
            x = torch.randn(16, 37, device="cuda", generator=torch.Generator(device="cuda").manual_seed(step))
            for net in (ours, ref):
                net.zero_grad(set_to_none=True)
                net(x).square().mean().backward()
            for g_ in opt.param_groups:
                g_["lr"] = noam_lr(hp["lr"], 4, step)
            norm_ref = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm)
            opt.step()
            ts.step()
            assert_close(ts.grad_norm, norm_ref.reshape(1), floor="max", what=f"grad norm step {step}")
            for (n1, p1), (n2, p2) in zip(ours.named_parameters(), ref.named_parameters()):
                assert_close(p1, p2, floor="max", what=f"{n1} after step {step} (max_norm {max_norm})")
        # a parameter without a gradient contributes zeros (and still decays)
        ours[2].bias.grad = None
        ref[2].bias.grad = None
