"""cm_reduce_batch (the batched fixed-order reducer) and kernels.deferred_reductions (the partial sums of a whole backward
pass reduced in one batch at its end): the kernel against torch sums over ragged job mixes, and the queued path against the
immediate one on whole models - bit-identical parameter gradients, with queued outputs NaN-poisoned until the flush so that a
premature read cannot pass."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _jobs(shapes, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    jobs, refs = [], []
    for rows, cols, pad, off in shapes:
        stride = cols + pad + off
        buf = torch.randn(rows, stride, device="cuda", generator=g)
        out = torch.full((cols,), float("nan"), device="cuda")
        jobs.append((buf, out, rows, cols, stride, off))
        refs.append(buf[:, off:off + cols].double().sum(0))
    return jobs, refs


@pytest.mark.parametrize("shapes", [
    [(1, 1, 0, 0)],
    [(592, 256, 0, 0), (128, 31 * 256, 0, 0), (33, 7, 0, 0), (64, 33, 5, 3)],                      # tall jobs, ragged
    [(4, 256 * 1024, 0, 0), (16, 48 * 512, 0, 0), (8, 4100, 0, 0), (3, 10, 0, 0), (32, 6, 2, 1)],   # wide jobs, vector + scalar
    [(8, 512, 0, 0), (700, 512, 0, 0), (2, 4096 * 3 + 4, 4, 0), (40, 1, 0, 0)] * 9,                 # > one batch, mixed
])
def test_reduce_batch_matches_torch_sums(shapes):
    from mamba_asr_b200 import kernels as K
    jobs, refs = _jobs(shapes, 3)
    K.reduce_many(jobs)
    again, _ = _jobs(shapes, 3)
    K.reduce_many(again)
    for (buf, out, rows, cols, stride, off), ref, (_, out2, *_r) in zip(jobs, refs, again):
        assert torch.isfinite(out).all()
        assert (out.double() - ref).abs().max() <= 1e-5 * max(1.0, rows ** 0.5)
        assert torch.equal(out, out2)                                     # fixed order: run-to-run identical


def test_reduce_batch_rejects_bad_jobs():
    import ctypes as C
    from mamba_asr_b200 import _cabi as cabi
    lib = cabi.lib()
    buf = torch.zeros(4, 8, device="cuda")
    out = torch.zeros(8, device="cuda")
    arr = (cabi.ReduceJob2 * 1)()
    arr[0].part, arr[0].out, arr[0].rows, arr[0].cols, arr[0].stride = buf.data_ptr(), out.data_ptr(), 4, 8, 4
    assert lib.cm_reduce_batch(arr, 1, None) == cabi.CM_ERR_BAD_ARG       # stride < cols
    assert lib.cm_reduce_batch(arr, 0, None) == cabi.CM_ERR_BAD_ARG
    assert lib.cm_reduce_batch(arr, cabi.CM_REDUCE_BATCH_MAX + 1, None) == cabi.CM_ERR_BAD_ARG


def _ctc_model():
    from mamba_asr_b200.encoder import build_model
    torch.manual_seed(5)
    m = build_model("conmamba_small_ctc", dropout=0.0, d_model=32, d_ffn=64, num_layers=2).cuda().train()
    g = torch.Generator().manual_seed(1)
    wav = (0.1 * torch.randn(3, 12000, generator=g)).cuda()
    tgt = torch.randint(1, 31, (3, 6), generator=g).cuda()

    def loss_of(model):
        logp = model(wav)
        L = logp.shape[1]
        return F.ctc_loss(logp.transpose(0, 1).float(), tgt, torch.full((3,), L), torch.full((3,), 6), blank=0, reduction="mean")
    return m, loss_of


def _s2s_model():
    from mamba_asr_b200.encoder import build_model, kldiv_loss
    torch.manual_seed(6)
    m = build_model("conmambamamba_large_s2s", dropout=0.0, d_model=32, d_ffn=64, num_layers=1, num_decoder_layers=2,
                    output_neurons=40).cuda().train()
    g = torch.Generator().manual_seed(2)
    wav = (0.1 * torch.randn(2, 9000, generator=g)).cuda()
    bos = torch.randint(3, 40, (2, 7), generator=g)
    bos[:, 0] = 1
    eos = torch.cat([bos[:, 1:], torch.full((2, 1), 2)], dim=1).cuda()
    bos = bos.cuda()

    def loss_of(model):
        p_ctc, p_seq = model(wav, bos)
        return kldiv_loss(p_seq, eos, label_smoothing=0.1) + 0.3 * p_ctc[..., 0].float().mean()
    return m, loss_of


@pytest.mark.parametrize("make", [_ctc_model, _s2s_model])
@pytest.mark.parametrize("autocast", [False, True])
def test_deferred_reductions_give_identical_parameter_gradients(make, autocast, monkeypatch):
    """torch.autograd.grad of a whole model with the reductions queued (and their outputs poisoned until the flush) against
    the same pass with every reduction run where it is produced: every parameter gradient bit-identical."""
    from mamba_asr_b200 import kernels as K
    model, loss_of = make()
    if autocast:
        model.enable_param_cache()
    params = [p for p in model.parameters() if p.requires_grad]

    def grads(defer):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast), K.deferred_reductions(defer):
            loss = loss_of(model)
            l0 = K.LAUNCHES
            gs = torch.autograd.grad(loss, params, allow_unused=True)
            n = K.LAUNCHES - l0
        torch.cuda.synchronize()
        return gs, n
    ref, n_ref = grads(False)
    monkeypatch.setattr(K, "_DEFER_POISON", True)
    got, n_got = grads(True)
    assert not K._PENDING
    for p, a, b in zip(params, ref, got):
        assert (a is None) == (b is None)
        if a is not None:
            assert torch.isfinite(b).all()
            assert torch.equal(a, b)
    assert n_got < n_ref                                    # fewer launches: the per-operator reducers are gone


def test_deferred_reductions_under_graph_capture_match_eager():
    """make_graphed_callables with the reductions queued (what bench.py replays): gradients of two replays equal the eager,
    immediately-reduced ones bit for bit."""
    from mamba_asr_b200 import kernels as K
    from mamba_asr_b200.graphs import graph_module
    model, _ = _ctc_model()
    g = torch.Generator().manual_seed(1)
    wav = (0.1 * torch.randn(3, 12000, generator=g)).cuda()

    def run(net):
        model.zero_grad(set_to_none=True)
        net(wav).float().square().mean().backward()
        torch.cuda.synchronize()
        return [None if p.grad is None else p.grad.clone() for p in model.parameters()]
    ref = run(model)
    with K.deferred_reductions():
        gnet = graph_module(model, (wav,), warmup=3)
    for _ in range(2):
        got = run(gnet)
        for a, b in zip(ref, got):
            assert (a is None) == (b is None)
            if a is not None:
                assert torch.equal(a, b)


def test_reduce_many_outside_a_backward_pass_runs_immediately():
    from mamba_asr_b200 import kernels as K
    part = torch.randn(40, 64, device="cuda")
    with K.deferred_reductions():
        out = K.sum_leading(part, defer=True)              # no backward pass running: nothing could flush a queue
        assert not K._PENDING
    assert torch.allclose(out, part.sum(0), atol=1e-5)
