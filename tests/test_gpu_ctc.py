"""GPU parity of cm_ctc_loss (mamba_asr_b200.ctc.ctc_loss) against torch.nn.functional.ctc_loss - the loss the reference
recipes reach through speechbrain.nnet.losses.ctc_loss (train_CTC.py:297-302, train_S2S.py:518-530).

The loss values are compared directly.  Gradients are compared at the LOGITS (through log_softmax): torch's native backward
returns exp(lp) - posterior for d/d log_probs (the derivative already folded with the softmax), the kernel returns the plain
derivative -posterior; both give the same logits gradient, which is what training sees.  Tolerances: fp32, 2e-4 relative on
the loss (hundreds of nats over 500 frames), 2e-4 absolute on logits gradients (values in [-1, 1])."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _case(Bt, T, Cn, S, seed, ragged=True, repeats=False):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(T, Bt, Cn, generator=g) * 2.0
    targets = torch.randint(1, Cn, (Bt, S), generator=g)
    if repeats and S >= 4:
        targets[:, 1] = targets[:, 0]
        targets[:, 3] = targets[:, 2]
    if ragged:
        il = torch.randint(max(1, T // 2), T + 1, (Bt,), generator=g)
        tl = torch.randint(0 if S == 0 else 1, S + 1, (Bt,), generator=g)
        il[0], tl[0] = T, S
    else:
        il = torch.full((Bt,), T, dtype=torch.long)
        tl = torch.full((Bt,), S, dtype=torch.long)
    return logits, targets, il, tl


def _both(logits, targets, il, tl, reduction, zero_infinity, blank=0):
    from mamba_asr_b200.ctc import ctc_loss
    dev = torch.device("cuda:0")
    out = []
    for fn in (ctc_loss, F.ctc_loss):
        x = logits.to(dev).clone().requires_grad_(True)
        lp = F.log_softmax(x, dim=-1)
        loss = fn(lp, targets.to(dev), il.to(dev), tl.to(dev), blank=blank, reduction=reduction, zero_infinity=zero_infinity)
        (loss.sum() * 1.7).backward()
        out.append((loss.detach().cpu(), x.grad.detach().cpu()))
    return out


@pytest.mark.parametrize("shape", [(4, 50, 31, 10), (64, 501, 31, 60), (3, 17, 5, 4), (2, 33, 1000, 12), (5, 16, 31, 7),
                                   (1, 1, 4, 1), (8, 200, 5000, 40), (2, 300, 31, 255), (6, 40, 7, 1)])
@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
def test_ctc_loss_matches_torch(shape, reduction):
    Bt, T, Cn, S = shape
    logits, targets, il, tl = _case(Bt, T, Cn, S, seed=Bt * 100 + T + S, ragged=True, repeats=True)
    (l1, g1), (l0, g0) = _both(logits, targets, il, tl, reduction, zero_infinity=True)
    torch.testing.assert_close(l1, l0, rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(g1, g0, rtol=1e-3, atol=2e-4)


def test_ctc_loss_full_lengths_and_other_blank():
    logits, targets, il, tl = _case(8, 120, 40, 20, seed=5, ragged=False)
    targets = targets.clamp(max=38)                     # blank = 39: labels in [1, 38]
    (l1, g1), (l0, g0) = _both(logits, targets, il, tl, "mean", zero_infinity=False, blank=39)
    torch.testing.assert_close(l1, l0, rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(g1, g0, rtol=1e-3, atol=2e-4)


def test_ctc_loss_infeasible_alignment_zero_infinity():
    # utterance 1 has fewer frames than labels (and repeated labels need a blank between them): infinite loss
    logits, targets, il, tl = _case(3, 12, 9, 10, seed=11, ragged=False)
    il[1] = 6
    targets[2, :] = 3
    il[2] = 12                                           # ten repeats of one label need 19 frames
    (l1, g1), (l0, g0) = _both(logits, targets, il, tl, "none", zero_infinity=True)
    assert l1[1] == 0 and l1[2] == 0
    torch.testing.assert_close(l1, l0, rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(g1, g0, rtol=1e-3, atol=2e-4)
    assert torch.all(g1[:, 1] == 0) and torch.all(g1[:, 2] == 0)
    (l1, _), (l0, _) = _both(logits, targets, il, tl, "none", zero_infinity=False)
    assert torch.isinf(l1[1]) and torch.isinf(l1[2]) and torch.isinf(l0[1])


def test_ctc_loss_empty_targets():
    logits, targets, il, tl = _case(4, 30, 11, 6, seed=3, ragged=True)
    tl[1] = 0
    tl[3] = 0
    (l1, g1), (l0, g0) = _both(logits, targets, il, tl, "sum", zero_infinity=True)
    torch.testing.assert_close(l1, l0, rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(g1, g0, rtol=1e-3, atol=2e-4)


def test_ctc_loss_is_deterministic_and_strided():
    from mamba_asr_b200.ctc import ctc_loss
    dev = torch.device("cuda:0")
    logits, targets, il, tl = _case(16, 250, 31, 50, seed=9)
    big = torch.zeros(250, 16, 64, device=dev)
    big[..., :31] = F.log_softmax(logits.to(dev), dim=-1)
    outs = []
    for view in (big[..., :31], big[..., :31].contiguous()):
        for _ in range(2):
            x = view.detach().clone().requires_grad_(True) if view.is_contiguous() else view.detach().requires_grad_(True)
            loss = ctc_loss(x, targets.to(dev), il.to(dev), tl.to(dev), reduction="none")
            g, = torch.autograd.grad(loss.sum(), x)
            outs.append((loss.cpu(), g.cpu()))
    for l, g in outs[1:]:
        assert torch.equal(l, outs[0][0]) and torch.equal(g, outs[0][1])


def test_ctc_loss_rejects_what_it_does_not_implement():
    from mamba_asr_b200.ctc import ctc_loss
    dev = torch.device("cuda:0")
    lp = torch.zeros(10, 2, 5, device=dev)
    with pytest.raises(RuntimeError):
        ctc_loss(lp.cpu(), torch.ones(2, 3, dtype=torch.long), [10, 10], [3, 3])
    # more than 255 labels per utterance: torch's CUDA kernels behind the same call
    lp = torch.log_softmax(torch.randn(700, 2, 9, device=dev), -1)
    tg = torch.randint(1, 9, (2, 300), device=dev)
    a = ctc_loss(lp, tg, torch.tensor([700, 650], device=dev), torch.tensor([300, 280], device=dev), reduction="sum")
    b = F.ctc_loss(lp, tg, torch.tensor([700, 650], device=dev), torch.tensor([300, 280], device=dev), reduction="sum")
    assert torch.equal(a, b)
