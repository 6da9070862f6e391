"""``ctc_loss`` - drop-in for ``torch.nn.functional.ctc_loss`` on the B200 path (the loss of train_CTC.py:297-302 and the
CTC branch of train_S2S.py:518-530, which reach it through ``speechbrain.nnet.losses.ctc_loss``).

One sm_100a kernel (``cm_ctc_loss``) computes the per-utterance negative log-likelihood AND its gradient with respect to the
log-probabilities (alpha and beta recursions run concurrently in one CTA per utterance); the autograd backward only scales
the stored gradient.  Arguments follow torch: ``log_probs`` (T, B, C) log-softmax outputs, ``targets`` (B, S) padded int64,
``input_lengths`` / ``target_lengths`` (B,), ``reduction`` in {"none", "sum", "mean"}, ``zero_infinity``.  The kernel gives
every extended-label state a thread: up to 255 labels per utterance; longer targets go to torch's CUDA kernels.  No CPU path."""
import torch

from . import kernels as K


class _CtcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_probs_btc, targets, input_lengths, target_lengths, blank, zero_infinity):
        need = ctx.needs_input_grad[0]
        nll, grad = K.ctc_nll_and_grad(log_probs_btc, targets, input_lengths, target_lengths, blank, need_grad=need)
        if zero_infinity:                                # the kernel already wrote a zero gradient for infeasible utterances
            nll = torch.where(torch.isinf(nll), torch.zeros_like(nll), nll)
        if need:
            ctx.save_for_backward(grad)
        return nll

    @staticmethod
    def backward(ctx, dnll):
        (grad,) = ctx.saved_tensors
        return grad * dnll.view(-1, 1, 1), None, None, None, None, None


def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank=0, reduction="mean", zero_infinity=False):
    """torch.nn.functional.ctc_loss semantics for (T, B, C) fp32 CUDA log-probabilities and padded (B, S) int64 targets."""
    if not log_probs.is_cuda:
        raise RuntimeError("mamba_asr_b200.ctc_loss runs on CUDA only (no CPU fallback)")
    if log_probs.dim() != 3:
        raise NotImplementedError("unbatched (T, C) input")
    if targets.dim() != 2:
        raise NotImplementedError("concatenated 1-D targets: pass the padded (B, S) form the recipes use")
    if not K.ctc_supported(targets.shape[1]):
        # more than 255 labels per utterance (e.g. 300 s of speech): torch's own CUDA kernels - same semantics, still no CPU
        return torch.nn.functional.ctc_loss(log_probs.float(), targets, input_lengths, target_lengths, blank=blank,
                                            reduction=reduction, zero_infinity=zero_infinity)
    dev = log_probs.device
    il = torch.as_tensor(input_lengths, dtype=torch.int64).to(dev).contiguous()
    tl = torch.as_tensor(target_lengths, dtype=torch.int64).to(dev).contiguous()
    lp = log_probs.float().transpose(0, 1)                       # (B, T, C) view: the kernel takes the strides
    nll = _CtcFn.apply(lp, targets.to(dev), il, tl, int(blank), bool(zero_infinity))
    if reduction == "none":
        return nll
    if reduction == "sum":
        return nll.sum()
    if reduction == "mean":
        return (nll / tl.clamp_min(1).to(nll.dtype)).mean()
    raise ValueError("reduction must be 'none', 'sum' or 'mean'")
