"""Optimizer side of a training step ("NoamScheduler stand-in", the optimizer step of
train_CTC.py:716-717 / speechbrain Brain.fit_batch): gradient clipping to ``max_grad_norm``, AdamW, Noam learning rate.

Reference objects: ``model_opt_class: torch.optim.AdamW(lr, betas=(0.9, 0.98), eps=1e-9, weight_decay)`` and
``NoamScheduler(lr_initial, n_warmup_steps)`` (hparams/CTC/conmamba_large.yaml:244-252), ``max_grad_norm: 5.0`` (:91).
"""
import torch


def noam_lr(lr_initial, n_warmup_steps, step):
    """speechbrain.nnet.schedulers.NoamScheduler without ``model_size``:
    lr = lr_initial * sqrt(n_warmup) * min(step^-0.5, step * n_warmup^-1.5), step counted from 1."""
    step = max(1, int(step))
    return lr_initial * (n_warmup_steps ** 0.5) * min(step ** -0.5, step * n_warmup_steps ** -1.5)


class TrainStep:
    """clip -> AdamW -> Noam on the parameters of ``model``; every op is stream-ordered on the device (no host sync)."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=5e-4, max_grad_norm=5.0,
                 n_warmup_steps=7500):
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.lr_initial, self.n_warmup_steps, self.max_grad_norm = lr, n_warmup_steps, max_grad_norm
        self.opt = torch.optim.AdamW(self.params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                     fused=self.params[0].is_cuda, foreach=None if self.params[0].is_cuda else False)
        self.steps = 0

    def step(self):
        self.steps += 1
        lr = noam_lr(self.lr_initial, self.n_warmup_steps, self.steps)
        for g in self.opt.param_groups:
            g["lr"] = lr
        if self.max_grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.params, self.max_grad_norm, foreach=self.params[0].is_cuda or None)
        self.opt.step()
        return lr
