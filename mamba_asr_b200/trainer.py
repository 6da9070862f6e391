"""Optimizer side of a training step ("NoamScheduler stand-in", the optimizer step of
train_CTC.py:716-717 / speechbrain Brain.fit_batch): gradient clipping to ``max_grad_norm``, AdamW, Noam learning rate.

Reference objects: ``model_opt_class: torch.optim.AdamW(lr, betas=(0.9, 0.98), eps=1e-9, weight_decay)`` and
``NoamScheduler(lr_initial, n_warmup_steps)`` (hparams/CTC/conmamba_large.yaml:244-252), ``max_grad_norm: 5.0`` (:91).

On CUDA the step runs on FLAT buffers: every parameter of the model is re-pointed to a view of one fp32 buffer (``state_dict``
keys and shapes are untouched), gradients are gathered into one flat fp32 buffer by a single multi-tensor copy, and the whole
update - global-norm clipping included - is two kernel launches (cm_sumsq_partial + cm_adamw_step).  The flat gradient buffer
is also what data-parallel training all-reduces, in place: no flatten / unflatten / divide passes (the 1 / world_size average
is folded into the update).  On CPU (the reference arm of bench.py) the same arithmetic runs through torch's own AdamW.
"""
import torch

from . import kernels as K
from .linear import invalidate_param_cache


def noam_lr(lr_initial, n_warmup_steps, step):
    """speechbrain.nnet.schedulers.NoamScheduler without ``model_size``:
    lr = lr_initial * sqrt(n_warmup) * min(step^-0.5, step * n_warmup^-1.5), step counted from 1."""
    step = max(1, int(step))
    return lr_initial * (n_warmup_steps ** 0.5) * min(step ** -0.5, step * n_warmup_steps ** -1.5)


class TrainStep:
    """clip -> AdamW -> Noam on the parameters of ``model``; every op is stream-ordered on the device (no host sync).

    Build it BEFORE capturing CUDA graphs of the model: on CUDA it moves the parameters into one flat buffer."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=5e-4, max_grad_norm=5.0,
                 n_warmup_steps=7500, world_size=1, group=None):
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.lr_initial, self.n_warmup_steps, self.max_grad_norm = lr, n_warmup_steps, max_grad_norm
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.world_size, self.group = world_size, group
        self.steps = 0
        self.flat = bool(self.params) and all(p.is_cuda and p.dtype == torch.float32 for p in self.params)
        if not self.flat:
            self.opt = torch.optim.AdamW(self.params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
            return
        dev = self.params[0].device
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + 7) // 8 * 8                    # 32-byte aligned views; the padding stays zero
        self.n = n
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.scratch = torch.empty(max(1, K.cabi.lib().cm_optim_num_part(n)), dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.g_views = []
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                view = self.flat_p[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view                                # same values, new storage: state_dict is unchanged
                self.g_views.append(self.flat_g[o:o + p.numel()].view(p.shape))

    def gather_grads(self):
        """One multi-tensor copy of the parameters' ``.grad`` into the flat gradient buffer (parameters without a gradient
        contribute zeros)."""
        src, dst, missing = [], [], []
        for p, gv in zip(self.params, self.g_views):
            if p.grad is None:
                missing.append(gv)
            else:
                src.append(p.grad)
                dst.append(gv)
        if missing:
            torch._foreach_zero_(missing)
        if dst:
            torch._foreach_copy_(dst, src)

    def allreduce(self, dist):
        """Sum the flat gradient buffer over the ranks, in place (the average is folded into the update)."""
        dist.all_reduce(self.flat_g, group=self.group)

    def step(self, dist=None):
        self.steps += 1
        lr = noam_lr(self.lr_initial, self.n_warmup_steps, self.steps)
        if not self.flat:
            for g in self.opt.param_groups:
                g["lr"] = lr
            if dist is not None and self.world_size > 1:
                from .dist_utils import allreduce_gradients
                allreduce_gradients(self.params, self.world_size, self.group)
            if self.max_grad_norm is not None:
                torch.nn.utils.clip_grad_norm_(self.params, self.max_grad_norm)
            self.opt.step()
            return lr
        self.gather_grads()
        if dist is not None and self.world_size > 1:
            self.allreduce(dist)
        K.adamw_step(self.flat_p, self.flat_g, self.m, self.v, lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                     self.steps, max_grad_norm=self.max_grad_norm or 0.0, grad_scale=1.0 / self.world_size,
                     scratch=self.scratch, norm_out=self.grad_norm)
        invalidate_param_cache()                             # the kernel wrote the parameters behind autograd's back
        return lr
