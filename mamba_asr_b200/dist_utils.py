"""Data-parallel plumbing for the training configs: the encoder shards by utterance, so the only
collective is the gradient all-reduce of a training step.  One flat all-reduce per step (NCCL over NVLink on the GPU
box, gloo in the CPU tests) instead of DDP's bucket hooks, because the forward / backward of the model are replayed as
CUDA graphs (torch.cuda.make_graphed_callables) and the 10-100 M parameter gradients of these models are a single
sub-millisecond NVLink transfer."""
import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


def allreduce_gradients(params, world_size=None, group=None):
    """Average the ``.grad`` of ``params`` over the process group with ONE all-reduce (grouped by dtype)."""
    if world_size is None:
        world_size = dist.get_world_size(group)
    by_dtype = {}
    for p in params:
        if p.grad is not None:
            by_dtype.setdefault(p.grad.dtype, []).append(p.grad)
    for grads in by_dtype.values():
        flat = _flatten_dense_tensors(grads)
        dist.all_reduce(flat, group=group)
        flat.div_(world_size)
        torch._foreach_copy_(grads, list(_unflatten_dense_tensors(flat, grads)))
