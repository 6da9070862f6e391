"""CUDA-graph capture of a whole training step (forward + loss + backward).

The small ConMamba configurations are launch-bound on the host (≈1.5 k kernel launches per step for a 12-layer
encoder); one graph replay removes that cost.  The C-ABI kernels are plain launches on the current stream with no
allocation or host sync, so they are captured like any torch op.
"""
import torch


class GraphedStep:
    """Capture ``fn(*static_inputs)`` once; ``__call__`` refreshes the static inputs and replays.

    ``fn`` must run forward, loss and ``backward()`` and return the loss tensor.  Gradients land in the ``.grad``
    tensors allocated during capture (graph-private pool) and are overwritten by every replay.
    """

    def __init__(self, fn, static_inputs, params, warmup=3):
        self.static_inputs = list(static_inputs)
        params = list(params)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                for p in params:
                    p.grad = None
                fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        for p in params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn(*self.static_inputs)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            if dst is not src:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.out


def graph_module(module, sample_args, warmup=3):
    """Forward and backward CUDA graphs of ``module`` (torch.cuda.make_graphed_callables): the returned callable
    replays the forward graph, and autograd replays the backward graph when the loss computed from its output is
    differentiated.  Use this when the loss itself cannot be captured (e.g. CTC with host-side length tensors)."""
    return torch.cuda.make_graphed_callables(module, tuple(sample_args), num_warmup_iters=warmup)
