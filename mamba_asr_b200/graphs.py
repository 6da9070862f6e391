"""CUDA-graph replay of a model's forward and backward.

The small ConMamba configurations are launch-bound on the host (about 1.5 k kernel launches per step for a 12-layer
encoder); graph replay removes that cost.  The C-ABI kernels are plain launches on the current stream with no
allocation or host sync, so they are captured like any torch op.
"""
import torch


def graph_module(module, sample_args, warmup=3):
    """Forward and backward CUDA graphs of ``module`` (torch.cuda.make_graphed_callables): the returned callable
    replays the forward graph, and autograd replays the backward graph when the loss computed from its output is
    differentiated.  Use this when the loss itself cannot be captured (e.g. CTC with host-side length tensors)."""
    return torch.cuda.make_graphed_callables(module, tuple(sample_args), num_warmup_iters=warmup)


def graph_forward(fn, static_inputs, warmup=3):
    """One CUDA graph of ``fn(*static_inputs)`` for gradient-free passes (evaluation, inference): returns ``run(*inputs)``
    that copies the inputs into the captured buffers, replays, and returns the captured output tensor(s) (overwritten by
    the next replay).  ``fn`` must be free of host synchronisation."""
    static_inputs = list(static_inputs)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):
            fn(*static_inputs)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = fn(*static_inputs)

    def run(*inputs):
        for dst, src in zip(static_inputs, inputs):
            if dst is not src:
                dst.copy_(src, non_blocking=True)
        graph.replay()
        return out

    run.graph = graph
    return run
