#!/usr/bin/env python
"""Kernel-level breakdown of one bench.py training step (torch.profiler, CUDA time by kernel name) - run on the B200 box.

    python tools/step_profile.py [--workload NAME] [--top 40]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mamba_asr_b200.encoder import CONFIGS, build_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default=bench.DEFAULT_WORKLOAD)
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--graphed", action="store_true",
                    help="profile the step bench.py times by default: forward and backward replayed as CUDA graphs, the "
                         "backward's partial sums reduced in one batch (kernels.deferred_reductions)")
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    cfg = CONFIGS[wl["model"]]
    dev = torch.device("cuda:0")
    model = build_model(wl["model"]).to(dev).train()
    model.enable_param_cache()
    wav, targets = bench.make_batch(cfg, wl["batch"], wl["seconds"], 1234, dev, cfg["output_neurons"])
    wav, targets = wav.to(dev), targets.to(dev)
    if args.graphed:
        from mamba_asr_b200 import kernels as K
        from mamba_asr_b200.graphs import graph_module
        is_s2s = hasattr(model, "decoder")
        sample = (wav,) if not is_s2s else (wav, torch.cat([torch.ones_like(targets[:, :1]), targets], dim=1))
        with torch.autocast("cuda", dtype=torch.bfloat16, cache_enabled=False), K.deferred_reductions():
            model = graph_module(model, sample, warmup=3)
    for _ in range(3):
        model.zero_grad(set_to_none=True)
        bench.ctc_step(model, wav, targets, True)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model.zero_grad(set_to_none=True)
        bench.ctc_step(model, wav, targets, True)
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0)
        if e.device_type == torch.autograd.DeviceType.CUDA and t > 0:
            rows.append((t, e.count, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print("total CUDA kernel time %.3f ms in %d launches (%d distinct kernels)" % (tot / 1e3, sum(r[1] for r in rows), len(rows)))
    for t, n, k in rows[:args.top]:
        print("%8.3f ms %5.1f%% %5d x  %s" % (t / 1e3, 100 * t / tot, n, k[:150]))


if __name__ == "__main__":
    main()
