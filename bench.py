#!/usr/bin/env python
"""bench.py - ConMamba encoder audio-seconds / second on N B200s, with the scan-kernel roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--sweep-L]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Default workload (BASELINE.json configs[2], the configuration the metric and the target are quoted on): ConMamba-large
CTC training step, bf16 autocast, batch 64 x 20 s of synthetic 16 kHz audio PER GPU; weak scaling (DDP gradient all-reduce
over NCCL at N > 1).  One step = Fbank -> normalise -> conv front-end -> 18 ConMamba layers -> CTC loss -> backward ->
gradient clipping + AdamW + Noam learning rate (hparams/CTC/conmamba_large.yaml:91-99, 244-252; train_CTC.py:716-717).
Other workloads (--workload): configs[0] forward only (8 x 10 s, small), configs[1] (small, 32 x 15 s, fwd+bwd),
configs[3] (S2S large), configs[4] (4 x 300 s inference; --sweep-L adds the kernel-level scan sweep over L = 1k..30k).

Printed JSON line (rank 0):
  value     audio-s/s of the whole job, inputs already resident in HBM, K steps timed with CUDA events between
            barrier + synchronize, max over ranks
  e2e       the same step driven from pinned HOST audio: H2D copy of every step's waveforms (issued on a copy stream,
            double-buffered, overlapping the previous step) and D2H read of every step's loss inside the timed region
  roofline  the dominant hand-written kernel (the selective scan): algorithmic bytes / launch divided by
            its average CUDA-event duration, against MEASURED_PEAKS.json.  The timed steps are CUDA-graph replays, whose
            kernels cannot carry events: the events come from 3 EAGER launches of the same step run right after the
            timed region (`config.roofline_timing` says so in the line itself); launches are grouped by (entry point,
            shape) so a model with several scan shapes (S2S) still quotes one kernel at one shape
  cpu_baseline  the reference CPU path (oracle port of selective_scan_ref + torch conv + Fbank inside the same
            module tree) on the host cores, bounded sample
``--impl reference`` times that CPU path alone (rank 0 only).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

METRIC = "ConMamba encoder audio-sec/sec (fwd+bwd)"
UNIT = "audio-sec/sec"

# mode: "train" = forward + loss + backward + optimizer ; "fwd" = forward + loss under no_grad (evaluation / inference)
WORKLOADS = {
    # BASELINE.json configs[0]: ConMamba-small CTC encoder forward, 8 x 10 s (the reference's own CPU-runnable case)
    "conmamba_small_ctc_fwd_b8x10s": dict(model="conmamba_small_ctc", batch=8, seconds=10.0, mode="fwd"),
    # BASELINE.json configs[1]: ConMamba-small CTC encoder fwd+bwd, 32 x 15 s on one GPU
    "conmamba_small_ctc_fwdbwd_b32x15s": dict(model="conmamba_small_ctc", batch=32, seconds=15.0, mode="train"),
    # BASELINE.json configs[2] - the configuration the metric and the target are quoted on: ConMamba-large CTC training
    # step, 64 x 20 s per GPU, data-parallel at 1/2/4/8 GPUs
    "conmamba_large_ctc_fwdbwd_b64x20s": dict(model="conmamba_large_ctc", batch=64, seconds=20.0, mode="train"),
    # BASELINE.json configs[3] shape: ConMambaMamba-large S2S (12 ConMamba encoder + 6 Mamba decoder layers, d_model 512,
    # vocab 5000) training step, 64 x 20 s per GPU, 3 target tokens per audio second
    "conmambamamba_large_s2s_fwdbwd_b64x20s": dict(model="conmambamamba_large_s2s", batch=64, seconds=20.0, mode="train"),
    # BASELINE.json configs[4]: long-form ConMamba-large encoder inference, 4 x 300 s (7501 encoder frames)
    "conmamba_large_ctc_infer_b4x300s": dict(model="conmamba_large_ctc", batch=4, seconds=300.0, mode="fwd"),
    # small smoke shapes for debugging
    "tiny": dict(model="conmamba_small_ctc", batch=2, seconds=2.0, mode="train"),
    "tiny_fwd": dict(model="conmamba_small_ctc", batch=2, seconds=2.0, mode="fwd"),
    "tiny_s2s": dict(model="conmambamamba_large_s2s", batch=2, seconds=2.0, mode="train"),
}
DEFAULT_WORKLOAD = "conmamba_large_ctc_fwdbwd_b64x20s"
# optimizer of the training workloads (hparams/CTC/conmamba_large.yaml:91-99, 244-252): AdamW, Noam schedule, clipping
OPT = dict(lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=5e-4, max_grad_norm=5.0, n_warmup_steps=7500)


# ------------------------------------------------------------------------------------------------ helpers
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(entry, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `entry` at this workload's shape, from the committed
    `ncu --set full` capture of the kernel that ships (profiles/r02_traffic.json, written by tools/ncu_traffic.py; the
    round-1 file holds the round-1 kernels); None if not captured."""
    key = "%s@%s" % (entry, workload)
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            continue
        if key in d:
            return d[key]
    return None


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for n in dir(nv):
            if n.startswith("nvmlClocksEventReason") or n.startswith("nvmlClocksThrottleReason"):
                v = getattr(nv, n)
                if isinstance(v, int) and v not in (0,):
                    names.setdefault(v, n.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if bit and (mask & bit) == bit and bin(bit).count("1") == 1:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        reasons = sorted(r for r in self.reasons if r not in ("GpuIdle", "None", "ApplicationsClocksSetting"))
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples)}


def make_batch(model_cfg, batch, seconds, seed, device, n_classes):
    g = torch.Generator().manual_seed(seed)
    n = int(round(16000 * seconds))
    wav = 0.1 * torch.randn(batch, n, generator=g)
    tgt_len = max(1, int(3 * seconds))
    targets = torch.randint(3, n_classes, (batch, tgt_len), generator=g)
    return wav, targets


def ctc_loss_of(logp, targets):
    """GPU arm: cm_ctc_loss (alpha, beta and the gradient in one launch); the CPU reference arm keeps torch's ctc_loss.
    CM_BENCH_TORCH_CTC=1 is the A/B switch back to torch's three kernels on the GPU."""
    Bt, L, _ = logp.shape
    if logp.is_cuda and os.environ.get("CM_BENCH_TORCH_CTC", "0") != "1":
        from mamba_asr_b200.ctc import ctc_loss
        in_len = torch.full((Bt,), L, dtype=torch.long, device=logp.device)
        tg_len = torch.full((Bt,), targets.shape[1], dtype=torch.long, device=logp.device)
        return ctc_loss(logp.transpose(0, 1), targets, in_len, tg_len, blank=0, reduction="mean", zero_infinity=True)
    in_len = torch.full((Bt,), L, dtype=torch.long)
    tg_len = torch.full((Bt,), targets.shape[1], dtype=torch.long)
    return F.ctc_loss(logp.float().transpose(0, 1), targets, in_len, tg_len, blank=0, reduction="mean",
                      zero_infinity=True)


def step_loss(model, wav, targets, autocast, is_s2s=None):
    """CTC recipe (train_CTC.py:285-302).  S2S recipe (train_S2S.py:344-361, 518-530): ctc_weight * CTC on the encoder +
    (1 - ctc_weight) * label-smoothed KL on the decoder, decoder input = <bos> + targets, decoder target = targets + <eos>."""
    if is_s2s is None:
        is_s2s = hasattr(model, "decoder")
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        if is_s2s:
            bos = torch.cat([torch.ones_like(targets[:, :1]), targets], dim=1)
            p_ctc, p_seq = model(wav, bos)
        else:
            logp = model(wav)                                       # (B, L, C)
    if is_s2s:
        from mamba_asr_b200.encoder import kldiv_loss
        eos = torch.cat([targets, torch.full_like(targets[:, :1], 2)], dim=1)
        loss = 0.3 * ctc_loss_of(p_ctc, targets) + 0.7 * kldiv_loss(p_seq.float(), eos, label_smoothing=0.1)
    else:
        loss = ctc_loss_of(logp, targets)
    return loss


def ctc_step(model, wav, targets, autocast, train=True, is_s2s=None):
    if not train:                       # evaluation / inference workloads: forward + loss, no gradients
        with torch.no_grad():
            return step_loss(model, wav, targets, autocast, is_s2s)
    loss = step_loss(model, wav, targets, autocast, is_s2s)
    loss.backward()
    return loss


def config_block(workload, world):
    """The part of `config` that names the workload - identical in the GPU arm and in the reference arm."""
    from mamba_asr_b200.encoder import CONFIGS
    wl = WORKLOADS[workload]
    cfg = CONFIGS[wl["model"]]
    T = 1 + int(round(16000 * wl["seconds"])) // 160
    L = (((T - 1) // 2 + 1) - 1) // 2 + 1
    layers = "%d" % cfg["num_layers"] + ("+%d decoder" % cfg["num_decoder_layers"] if "num_decoder_layers" in cfg else "")
    if wl["mode"] == "train":
        step = ("Fbank+norm+CNN+%s ConMamba layers+loss, forward+backward, bf16 autocast, clip %.1f + AdamW + Noam"
                % (layers, OPT["max_grad_norm"]))
    else:
        step = "Fbank+norm+CNN+%s ConMamba layers+loss, forward only (no_grad), bf16 autocast" % layers
    return {"workload": workload, "per_gpu_batch": wl["batch"], "audio_seconds": wl["seconds"], "encoder_frames": L,
            "d_inner": 2 * cfg["d_model"], "layers": cfg["num_layers"], "mode": wl["mode"], "step": step,
            "parallelism": ("dp%d (utterance sharding; gradient all-reduce over NCCL)" % world) if world > 1 else "single GPU"}


def scan_algorithmic_bytes(batch, L, D, N, s, ndir, bwd):
    """bytes per (b, d, l) position of one fused-bidirectional scan launch."""
    if ndir == 2:
        per = (11 + 8.0 * N / D) * s if bwd else (6 + 4.0 * N / D) * s
    else:
        per = (7 + 4.0 * N / D) * s if bwd else (4 + 2.0 * N / D) * s
    return per * batch * L * D


def conv_algorithmic_bytes(batch, L, D, s, ndir, bwd):
    per = (4 if bwd else 3) * s if ndir == 2 else (3 if bwd else 2) * s
    return per * batch * L * D


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_run(workload, steps, warmup, budget_s=200.0):
    """Times the reference CPU path on a bounded sample of the workload; returns (audio-s/s, description dict)."""
    from mamba_asr_b200.encoder import CONFIGS, build_model
    from oracle.cpu_encoder import to_cpu_reference
    wl = WORKLOADS[workload]
    cfg = CONFIGS[wl["model"]]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = to_cpu_reference(build_model(wl["model"]), cfg["n_fft"], cfg["n_mels"], cfg["win_length"])
    train = wl["mode"] == "train"
    model.train(train)
    seconds = wl["seconds"]
    opt = None
    if train:
        from mamba_asr_b200.trainer import TrainStep
        opt = TrainStep(model, **OPT)

    def one(sec):
        wav, targets = make_batch(cfg, 1, sec, 1234, "cpu", cfg["output_neurons"])
        t0 = time.perf_counter()
        if train:
            model.zero_grad(set_to_none=True)
            step_loss(model, wav, targets, False).backward()
            opt.step()
        else:
            with torch.no_grad():
                step_loss(model, wav, targets, False)
        return time.perf_counter() - t0

    # pick the largest sample (1 utterance of the workload's duration, else a shorter cut) that fits the budget
    sec = seconds
    t_probe = one(sec)                                        # also the first warm-up step
    total = steps + warmup
    while t_probe * total > budget_s and sec > 1.0:
        sec = max(1.0, sec / 2.0)
        t_probe = one(sec)
    for _ in range(max(0, warmup - 1)):
        one(sec)
    times = [one(sec) for _ in range(steps)]
    dt = sum(times)
    value = sec * steps / dt
    desc = {"kind": "port", "cores": threads, "value": value, "unit": UNIT,
            "sample": "1 utterance x %.2f s of the %s workload per step (the workload's batch is %d such utterances; on the CPU path "
                      "audio-s/s moves by about 20 %% between batch 1 and 4, measured): oracle port of selective_scan_ref + torch conv + Fbank "
                      "inside the same %d-layer module tree, %s, fp32, %d steps, %.2f s/step"
                      % (sec, workload, wl["batch"], cfg["num_layers"],
                         "fwd+bwd+optimizer" if train else "forward only", steps, dt / steps)}
    return value, dt / steps * 1e3, desc


# ------------------------------------------------------------------------------------------------ scan sweep (config 5)
def scan_length_sweep(dev, peak, lengths=(1024, 2048, 4096, 8192, 16384, 30001), iters=10):
    """BASELINE.json configs[4] kernel-level part: the fused bidirectional scan forward (inference launch: no
    checkpoints, chunk-parallel over time windows when that helps) at batch 4 x D 512 for L = 1k .. 30k, bf16 and fp32;
    algorithmic GB/s against the measured HBM peak.  L2 is flushed before every launch."""
    from mamba_asr_b200 import kernels as K
    Bt, D, N = 4, 512, 16
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = []
    for dt, s in ((torch.bfloat16, 2), (torch.float32, 4)):
        for L in lengths:
            g = torch.Generator(device=dev).manual_seed(L)
            rn = lambda *sh: torch.randn(*sh, device=dev, generator=g)
            cl = lambda: rn(Bt, L, D).to(dt).transpose(1, 2)
            z = cl()
            dirs = []
            for rev in (False, True):
                xdbl = rn(Bt, L, 2 * N + 16).to(dt)
                dirs.append(dict(u=cl(), delta=(0.5 * rn(Bt, L, D)).to(dt).transpose(1, 2), A=-torch.exp(0.3 * rn(D, N)),
                                 B=xdbl[..., :N].transpose(1, 2), C=xdbl[..., N:2 * N].transpose(1, 2),
                                 D=torch.ones(D, device=dev), delta_bias=torch.full((D,), -4.0, device=dev), reverse=rev))
            f = lambda: K.scan_forward(dirs, z=z, out_scale=0.5, delta_softplus=True)
            for _ in range(3):
                f()
            ts = []
            for _ in range(iters):
                flush.zero_()
                K.start_timing()
                f()
                ts += K.stop_timing()["cm_scan_fwd"]
            ts.sort()
            byts = scan_algorithmic_bytes(Bt, L, D, N, s, 2, False)
            med = ts[len(ts) // 2]
            out.append({"dtype": "bf16" if s == 2 else "f32", "L": L, "ms": med, "alg_bytes": byts,
                        "achieved_gbs": byts / (med * 1e-3) / 1e9, "frac": byts / (med * 1e-3) / 1e9 / peak})
    return out


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true", help="(default) replay the model forward / backward as CUDA graphs: the eager step is host-launch-bound (2.4 k launches, 41 ms wall for 30 ms of kernels on B200)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly")
    ap.add_argument("--no-defer-reduce", action="store_true",
                    help="A/B: reduce every backward kernel's partial sums where they are produced (one launch per operator) "
                         "instead of in one batch at the end of the backward graph")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-param-cache", action="store_true", help="per-use autocast-style parameter casts (A/B)")
    ap.add_argument("--no-optimizer", action="store_true", help="training workloads: stop after backward (A/B)")
    ap.add_argument("--sweep-L", action="store_true", help="add the kernel-level scan sweep over L = 1k..30k (configs[4]) "
                                                           "to the JSON line as `scan_sweep`")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    train = wl["mode"] == "train"

    if args.impl == "reference":
        if rank != 0:
            return 0
        value, ms, desc = cpu_reference_run(args.workload, args.steps, max(1, args.warmup))
        line = {"impl": "reference", "metric": METRIC if train else METRIC.replace("(fwd+bwd)", "(fwd)"), "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_block(args.workload, max(1, args.gpus)),
                "cpu_baseline": desc,
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: the product path has no CPU fallback")
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from mamba_asr_b200 import kernels as K
    from mamba_asr_b200.encoder import CONFIGS, build_model
    from mamba_asr_b200.trainer import TrainStep
    cfg = CONFIGS[wl["model"]]
    model = build_model(wl["model"]).to(dev)
    model.train(train)
    if not args.no_param_cache:
        model.enable_param_cache()          # bf16 parameter copies refreshed by one multi-tensor copy per step
    net = model
    if world > 1 and train:
        from mamba_asr_b200.dist_utils import allreduce_gradients
        for p_ in model.parameters():                          # identical replicas: rank 0's initialisation
            dist.broadcast(p_.data, 0)
    n_params = sum(p.numel() for p in model.parameters())
    is_s2s = hasattr(model, "decoder")
    opt = TrainStep(model, world_size=world, **OPT) if (train and not args.no_optimizer) else None

    batch, seconds = wl["batch"], wl["seconds"]
    wav_h, tgt_h = make_batch(cfg, batch, seconds, cfg["seed"] + rank, dev, cfg["output_neurons"])
    wav_pin, tgt_pin = wav_h.pin_memory(), tgt_h.pin_memory()
    wav_d, tgt_d = wav_h.to(dev), tgt_h.to(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    eager_forward = net.forward      # torch.cuda.make_graphed_callables swaps net.forward for the graph replay

    def finish_step():
        """what follows backward in a training step: gather the gradients into the flat buffer, the step's only collective
        (one in-place NCCL all-reduce of that buffer over NVLink), then clip + AdamW (two launches)"""
        if opt is not None:
            opt.step(dist if world > 1 else None)
        elif world > 1:
            allreduce_gradients(model.parameters(), world)

    def step_eager(w, t):
        if not train:
            return ctc_step(net, w, t, True, train=False)
        model.zero_grad(set_to_none=True)
        cur, net.forward = net.forward, eager_forward
        try:
            loss = ctc_step(net, w, t, True)
        finally:
            net.forward = cur
        finish_step()
        return loss

    # CUDA graphs: the eager step is host-launch-bound.  Training: forward graph + backward graph of the whole model
    # (torch.cuda.make_graphed_callables); the CTC loss between them stays eager because its length tensors live on the
    # host.  Forward-only workloads: one graph of the model forward.  Eager fallback if the capture is refused.
    graph_note = "eager launches"
    step_fn = step_eager
    launches_per_step = None
    if not args.no_graph:
        try:
            from mamba_asr_b200.graphs import graph_forward, graph_module
            l0 = K.LAUNCHES
            sample = (wav_d,) if not is_s2s else (wav_d, torch.cat([torch.ones_like(tgt_d[:, :1]), tgt_d], dim=1))
            if train:
                # capture under the same autocast policy the step runs with (weight-cast caching off: the cached casts
                # of a warm-up iteration would otherwise be baked out of the graph)
                # the partial sums of all backward kernels are reduced in one batch at the end of the backward graph
                # (kernels.deferred_reductions: valid here because the graph hands the gradients out after the pass)
                with torch.autocast("cuda", dtype=torch.bfloat16, cache_enabled=False), \
                        K.deferred_reductions(not args.no_defer_reduce):
                    gnet = graph_module(net, sample, warmup=3)
                launches_per_step = (K.LAUNCHES - l0) // 4          # 3 warm-ups + 1 capture

                def step_graphed(w, t):
                    model.zero_grad(set_to_none=True)
                    loss = ctc_step(gnet, w, t, True)
                    finish_step()
                    return loss
                graph_note = "model forward and backward replayed as CUDA graphs (loss, collective and optimizer eager)"
            else:
                def fwd_only(*xs):
                    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, cache_enabled=False):
                        return net(*xs)
                grun = graph_forward(fwd_only, [x.clone() for x in sample], warmup=3)
                launches_per_step = (K.LAUNCHES - l0) // 4

                def step_graphed(w, t):
                    return ctc_step(grun, w, t, True, train=False, is_s2s=is_s2s)
                graph_note = "model forward replayed as one CUDA graph (loss eager)"
            step_graphed(wav_d, tgt_d)
            torch.cuda.synchronize()
            step_fn = step_graphed
        except Exception as ex:
            torch.cuda.synchronize()
            graph_note = "eager launches (graph capture refused: %s)" % (repr(ex)[:160])
            step_fn = step_eager

    for _ in range(warmup):
        step_fn(wav_d, tgt_d)
    barrier()
    torch.cuda.reset_peak_memory_stats()

    # ---- timed region: K steps, device events -------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    launches0 = K.LAUNCHES
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step_fn(wav_d, tgt_d)
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = (K.LAUNCHES - launches0) if launches_per_step is None else launches_per_step * steps
    elapsed_ms = e0.elapsed_time(e1)
    peak_mem = torch.cuda.max_memory_allocated()
    loss_val = float(loss.item())

    # ---- per-kernel CUDA events for the roofline: the same step, launched eagerly right after the timed region ------
    for _ in range(2):
        step_eager(wav_d, tgt_d)
    K.start_timing()
    for _ in range(3):
        step_eager(wav_d, tgt_d)
    ktimes = K.stop_timing(by_shape=True)

    # ---- end-to-end: pinned host audio -> H2D -> step -> loss D2H, every step ---------------------------------
    # Every step's inputs come from pinned host memory and its loss goes back to the host, all inside the timed region.
    # The copy of step i+1 is issued on a copy stream while step i computes (two device buffers, events both ways), as
    # a data loader feeding a trainer would.
    copy_stream = torch.cuda.Stream()
    bufs = [(torch.empty_like(wav_d), torch.empty_like(tgt_d)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def issue_copy(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])              # the step that last read this buffer is done
            bufs[i % 2][0].copy_(wav_pin, non_blocking=True)     # host -> device copy of step i's inputs
            bufs[i % 2][1].copy_(tgt_pin, non_blocking=True)
            ready[i % 2].record(copy_stream)

    loss_pin = torch.zeros(max(steps, 2), dtype=torch.float32).pin_memory()

    def run_e2e(n):
        cur = torch.cuda.current_stream()
        for b in range(2):
            consumed[b].record(cur)
        issue_copy(0)
        for i in range(n):
            if i + 1 < n:
                issue_copy(i + 1)
            cur.wait_event(ready[i % 2])
            loss_i = step_fn(*bufs[i % 2])
            consumed[i % 2].record(cur)
            # device -> host read of every step's loss: stream-ordered copy into pinned memory (a trainer's logging
            # path), so the host does not stall the launches of the next step; all n values are on the host when the
            # closing event of the timed region has completed
            loss_pin[i:i + 1].copy_(loss_i.detach().float().reshape(1), non_blocking=True)

    run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    run_e2e(steps)
    g1.record()
    barrier()
    e2e_ms = g0.elapsed_time(g1)
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_losses = [float(v) for v in loss_pin[:steps]]
    assert all(v == v and abs(v) < 1e30 for v in e2e_losses), e2e_losses

    if dist is not None:
        t = torch.tensor([elapsed_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = float(t[0]), float(t[1])
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt)

    audio_s = batch * seconds * world * steps
    value = audio_s / (elapsed_ms / 1e3)
    e2e_value = audio_s / (e2e_ms / 1e3)

    if rank == 0:
        # ---- roofline of the hand-written kernels (rank 0's launches), grouped by (entry point, launch shape) ----------
        hbm_peak, peak_src = peaks()
        N, s = 16, 2
        alg_of = {"cm_scan_fwd": lambda b, d, l, nd: scan_algorithmic_bytes(b, l, d, N, s, nd, False),
                  "cm_scan_bwd": lambda b, d, l, nd: scan_algorithmic_bytes(b, l, d, N, s, nd, True),
                  "cm_conv_fwd": lambda b, d, l, nd: conv_algorithmic_bytes(b, l, d, s, nd, False),
                  "cm_conv_bwd": lambda b, d, l, nd: conv_algorithmic_bytes(b, l, d, s, nd, True)}
        kern, share = {}, {}
        for (name, tag), ts in ktimes.items():
            share[name] = share.get(name, 0.0) + sum(ts)
            if name in alg_of and tag is not None and ts:
                byts = alg_of[name](*tag)
                avg = sum(ts) / len(ts)
                kern["%s@B%d_D%d_L%d_dirs%d" % ((name,) + tuple(tag))] = {
                    "entry": name, "launches": len(ts), "avg_ms": avg, "total_ms": sum(ts), "alg_bytes": byts,
                    "achieved_gbs": byts / (avg * 1e-3) / 1e9, "frac": byts / (avg * 1e-3) / 1e9 / hbm_peak}
        dom = max(kern, key=lambda k: kern[k]["total_ms"]) if kern else None
        roofline = None
        if dom:
            roofline = {"kernel": dom, "bound": "hbm", "achieved": kern[dom]["achieved_gbs"], "peak": hbm_peak,
                        "unit": "GB/s", "frac": kern[dom]["frac"], "traffic": ncu_traffic(kern[dom]["entry"], args.workload),
                        "peak_source": peak_src, "alg_bytes_per_launch": kern[dom]["alg_bytes"],
                        "avg_launch_ms": kern[dom]["avg_ms"],
                        "note": "fp32 state update is MUFU/issue-bound before HBM (SURVEY.md 0.8); see `kernels`"}
        cpu_desc = None
        if not args.no_cpu_baseline and world == 1:
            try:
                _, _, cpu_desc = cpu_reference_run(args.workload, args.cpu_steps, 1, budget_s=60.0)
            except Exception as ex:                               # the baseline must never take the GPU number down
                cpu_desc = {"kind": "port", "error": repr(ex)}
        config = config_block(args.workload, world)
        config.update({"params": n_params, "launch": graph_note,
                       "optimizer": ("flat-buffer clip + AdamW kernels (cm_sumsq_partial, cm_adamw_step) + Noam lr, inside the timed step" if opt is not None
                                     else ("none (forward-only workload)" if not train else "disabled (--no-optimizer)")),
                       "roofline_timing": "per-kernel CUDA events from 3 eager launches of the same step right after "
                                          "the timed region (kernels inside a graph replay cannot carry events)",
                       "l2": "no flush: activations touched per step (peak %.2f GB allocated) exceed the 126 MB L2"
                             % (peak_mem / 1e9)})
        line = {
            "metric": METRIC if train else METRIC.replace("(fwd+bwd)", "(fwd)"), "value": value, "unit": UNIT,
            "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": elapsed_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / steps,
                    "host_wall_ms_per_step": e2e_wall_ms / steps,
                    "h2d_bytes_per_step": int(wav_pin.numel() * 4 + tgt_pin.numel() * 8), "d2h_bytes_per_step": 4,
                    "pipeline": "step i+1's H2D on a copy stream under step i; loss D2H stream-ordered into pinned memory"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "kernels": kern,
            "kernel_time_share_ms": share,
            "cpu_baseline": cpu_desc,
            "loss": loss_val,
        }
        if args.sweep_L:
            line["scan_sweep"] = scan_length_sweep(dev, hbm_peak)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
