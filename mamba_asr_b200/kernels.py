"""Tensor-level launchers over the C ABI (include/conmamba_b200.h).

Each function takes torch CUDA tensors, fills the POD argument block with raw pointers / element strides and
enqueues the sm_100a kernels on torch's current stream.  Logical shapes follow the reference
(u, delta, z: (B, D, L); variable B, C: (B, N, L)), but ANY strides are accepted - channel-last memory
(a ``(B, L, D)`` buffer viewed through ``.transpose(1, 2)``) is the fast path.  No host synchronisation.
"""
import ctypes as C
import functools
import os

import torch

from . import _cabi as cabi

__all__ = ["scan_forward", "scan_backward", "conv_forward", "conv_backward", "conv_update", "fbank_logmel",
           "empty_like_bdl", "num_ckpt"]


# ---- launch accounting (bench.py reads these; they never change what is computed) -------------------------
LAUNCHES = 0          # kernels launched through the C ABI since import (one per entry-point call)
_TIMING = None        # None, or {entry point: [(start_event, end_event), ...]} while bench.py profiles a step


def start_timing():
    """Record a CUDA event pair around every C-ABI launch on the launching stream (used by bench.py's roofline)."""
    global _TIMING
    _TIMING = {}


def stop_timing(by_shape=False):
    """-> {entry point: [ms per launch, ...]} ; synchronises.  ``by_shape=True`` keys the scan / conv launches by
    (entry point, (batch, dim, seqlen, ndir)) instead, so a model that launches one entry point at several shapes (the
    S2S decoder's unidirectional scans next to the encoder's bidirectional ones) can quote one kernel at one shape."""
    global _TIMING
    t, _TIMING = _TIMING, None
    torch.cuda.synchronize()
    out = {}
    for (name, tag), v in (t or {}).items():
        key = (name, tag) if by_shape else name
        out.setdefault(key, []).extend(a.elapsed_time(b) for a, b in v)
    return out


_NVTX = os.environ.get("CM_NVTX") not in (None, "", "0")   # CM_NVTX=1: one NVTX range per C-ABI entry point (nsys / ncu --nvtx)


def _call(name, fn, *args, tag=None):
    global LAUNCHES
    LAUNCHES += 1
    if _NVTX:
        torch.cuda.nvtx.range_push(name if tag is None else "%s %s" % (name, tag))
        try:
            cabi.check(fn(*args), name)
        finally:
            torch.cuda.nvtx.range_pop()
        return
    if _TIMING is None:
        cabi.check(fn(*args), name)
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    cabi.check(fn(*args), name)
    e.record()
    _TIMING.setdefault((name, tag), []).append((s, e))


def grad_cast(t, dtype):
    """``t.to(dtype)`` for a gradient that may still be queued (deferred_reductions): a real cast reads it, so the queue is
    run first; the usual fp32 -> fp32 case returns ``t`` itself and leaves the queue alone."""
    if t is None or t.dtype == dtype:
        return t
    flush_reductions()
    return t.to(dtype)


# ---- fixed-order sums of partial buffers, optionally queued until the end of the backward pass --------------------------
# A job is (part, out) - part (rows, ...) fp32 contiguous, out fp32 with part[0].numel() elements - or
# (part, out, rows, cols, stride) with explicit geometry (part[r * stride + c], c < cols).
_DEFER = False          # set by deferred_reductions()
_DEFER_POISON = os.environ.get("CM_DEFER_POISON") not in (None, "", "0")   # tests: NaN-fill queued outputs until the flush
_PENDING = []


class deferred_reductions:
    """Context manager: inside it, ``reduce_many(jobs, defer=True)`` calls made during a backward pass are queued and run in
    ONE batch (cm_reduce_batch, 64 jobs per launch) by an autograd-engine callback at the end of that pass, instead of one
    reducer launch per operator (440 launches of ~4 us per ConMamba-large step).

    Contract (why this is opt-in): a queued output holds no data until the pass ends, so it must not be READ inside the pass.
    The library's own backward functions only mark jobs whose output goes straight to a parameter gradient; the caller
    guarantees that nothing accumulates into those gradients during the pass: use it around ``torch.autograd.grad`` /
    ``torch.cuda.make_graphed_callables`` capture (gradients are handed out at the end of the pass), with every parameter
    used by one autograd node.  Plain ``loss.backward()`` into existing ``.grad`` tensors (accumulation) is NOT covered."""

    def __init__(self, enabled=True):
        self.enabled = enabled

    def __enter__(self):
        global _DEFER
        self.prev, _DEFER = _DEFER, self.enabled
        return self

    def __exit__(self, exc_type, *exc):
        global _DEFER
        _DEFER = self.prev
        if exc_type is not None:
            _PENDING.clear()            # a pass that raised: its queued jobs point at tensors nobody will use
        flush_reductions()
        return False


def _norm_job(job):
    if len(job) == 2:
        part, out = job
        rows = part.shape[0]
        cols = part.numel() // rows
        return part, out, part.data_ptr(), out.data_ptr(), rows, cols, cols
    part, out, rows, cols, stride, off = job
    return part, out, part.data_ptr() + 4 * off, out.data_ptr(), rows, cols, stride


def _launch_reduce(jobs):
    lib = cabi.lib()
    st = cabi.stream_ptr()
    for i in range(0, len(jobs), cabi.CM_REDUCE_BATCH_MAX):
        chunk = jobs[i:i + cabi.CM_REDUCE_BATCH_MAX]
        arr = (cabi.ReduceJob2 * len(chunk))()
        for k, (_, _, pp, op, rows, cols, stride) in enumerate(chunk):
            arr[k].part, arr[k].out, arr[k].rows, arr[k].cols, arr[k].stride = pp, op, rows, cols, stride
        _call("cm_reduce_batch", lib.cm_reduce_batch, arr, len(chunk), st)


def flush_reductions():
    """Run every queued reduction now (the engine callback; also safe to call by hand, and a no-op on an empty queue)."""
    if _PENDING:
        jobs = list(_PENDING)
        _PENDING.clear()
        _launch_reduce(jobs)


def reduce_many(jobs, defer=False):
    """Fixed-order column sums of partial buffers (cm_reduce_batch).  ``defer=True`` marks outputs that nobody reads before the
    end of the current backward pass: under ``deferred_reductions()`` they are queued (see there), otherwise run now."""
    jobs = [_norm_job(j) for j in jobs]
    if not jobs:
        return
    if defer and _DEFER:
        queued = bool(_PENDING)                     # a non-empty queue already has its end-of-pass callback
        if not queued:
            try:                                    # (one more callback after a mid-pass flush is a no-op at the end)
                torch.autograd.Variable._execution_engine.queue_callback(flush_reductions)
                queued = True
            except RuntimeError:                    # not inside a backward pass: nothing would ever flush the queue
                pass
        if queued:
            if _DEFER_POISON:
                for j in jobs:
                    j[1].view(-1)[:j[5]].fill_(float("nan"))
            _PENDING.extend(jobs)                   # the tensors stay referenced (and their memory theirs) until the flush
            return
    _launch_reduce(jobs)


def sum_leading(part, defer=False):
    """part.sum(0) for a contiguous fp32 (rows, ...) CUDA tensor through the batched reducer (the split-K weight-gradient
    sums: 219 at::reduce_kernel launches per ConMamba-large step otherwise)."""
    out = torch.empty(part.shape[1:], dtype=torch.float32, device=part.device)
    reduce_many([(part, out)], defer=defer)
    return out


def _require_cuda(t, name, like=None):
    if not t.is_cuda:
        raise RuntimeError("mamba_asr_b200: %s must be a CUDA tensor - the B200 kernels have no CPU fallback" % name)
    if like is not None and t.device != like.device:
        raise RuntimeError("mamba_asr_b200: %s is on %s but the launch runs on %s" % (name, t.device, like.device))


def _same_device(like, **tensors):
    """Every (non-None) parameter / buffer of a launch must live on the device of its activations: a host or foreign-device
    pointer would otherwise reach the kernel and fault the context instead of raising here."""
    for name, t in tensors.items():
        if t is not None:
            _require_cuda(t, name, like)


def empty_like_bdl(t, dtype=None):
    """New (B, D, L) tensor with t's memory order (channel-last stays channel-last)."""
    Bt, D, L = t.shape
    dtype = dtype or t.dtype
    if t.stride(1) == 1 and D > 1:
        return torch.empty((Bt, L, D), dtype=dtype, device=t.device).transpose(1, 2)
    return torch.empty((Bt, D, L), dtype=dtype, device=t.device)


def num_ckpt(L, ndir):
    return cabi.lib().cm_scan_num_ckpt(L, ndir)


def _f32c(t, name):
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32 (reference keeps A, D, delta_bias in fp32, bimamba.py:200,232-233)" % name)
    return t if t.is_contiguous() else t.contiguous()


def _fill_scan_dir(sd, d, keep, const_bc):
    u, delta = d["u"], d["delta"]
    sd.reverse = 1 if d.get("reverse", False) else 0
    sd.bc_const = 1 if const_bc else 0
    sd.u = cabi.t3(u)
    sd.delta = cabi.t3(delta)
    Bm, Cm = d["B"], d["C"]
    if const_bc:
        # (D, N) fp32 constants: the "batch" stride slot carries the channel stride
        Bm, Cm = Bm.float().contiguous(), Cm.float().contiguous()
        keep += [Bm, Cm]
        sd.Bm = cabi.Tensor3(Bm.data_ptr(), Bm.stride(0), Bm.stride(1), 0)
        sd.Cm = cabi.Tensor3(Cm.data_ptr(), Cm.stride(0), Cm.stride(1), 0)
    else:
        sd.Bm = cabi.t3(Bm)
        sd.Cm = cabi.t3(Cm)
    A = _f32c(d["A"], "A")
    Dk = _f32c(d.get("D"), "D")
    bias = _f32c(d.get("delta_bias"), "delta_bias")
    _same_device(u, A=A, D=Dk, delta_bias=bias, B=Bm, C=Cm, delta=delta)
    keep += [A, Dk, bias]
    sd.A = A.data_ptr()
    sd.A_sd, sd.A_sn = A.stride(0), A.stride(1)
    sd.Dskip = cabi.ptr(Dk)
    sd.delta_bias = cabi.ptr(bias)


def _check_dirs(dirs):
    if len(dirs) not in (1, 2):
        raise ValueError("1 or 2 scan directions")
    u0 = dirs[0]["u"]
    Bt, D, L = u0.shape
    N = dirs[0]["A"].shape[1]
    const_bc = dirs[0]["B"].dim() == 2
    for d in dirs:
        for k in ("u", "delta", "B", "C"):
            _require_cuda(d[k], k)
        if d["u"].shape != (Bt, D, L) or d["delta"].shape != (Bt, D, L):
            raise ValueError("u / delta must be (B, D, L)")
        if d["u"].dtype != u0.dtype or d["delta"].dtype != u0.dtype:
            raise TypeError("u and delta must share one dtype")
        if d["A"].shape != (D, N):
            raise ValueError("A must be (D, N)")
        if (d["B"].dim() == 2) != const_bc or (d["C"].dim() == 2) != const_bc:
            raise NotImplementedError("B and C must both be input-dependent (B, N, L) or both constant (D, N)")
        if const_bc:
            if d["B"].shape != (D, N) or d["C"].shape != (D, N):
                raise ValueError("constant B / C must be (D, N)")
        else:
            if d["B"].shape != (Bt, N, L) or d["C"].shape != (Bt, N, L):
                raise ValueError("variable B / C must be (B, N, L)")
        if not const_bc and (d["B"].dtype != u0.dtype or d["C"].dtype != u0.dtype):
            raise TypeError("input-dependent B and C must have the dtype of u")
    return Bt, D, L, N, const_bc


def scan_forward(dirs, z=None, out_scale=1.0, delta_softplus=False, need_ckpt=False, need_last_state=False,
                 need_out_pre=False, lanes=0):
    """Fused selective scan forward (cm_scan_fwd).

    dirs: list of 1 or 2 dicts {u, delta, A, B, C, D=None, delta_bias=None, reverse=False}.
    Returns dict(out, out_pre, ckpt[list], last_state[list]).
    """
    lib = cabi.lib()
    Bt, D, L, N, const_bc = _check_dirs(dirs)
    u0 = dirs[0]["u"]
    a = cabi.ScanFwdArgs()
    a.batch, a.dim, a.seqlen, a.dstate = Bt, D, L, N
    a.ndir = len(dirs)
    a.dtype = cabi.dtype_code(u0.dtype)
    a.flags = cabi.CM_FLAG_DELTA_SOFTPLUS if delta_softplus else 0
    a.out_scale = float(out_scale)
    a.lanes_per_channel = int(lanes) or int(os.environ.get("CM_SCAN_LANES", "0"))   # 0 = library heuristic
    keep = []
    ckpts, lasts = [], []
    nck = lib.cm_scan_num_ckpt(L, len(dirs))
    for r, d in enumerate(dirs):
        sd = a.dir[r]
        _fill_scan_dir(sd, d, keep, const_bc)
        if need_ckpt:
            ck = torch.empty((Bt, D, nck, 16), dtype=torch.float32, device=u0.device)
            sd.ckpt = ck.data_ptr()
            sd.ckpt_sb, sd.ckpt_sd = ck.stride(0), ck.stride(1)
            ckpts.append(ck)
        if need_last_state:
            ls = torch.empty((Bt, D, N), dtype=torch.float32, device=u0.device)
            sd.last_state = ls.data_ptr()
            sd.ls_sb, sd.ls_sd, sd.ls_sn = ls.stride()
            lasts.append(ls)
    if z is not None:
        _require_cuda(z, "z")
        if z.shape != (Bt, D, L) or z.dtype != u0.dtype:
            raise ValueError("z must match u in shape and dtype")
    a.z = cabi.t3(z)
    out = empty_like_bdl(u0)
    a.out = cabi.t3(out)
    out_pre = empty_like_bdl(u0) if need_out_pre else None
    a.out_pre = cabi.t3(out_pre)
    ws_bytes = lib.cm_scan_fwd_workspace_bytes(C.byref(a))   # > 0: few long sequences, no checkpoints -> time windows
    if ws_bytes > 0:
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=u0.device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
        keep.append(ws)
    _call("cm_scan_fwd", lib.cm_scan_fwd, C.byref(a), cabi.stream_ptr(), tag=(Bt, D, L, len(dirs)))
    return dict(out=out, out_pre=out_pre, ckpt=ckpts, last_state=lasts)


def scan_backward(dirs, ckpts, dout, z=None, out_pre=None, out_scale=1.0, delta_softplus=False, lanes=0,
                  dz_out=None, dBC_like=None, defer=False, defer_dA=False):
    """Fused selective scan backward (cm_scan_bwd + deterministic reducers).

    Returns dict(du[list], ddelta[list], dz, dB[list], dC[list], dA[list], dD[list], dbias[list]).
    ``dz_out``: optional pre-allocated (B, D, L) view to receive dz (e.g. a slice of dxz).
    ``dBC_like``: optional list (per direction) of (dB, dC) tensors to fill instead of allocating.
    """
    lib = cabi.lib()
    Bt, D, L, N, const_bc = _check_dirs(dirs)
    u0 = dirs[0]["u"]
    dev = u0.device
    a = cabi.ScanBwdArgs()
    a.batch, a.dim, a.seqlen, a.dstate = Bt, D, L, N
    a.ndir = len(dirs)
    a.dtype = cabi.dtype_code(u0.dtype)
    a.flags = cabi.CM_FLAG_DELTA_SOFTPLUS if delta_softplus else 0
    a.out_scale = float(out_scale)
    a.lanes_per_channel = int(lanes) or int(os.environ.get("CM_SCAN_LANES", "0"))     # 0 = library default
    keep = []
    res = dict(du=[], ddelta=[], dB=[], dC=[], dA=[], dD=[], dbias=[], dz=None)
    parts = []
    for r, d in enumerate(dirs):
        bd = a.dir[r]
        _fill_scan_dir(bd.inp, d, keep, const_bc)
        ck = ckpts[r]
        bd.inp.ckpt = ck.data_ptr()
        bd.inp.ckpt_sb, bd.inp.ckpt_sd = ck.stride(0), ck.stride(1)
        du = empty_like_bdl(d["u"])
        ddelta = empty_like_bdl(d["delta"])
        bd.du, bd.ddelta = cabi.t3(du), cabi.t3(ddelta)
        res["du"].append(du)
        res["ddelta"].append(ddelta)
    _require_cuda(dout, "dout")
    if dout.shape != (Bt, D, L):
        raise ValueError("dout must be (B, D, L)")
    if dout.dtype != u0.dtype:
        dout = dout.to(u0.dtype)
    a.dout = cabi.t3(dout)
    if z is not None:
        if out_pre is None:
            raise ValueError("out_pre (pre-gate output saved by forward) is required when z is given")
        dz = dz_out if dz_out is not None else empty_like_bdl(z)
        a.z, a.out_pre, a.dz = cabi.t3(z), cabi.t3(out_pre), cabi.t3(dz)
        res["dz"] = dz
    # the slab width of the dB/dC partial tensor depends on the kernel the library will take for these tensors
    slab_ch = lib.cm_scan_bwd_slab_channels(C.byref(a))
    if slab_ch <= 0:
        cabi.check(slab_ch, "cm_scan_bwd_slab_channels")
    n_slab = (D + slab_ch - 1) // slab_ch
    for r, d in enumerate(dirs):
        bd = a.dir[r]
        if const_bc:
            bc_part = torch.empty((Bt, D, 32), dtype=torch.float32, device=dev)
        else:
            bc_part = torch.empty((Bt, n_slab, L, 32), dtype=torch.float32, device=dev)
        dA_part = torch.empty((Bt, D, 16), dtype=torch.float32, device=dev)
        dD_part = torch.empty((Bt, D), dtype=torch.float32, device=dev) if d.get("D") is not None else None
        db_part = torch.empty((Bt, D), dtype=torch.float32, device=dev) if d.get("delta_bias") is not None else None
        bd.dBC_part, bd.dA_part = bc_part.data_ptr(), dA_part.data_ptr()
        bd.dD_part, bd.dbias_part = cabi.ptr(dD_part), cabi.ptr(db_part)
        parts.append((bc_part, dA_part, dD_part, db_part))
    st = cabi.stream_ptr()
    _call("cm_scan_bwd", lib.cm_scan_bwd, C.byref(a), st, tag=(Bt, D, L, len(dirs)))

    # defer: dD / d(delta_bias) go straight to parameter gradients (see deferred_reductions); dA usually does not - the node
    # that made A = -exp(A_log) reads it - unless the caller says so (defer_dA: that node runs the queue before reading)
    jobs, jobs_now = [], []
    outs = []
    for r, d in enumerate(dirs):
        bc_part, dA_part, dD_part, db_part = parts[r]
        dA = torch.empty((D, 16), dtype=torch.float32, device=dev)
        (jobs if defer_dA else jobs_now).append((dA_part, dA))
        dD = db = dBC = None
        if dD_part is not None:
            dD = torch.empty((D,), dtype=torch.float32, device=dev)
            jobs.append((dD_part, dD))
        if db_part is not None:
            db = torch.empty((D,), dtype=torch.float32, device=dev)
            jobs.append((db_part, db))
        if const_bc:
            dBC = torch.empty((D, 32), dtype=torch.float32, device=dev)
            jobs_now.append((bc_part, dBC))                   # constant B / C gradients are cast (read) below
        outs.append((dA, dD, db, dBC))
    reduce_many(jobs, defer=defer)
    reduce_many(jobs_now)
    for r, d in enumerate(dirs):
        bc_part = parts[r][0]
        dA, dD, db, dBC = outs[r]
        res["dA"].append(dA[:, :N])
        res["dD"].append(dD)
        res["dbias"].append(db)
        if const_bc:
            res["dB"].append(dBC[:, :N].to(d["B"].dtype))
            res["dC"].append(dBC[:, 16:16 + N].to(d["C"].dtype))
        else:
            if dBC_like is not None:
                dB, dC = dBC_like[r]
            else:
                dB = torch.empty_strided(d["B"].shape, _dense_strides(d["B"]), dtype=d["B"].dtype, device=dev)
                dC = torch.empty_strided(d["C"].shape, _dense_strides(d["C"]), dtype=d["C"].dtype, device=dev)
            _call("cm_reduce_dbc", lib.cm_reduce_dbc, bc_part.data_ptr(), Bt, n_slab, L, N, a.dtype, cabi.t3(dB), cabi.t3(dC), st)
            res["dB"].append(dB)
            res["dC"].append(dC)
    return res


def _dense_strides(t):
    """Strides of a dense tensor with t's dimension order (views into wider buffers become compact)."""
    order = sorted(range(t.dim()), key=lambda i: (t.stride(i), -i), reverse=True)
    strides = [0] * t.dim()
    acc = 1
    for i in reversed(order):
        strides[i] = acc
        acc *= t.shape[i]
    return strides


# ------------------------------------------------------------------------------------------------------
def _conv_common(x, dirs, silu):
    lib = cabi.lib()
    _require_cuda(x, "x")
    Bt, D, L = x.shape
    W = dirs[0]["weight"].shape[1]
    a = cabi.ConvArgs()
    a.batch, a.dim, a.seqlen, a.width = Bt, D, L, W
    a.ndir = len(dirs)
    a.dtype = cabi.dtype_code(x.dtype)
    a.flags = cabi.CM_FLAG_SILU if silu else 0
    a.x = cabi.t3(x)
    keep = []
    for r, d in enumerate(dirs):
        w = d["weight"]
        if w.shape != (D, W):
            raise ValueError("conv weight must be (D, W)")
        _same_device(x, conv_weight=w, conv_bias=d.get("bias"))
        w = w.float().contiguous()
        b = d.get("bias")
        b = None if b is None else b.float().contiguous()
        keep += [w, b]
        cd = a.dir[r]
        cd.anticausal = 1 if d.get("anticausal", False) else 0
        cd.weight = w.data_ptr()
        cd.bias = cabi.ptr(b)
    return lib, a, keep, (Bt, D, L, W)


def conv_forward(x, dirs, silu=True, outs=None):
    """cm_conv_fwd.  x: (B, D, L) any strides; dirs: 1-2 dicts {weight (D, W), bias, anticausal}.
    Returns the list of outputs (same memory order as x)."""
    lib, a, keep, _ = _conv_common(x, dirs, silu)
    res = []
    for r in range(len(dirs)):
        o = outs[r] if outs is not None else empty_like_bdl(x)
        a.dir[r].out = cabi.t3(o)
        res.append(o)
    _call("cm_conv_fwd", lib.cm_conv_fwd, C.byref(a), cabi.stream_ptr(), tag=(a.batch, a.dim, a.seqlen, len(dirs)))
    return res


def conv_backward(x, dirs, douts, silu=True, dx_out=None, defer=False):
    """cm_conv_bwd.  Returns (dx, [dweight (D, W) fp32], [dbias (D,) fp32 or None])."""
    lib, a, keep, (Bt, D, L, W) = _conv_common(x, dirs, silu)
    dev = x.device
    dx = dx_out if dx_out is not None else empty_like_bdl(x)
    a.dx = cabi.t3(dx)
    npart = lib.cm_conv_num_part(Bt, L)
    parts = []
    for r, d in enumerate(dirs):
        g = douts[r]
        if g.shape != x.shape:
            raise ValueError("dout must match x")
        if g.dtype != x.dtype:
            g = g.to(x.dtype)
        keep.append(g)
        a.dir[r].out = cabi.t3(g)
        wp = torch.empty((npart, D, W), dtype=torch.float32, device=dev)
        bp = torch.empty((npart, D), dtype=torch.float32, device=dev) if d.get("bias") is not None else None
        a.dir[r].dweight_part = wp.data_ptr()
        a.dir[r].dbias_part = cabi.ptr(bp)
        parts.append((wp, bp))
    st = cabi.stream_ptr()
    _call("cm_conv_bwd", lib.cm_conv_bwd, C.byref(a), st, tag=(Bt, D, L, len(dirs)))
    dws, dbs, jobs = [], [], []
    for wp, bp in parts:
        dw = torch.empty((D, W), dtype=torch.float32, device=dev)
        jobs.append((wp, dw))
        dws.append(dw)
        if bp is not None:
            db = torch.empty((D,), dtype=torch.float32, device=dev)
            jobs.append((bp, db))
            dbs.append(db)
        else:
            dbs.append(None)
    reduce_many(jobs, defer=defer)
    return dx, dws, dbs


def conv_update(x, conv_state, weight, bias=None, silu=False):
    """cm_conv_update: x (B, D), conv_state (B, D, W) contiguous, updated in place; returns (B, D)."""
    lib = cabi.lib()
    _require_cuda(x, "x")
    Bt, D = x.shape
    W = weight.shape[1]
    if not conv_state.is_contiguous() or conv_state.shape != (Bt, D, W) or conv_state.dtype != x.dtype:
        raise ValueError("conv_state must be a contiguous (B, D, W) tensor of x's dtype")
    x = x.contiguous()
    w = weight.float().contiguous()
    b = None if bias is None else bias.float().contiguous()
    out = torch.empty_like(x)
    _call("cm_conv_update", lib.cm_conv_update, x.data_ptr(), conv_state.data_ptr(), w.data_ptr(), cabi.ptr(b), out.data_ptr(), Bt, D, W,
                                  cabi.dtype_code(x.dtype), cabi.CM_FLAG_SILU if silu else 0, cabi.stream_ptr())
    return out


# ------------------------------------------------------------------------------------------------------
def ssm_step(state, x, dt, A, Bm, Cm, D=None, z=None, dt_bias=None, dt_softplus=False):
    """cm_ssm_step: one decoding token.  state (B, D, N) contiguous, updated in place; x, dt, z (B, D); Bm, Cm (B, N);
    A (D, N); D, dt_bias (D,).  Returns out (B, D) = (<state, C> + D*x) * silu(z)."""
    lib = cabi.lib()
    _require_cuda(x, "x")
    _require_cuda(state, "state")
    Bt, Dm = x.shape
    N = A.shape[-1]
    if state.shape != (Bt, Dm, N) or not state.is_contiguous():
        raise ValueError("ssm_state must be a contiguous (B, D, N) tensor")
    if A.shape != (Dm, N) or Bm.shape != (Bt, N) or Cm.shape != (Bt, N) or dt.shape != (Bt, Dm):
        raise ValueError("ssm_step: shapes must be x, dt (B, D); A (D, N); B, C (B, N)")
    act = x.dtype

    def row(t, name):
        if t.dtype != act:
            t = t.to(act)
        if t.stride(-1) != 1:
            t = t.contiguous()
        return t

    x, dt, Bm, Cm = row(x, "x"), row(dt, "dt"), row(Bm, "B"), row(Cm, "C")
    z = None if z is None else row(z, "z")
    Af = _f32c(A.float(), "A")
    Df = None if D is None else _f32c(D.float(), "D")
    bf = None if dt_bias is None else _f32c(dt_bias.float(), "dt_bias")
    out = torch.empty((Bt, Dm), dtype=act, device=x.device)
    a = cabi.SsmStepArgs()
    a.batch, a.dim, a.dstate = Bt, Dm, N
    a.dtype, a.state_dtype = cabi.dtype_code(act), cabi.dtype_code(state.dtype)
    a.flags = cabi.CM_FLAG_DELTA_SOFTPLUS if dt_softplus else 0
    a.state, a.x, a.dt, a.z = state.data_ptr(), x.data_ptr(), dt.data_ptr(), cabi.ptr(z)
    a.Bm, a.Cm, a.out = Bm.data_ptr(), Cm.data_ptr(), out.data_ptr()
    a.x_sb, a.dt_sb, a.z_sb = x.stride(0), dt.stride(0), (0 if z is None else z.stride(0))
    a.b_sb, a.c_sb, a.out_sb = Bm.stride(0), Cm.stride(0), out.stride(0)
    a.A, a.Dskip, a.dt_bias = Af.data_ptr(), cabi.ptr(Df), cabi.ptr(bf)
    _call("cm_ssm_step", lib.cm_ssm_step, C.byref(a), cabi.stream_ptr())
    return out


def fbank_logmel(stft, fbank, top_db=80.0, amin=1e-10, multiplier=10.0, db_offset=0.0):
    """cm_fbank_logmel + cm_fbank_floor.  stft: complex64 (B, F, T) as returned by torch.stft; fbank: (F, M) fp32.
    Returns (B, T, M) fp32 log-mel features with the per-utterance top_db floor applied."""
    lib = cabi.lib()
    _require_cuda(stft, "stft")
    if stft.dtype != torch.complex64:
        raise TypeError("stft must be complex64")
    Bt, F, T = stft.shape
    M = fbank.shape[1]
    if fbank.shape[0] != F:
        raise ValueError("fbank must be (n_bins, n_mels)")
    _same_device(stft, fbank=fbank)
    fb = fbank.float().contiguous()
    out = torch.empty((Bt, T, M), dtype=torch.float32, device=stft.device)
    umax = torch.full((Bt,), float("-inf"), dtype=torch.float32, device=stft.device)
    a = cabi.FbankArgs()
    a.batch, a.frames, a.nbins, a.nmels = Bt, T, F, M
    a.stft = stft.data_ptr()
    a.s_b, a.s_f, a.s_t = stft.stride()
    a.fbank, a.out, a.utt_max = fb.data_ptr(), out.data_ptr(), umax.data_ptr()
    a.amin, a.multiplier, a.db_offset, a.top_db = amin, multiplier, db_offset, top_db
    st = cabi.stream_ptr()
    _call("cm_fbank_logmel", lib.cm_fbank_logmel, C.byref(a), st)
    _call("cm_fbank_floor", lib.cm_fbank_floor, C.byref(a), st)
    return out


def fbank_wav_supported(n_fft):
    return bool(cabi.lib().cm_fbank_wav_supported(int(n_fft)))


def fbank_wav_logmel(wav, window, fbank, band, n_fft, hop, top_db=80.0, amin=1e-10, multiplier=10.0, db_offset=0.0):
    """cm_fbank_wav_logmel + cm_fbank_floor: the whole Fbank front-end from the samples (windowed DFT in the kernel, no cuFFT,
    no STFT tensor).  wav: (B, n_samples) fp32 CUDA; window: (n_fft,) fp32 (analysis window zero-padded to n_fft);
    fbank: (n_fft // 2 + 1, M) fp32; band: (M, 2) int32 support of each filter.  Returns (B, 1 + n_samples // hop, M) fp32."""
    lib = cabi.lib()
    _require_cuda(wav, "wav")
    if wav.dim() != 2 or wav.dtype != torch.float32:
        raise TypeError("wav must be a (B, n_samples) float32 tensor")
    _same_device(wav, window=window, fbank=fbank, band=band)
    if wav.stride(1) != 1:
        wav = wav.contiguous()
    Bt, n = wav.shape
    M = fbank.shape[1]
    if fbank.shape[0] != n_fft // 2 + 1 or window.numel() != n_fft or tuple(band.shape) != (M, 2) or band.dtype != torch.int32:
        raise ValueError("window (n_fft,), fbank (n_fft // 2 + 1, M) and band (M, 2) int32 expected")
    T = 1 + n // hop
    fb, win, bd = fbank.float().contiguous(), window.float().contiguous(), band.contiguous()
    out = torch.empty((Bt, T, M), dtype=torch.float32, device=wav.device)
    umax = torch.full((Bt,), float("-inf"), dtype=torch.float32, device=wav.device)
    a = cabi.FbankWavArgs()
    a.batch, a.n_samples, a.frames, a.nmels = Bt, n, T, M
    a.n_fft, a.hop = int(n_fft), int(hop)
    a.wav, a.wav_sb = wav.data_ptr(), wav.stride(0)
    a.window, a.fbank, a.band, a.out, a.utt_max = win.data_ptr(), fb.data_ptr(), bd.data_ptr(), out.data_ptr(), umax.data_ptr()
    a.amin, a.multiplier, a.db_offset, a.top_db = amin, multiplier, db_offset, top_db
    f = cabi.FbankArgs()
    f.batch, f.frames, f.nbins, f.nmels = Bt, T, n_fft // 2 + 1, M
    f.out, f.utt_max, f.top_db = out.data_ptr(), umax.data_ptr(), top_db
    st = cabi.stream_ptr()
    _call("cm_fbank_wav_logmel", lib.cm_fbank_wav_logmel, C.byref(a), st)
    _call("cm_fbank_floor", lib.cm_fbank_floor, C.byref(f), st)
    return out


# ------------------------------------------------------------------------------------------------ CTC
def ctc_supported(max_target):
    """The kernel gives every extended-label state a thread in each of its two groups: up to 255 labels per utterance."""
    return ((2 * int(max_target) + 1 + 31) // 32 * 32) * 2 <= 1024


def ctc_nll_and_grad(log_probs, targets, input_lengths=None, target_lengths=None, blank=0, need_grad=True):
    """cm_ctc_loss: per-utterance negative log-likelihood and its gradient in one launch.

    log_probs: (B, T, C) fp32 CUDA, unit class stride; targets: (B, S) int64 (padded); lengths: (B,) int64 CUDA or None
    (= full).  Returns (nll (B,) fp32, grad (B, T, C) fp32 or None) with grad[b] = d nll[b] / d log_probs[b]."""
    lib = cabi.lib()
    _require_cuda(log_probs, "log_probs")
    if log_probs.dim() != 3 or log_probs.dtype != torch.float32:
        raise TypeError("log_probs must be a (B, T, C) float32 tensor")
    if log_probs.stride(2) != 1:
        log_probs = log_probs.contiguous()
    Bt, T, Cn = log_probs.shape
    if targets.dim() != 2 or targets.shape[0] != Bt or targets.dtype != torch.int64:
        raise TypeError("targets must be a (B, S) int64 tensor")
    _same_device(log_probs, targets=targets, input_lengths=input_lengths, target_lengths=target_lengths)
    if targets.stride(1) != 1:
        targets = targets.contiguous()
    S = targets.shape[1]
    for name, t in (("input_lengths", input_lengths), ("target_lengths", target_lengths)):
        if t is not None and (t.dtype != torch.int64 or t.shape != (Bt,) or not t.is_contiguous()):
            raise TypeError("%s must be a contiguous (B,) int64 tensor" % name)
    dev = log_probs.device
    nll = torch.empty((Bt,), dtype=torch.float32, device=dev)
    grad = torch.empty((Bt, T, Cn), dtype=torch.float32, device=dev) if need_grad else None
    ws = torch.empty((max(1, lib.cm_ctc_workspace_floats(Bt, T, S)),), dtype=torch.float32, device=dev)
    a = cabi.CtcArgs()
    a.batch, a.max_time, a.classes, a.max_target = Bt, T, Cn, S
    a.blank, a.ws_states = int(blank), 2 * S + 1
    a.log_probs, a.lp_sb, a.lp_st = log_probs.data_ptr(), log_probs.stride(0), log_probs.stride(1)
    a.targets, a.tg_sb = targets.data_ptr(), targets.stride(0)
    a.input_lengths, a.target_lengths = cabi.ptr(input_lengths), cabi.ptr(target_lengths)
    a.nll = nll.data_ptr()
    a.grad = cabi.ptr(grad)
    if grad is not None:
        a.g_sb, a.g_st = grad.stride(0), grad.stride(1)
    a.workspace = ws.data_ptr()
    _call("cm_ctc_loss", lib.cm_ctc_loss, C.byref(a), cabi.stream_ptr(), tag=(Bt, T, Cn, S))
    return nll, grad


# ------------------------------------------------------------------------------------------------ LayerNorm
def layernorm_gelu_supported(x2d):
    """True when the GELU epilogue of cm_layernorm_fwd / _bwd takes these rows (the pair-vectorised kernels: even column
    count and row stride, 8-byte aligned)."""
    return (x2d.is_cuda and x2d.dim() == 2 and x2d.stride(1) == 1 and x2d.shape[1] % 2 == 0 and x2d.shape[1] <= 1024
            and x2d.stride(0) % 2 == 0 and x2d.data_ptr() % 8 == 0)


def layernorm_forward(x2d, weight, bias, eps, out_dtype, gelu=False):
    """cm_layernorm_fwd over the rows of a (rows, C) CUDA tensor with unit column stride.
    Returns (y (rows, C) in out_dtype, mean (rows,) fp32, rstd (rows,) fp32).  gelu=True: y = gelu(LayerNorm(x))."""
    lib = cabi.lib()
    _require_cuda(x2d, "x")
    rows, Cn = x2d.shape
    if x2d.stride(1) != 1:
        raise ValueError("layernorm: the normalised dimension must be contiguous")
    _same_device(x2d, weight=weight, bias=bias)
    y = torch.empty((rows, Cn), dtype=out_dtype, device=x2d.device)
    mean = torch.empty((rows,), dtype=torch.float32, device=x2d.device)
    rstd = torch.empty((rows,), dtype=torch.float32, device=x2d.device)
    a = cabi.LayerNormArgs()
    a.rows, a.cols = rows, Cn
    a.x_dtype, a.y_dtype = cabi.dtype_code(x2d.dtype), cabi.dtype_code(out_dtype)
    a.eps = float(eps)
    a.x, a.x_stride = x2d.data_ptr(), x2d.stride(0)
    a.y, a.y_stride = y.data_ptr(), y.stride(0)
    a.gamma, a.beta = cabi.ptr(weight), cabi.ptr(bias)
    a.mean, a.rstd = mean.data_ptr(), rstd.data_ptr()
    a.act = cabi.CM_LN_OUT_GELU if gelu else 0
    _call("cm_layernorm_fwd", lib.cm_layernorm_fwd, C.byref(a), cabi.stream_ptr())
    return y, mean, rstd


def layernorm_backward(x2d, dy2d, weight, mean, rstd, need_wgrad=True, defer=False, gelu=False, bias=None):
    """cm_layernorm_bwd + deterministic reduction of the per-CTA dgamma / dbeta partial rows.
    Returns (dx in x's dtype, dgamma fp32 (C,), dbeta fp32 (C,)).  gelu=True: dy is the gradient of gelu(LayerNorm(x))
    (the forward's GELU epilogue; ``bias`` = the LayerNorm's beta, needed to recompute the pre-activation)."""
    lib = cabi.lib()
    rows, Cn = x2d.shape
    if dy2d.stride(1) != 1:
        dy2d = dy2d.contiguous()
    dx = torch.empty((rows, Cn), dtype=x2d.dtype, device=x2d.device)
    n_part = lib.cm_layernorm_num_part2(rows, Cn)
    dg_part = torch.empty((n_part, Cn), dtype=torch.float32, device=x2d.device)
    db_part = torch.empty((n_part, Cn), dtype=torch.float32, device=x2d.device)
    a = cabi.LayerNormArgs()
    a.rows, a.cols = rows, Cn
    a.x_dtype, a.y_dtype = cabi.dtype_code(x2d.dtype), cabi.dtype_code(dy2d.dtype)
    a.x, a.x_stride = x2d.data_ptr(), x2d.stride(0)
    a.gamma = cabi.ptr(weight)
    a.mean, a.rstd = mean.data_ptr(), rstd.data_ptr()
    a.dy, a.dy_stride = dy2d.data_ptr(), dy2d.stride(0)
    a.dx, a.dx_stride = dx.data_ptr(), dx.stride(0)
    a.dgamma_part, a.dbeta_part = dg_part.data_ptr(), db_part.data_ptr()
    a.n_part = n_part
    if gelu:
        a.act, a.beta = cabi.CM_LN_OUT_GELU, cabi.ptr(bias)
    _call("cm_layernorm_bwd", lib.cm_layernorm_bwd, C.byref(a), cabi.stream_ptr())
    if not need_wgrad:
        return dx, None, None
    dg = torch.empty((Cn,), dtype=torch.float32, device=x2d.device)
    db = torch.empty((Cn,), dtype=torch.float32, device=x2d.device)
    reduce_many([(dg_part, dg), (db_part, db)], defer=defer)
    return dx, dg, db


# ------------------------------------------------------------------------------------------------ wide LayerNorm + LeakyReLU
LN_ACT_MAX_COLS = 2560


def ln_act_supported(x2d):
    """cm_ln_act_* envelope: dense (rows, C) CUDA matrix, C a multiple of 4 and <= 2560, 16-byte aligned."""
    return (x2d.is_cuda and x2d.dim() == 2 and x2d.is_contiguous() and x2d.dtype in cabi._DTYPES
            and x2d.shape[1] % 4 == 0 and 0 < x2d.shape[1] <= LN_ACT_MAX_COLS and x2d.shape[0] > 0
            and x2d.data_ptr() % 16 == 0)


LN_ACTS = {"leaky_relu": cabi.CM_LN_ACT_LEAKY_RELU, "gelu": cabi.CM_LN_ACT_GELU}


def _ln_act_args(x2d, weight, bias, eps, slope, mean, rstd, act, pre_bias):
    _same_device(x2d, weight=weight, bias=bias, pre_bias=pre_bias)
    a = cabi.LnActArgs()
    a.rows, a.cols = x2d.shape
    a.dtype = cabi.dtype_code(x2d.dtype)
    a.eps, a.slope = float(eps), float(slope)
    a.act = LN_ACTS[act]
    if pre_bias is not None:
        pre_bias = _f32c(pre_bias, "pre_bias")
        n = pre_bias.numel()
        if n % 4 != 0 or x2d.shape[1] % n != 0 or pre_bias.data_ptr() % 16 != 0:
            raise NotImplementedError("ln_act: pre_bias needs a length that is a multiple of 4 and divides the row length")
        a.pre_bias, a.pre_bias_n = pre_bias.data_ptr(), n
    a.x = x2d.data_ptr()
    a.gamma, a.beta = _f32c(weight, "weight").data_ptr(), _f32c(bias, "bias").data_ptr()
    a.mean, a.rstd = mean.data_ptr(), rstd.data_ptr()
    return a


def ln_act_forward(x2d, weight, bias, eps, slope=0.01, act="leaky_relu", pre_bias=None):
    """cm_ln_act_fwd: y = act(layer_norm(x + pre_bias)) over the rows of a dense (rows, C) matrix, y in x's dtype;
    act = "leaky_relu" (negative slope `slope`) or "gelu" (exact); pre_bias (n,) fp32 repeats every n columns.
    Returns (y, mean (rows,) fp32, rstd (rows,) fp32)."""
    lib = cabi.lib()
    _require_cuda(x2d, "x")
    rows = x2d.shape[0]
    y = torch.empty_like(x2d)
    mean = torch.empty((rows,), dtype=torch.float32, device=x2d.device)
    rstd = torch.empty((rows,), dtype=torch.float32, device=x2d.device)
    a = _ln_act_args(x2d, weight, bias, eps, slope, mean, rstd, act, pre_bias)
    a.y = y.data_ptr()
    _call("cm_ln_act_fwd", lib.cm_ln_act_fwd, C.byref(a), cabi.stream_ptr())
    return y, mean, rstd


def ln_act_backward(x2d, dy2d, weight, bias, mean, rstd, slope=0.01, act="leaky_relu", pre_bias=None, need_wgrad=True,
                    defer=False):
    """cm_ln_act_bwd + deterministic reduction of the per-CTA dgamma / dbeta partial rows.
    Returns (dx in x's dtype (= the gradient of x + pre_bias), dgamma fp32 (C,), dbeta fp32 (C,))."""
    lib = cabi.lib()
    rows, Cn = x2d.shape
    if dy2d.dtype != x2d.dtype or not dy2d.is_contiguous():
        dy2d = dy2d.to(x2d.dtype).contiguous()
    dx = torch.empty_like(x2d)
    n_part = lib.cm_ln_act_num_part(rows, Cn)
    dg_part = torch.empty((n_part, Cn), dtype=torch.float32, device=x2d.device)
    db_part = torch.empty((n_part, Cn), dtype=torch.float32, device=x2d.device)
    a = _ln_act_args(x2d, weight, bias, 0.0, slope, mean, rstd, act, pre_bias)
    a.dy, a.dx = dy2d.data_ptr(), dx.data_ptr()
    a.dgamma_part, a.dbeta_part = dg_part.data_ptr(), db_part.data_ptr()
    _call("cm_ln_act_bwd", lib.cm_ln_act_bwd, C.byref(a), cabi.stream_ptr())
    if not need_wgrad:
        return dx, None, None
    dg = torch.empty((Cn,), dtype=torch.float32, device=x2d.device)
    db = torch.empty((Cn,), dtype=torch.float32, device=x2d.device)
    reduce_many([(dg_part, dg), (db_part, db)], defer=defer)
    return dx, dg, db


# ------------------------------------------------------------------------------------------------ front-end block 1
def stem_supported(feats, conv_weight):
    """True when cm_stem_* implements Conv2d(1 -> C, 3 x 3, stride 2, padding 1) + LayerNorm([F', C]) + LeakyReLU for this input."""
    if not (feats.is_cuda and feats.dim() == 3 and feats.dtype in cabi._DTYPES):
        return False
    if conv_weight.dim() != 4 or tuple(conv_weight.shape[1:]) != (1, 3, 3):
        return False
    return bool(cabi.lib().cm_stem_supported(feats.shape[2], conv_weight.shape[0]))


def _stem_args(feats, weight, bias, gamma, beta, eps, slope, out_dtype, mean, rstd):
    _require_cuda(feats, "feats")
    _same_device(feats, weight=weight, bias=bias, gamma=gamma, beta=beta)
    a = cabi.StemArgs()
    a.batch, a.frames, a.feats = feats.shape
    a.channels = weight.shape[0]
    a.in_dtype, a.out_dtype = cabi.dtype_code(feats.dtype), cabi.dtype_code(out_dtype)
    a.eps, a.slope = float(eps), float(slope)
    a.inp = feats.data_ptr()
    a.weight = _f32c(weight, "weight").data_ptr()
    a.bias = cabi.ptr(None if bias is None else _f32c(bias, "bias"))
    a.gamma, a.beta = _f32c(gamma, "gamma").data_ptr(), _f32c(beta, "beta").data_ptr()
    a.mean, a.rstd = mean.data_ptr(), rstd.data_ptr()
    return a


def stem_forward(feats, weight, bias, gamma, beta, eps, slope, out_dtype):
    """cm_stem_fwd: feats (B, T, F) dense -> y (B, T', F', C) in out_dtype, mean / rstd (B * T') fp32.
    weight (C, 1, 3, 3) / bias (C) / gamma, beta (F', C): fp32 contiguous."""
    lib = cabi.lib()
    if not feats.is_contiguous():
        feats = feats.contiguous()
    Bt, T, Fd = feats.shape
    Cn = weight.shape[0]
    To, Fo = (T - 1) // 2 + 1, (Fd - 1) // 2 + 1
    y = torch.empty((Bt, To, Fo, Cn), dtype=out_dtype, device=feats.device)
    mean = torch.empty((Bt * To,), dtype=torch.float32, device=feats.device)
    rstd = torch.empty((Bt * To,), dtype=torch.float32, device=feats.device)
    a = _stem_args(feats, weight, bias, gamma, beta, eps, slope, out_dtype, mean, rstd)
    a.y = y.data_ptr()
    _call("cm_stem_fwd", lib.cm_stem_fwd, C.byref(a), cabi.stream_ptr(), tag=(Bt, T, Fd, Cn))
    return y, mean, rstd


def stem_backward(feats, dy, weight, bias, gamma, beta, mean, rstd, slope, defer=False):
    """cm_stem_bwd + one deterministic reduction launch.  Returns (dweight (C, 1, 3, 3), dbias (C), dgamma, dbeta (F', C)) fp32."""
    lib = cabi.lib()
    Bt, T, Fd = feats.shape
    Cn = weight.shape[0]
    Fo = (Fd - 1) // 2 + 1
    if not dy.is_contiguous():
        dy = dy.contiguous()
    dev = feats.device
    n_part = lib.cm_stem_num_part(Bt, T)
    dg_part = torch.empty((n_part, Fo * Cn), dtype=torch.float32, device=dev)
    db_part = torch.empty((n_part, Fo * Cn), dtype=torch.float32, device=dev)
    dw_part = torch.empty((n_part, Cn * 9), dtype=torch.float32, device=dev)
    dcb_part = torch.empty((n_part, Cn), dtype=torch.float32, device=dev)
    a = _stem_args(feats, weight, bias, gamma, beta, 0.0, slope, dy.dtype, mean, rstd)
    a.dy = dy.data_ptr()
    a.dgamma_part, a.dbeta_part = dg_part.data_ptr(), db_part.data_ptr()
    a.dweight_part, a.dbias_part = dw_part.data_ptr(), dcb_part.data_ptr()
    _call("cm_stem_bwd", lib.cm_stem_bwd, C.byref(a), cabi.stream_ptr(), tag=(Bt, T, Fd, Cn))
    dg = torch.empty((Fo, Cn), dtype=torch.float32, device=dev)
    db = torch.empty((Fo, Cn), dtype=torch.float32, device=dev)
    dw = torch.empty((Cn, 1, 3, 3), dtype=torch.float32, device=dev)
    dcb = torch.empty((Cn,), dtype=torch.float32, device=dev)
    reduce_many([(dg_part, dg), (db_part, db), (dw_part, dw), (dcb_part, dcb)], defer=defer)
    return dw, dcb, dg, db


# ------------------------------------------------------------------------------------------------ add + dropout + LayerNorm
ADD_LN_COMBOS = {(torch.float32, torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16, torch.float32),
                 (torch.float32, torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16, torch.bfloat16),
                 (torch.bfloat16, torch.bfloat16, torch.float32)}


def add_ln_supported(a2d, b2d, out_dtype):
    """True when cm_add_ln_* implements this (dtype, shape, alignment) combination."""
    rows, Cn = a2d.shape
    bd = a2d.dtype if b2d is None else b2d.dtype
    if b2d is None and a2d.dtype == torch.float32 and out_dtype == torch.bfloat16:
        bd = torch.bfloat16
    if (a2d.dtype, bd, out_dtype) not in ADD_LN_COMBOS or Cn % 2 or Cn > 1024:
        return False
    for t in (a2d, b2d):
        if t is not None and (t.stride(1) != 1 or t.stride(0) % 2 or t.data_ptr() % 8):
            return False
    return True


def add_ln_forward(a2d, b2d, weight, bias, eps, alpha, p_drop, seed, call_id, out_dtype, need_s=True, store_mask=None):
    """cm_add_ln_fwd: s = a + alpha * dropout_p(b); y = LayerNorm(s).  a2d (rows, C); b2d (rows, C) or None; seed: int64
    CUDA scalar tensor or None.  Returns (s, y, mean, rstd, saved): ``saved`` is what backward needs to rebuild the dropout
    mask - None (no dropout), a (1,) int32 key tensor (default: regenerated, nothing stored) or the uint8 byte mask
    (``store_mask=True`` / CM_DROPOUT_STORE_MASK=1)."""
    lib = cabi.lib()
    _require_cuda(a2d, "a")
    rows, Cn = a2d.shape
    dev = a2d.device
    s = torch.empty((rows, Cn), dtype=a2d.dtype, device=dev) if (need_s or b2d is not None) else None
    y = torch.empty((rows, Cn), dtype=out_dtype, device=dev)
    mean = torch.empty((rows,), dtype=torch.float32, device=dev)
    rstd = torch.empty((rows,), dtype=torch.float32, device=dev)
    drop = p_drop > 0.0 and b2d is not None
    if store_mask is None:
        store_mask = os.environ.get("CM_DROPOUT_STORE_MASK", "0") == "1"
    saved = None
    if drop:
        saved = (torch.empty((rows, Cn), dtype=torch.uint8, device=dev) if store_mask
                 else torch.empty((1,), dtype=torch.int32, device=dev))
    _same_device(a2d, b=b2d, weight=weight, bias=bias, seed=seed)
    a = cabi.AddLnArgs()
    a.rows, a.cols = rows, Cn
    a.a_dtype, a.y_dtype = cabi.dtype_code(a2d.dtype), cabi.dtype_code(out_dtype)
    a.b_dtype = cabi.dtype_code(b2d.dtype) if b2d is not None else a.a_dtype
    a.eps, a.alpha, a.p_drop = float(eps), float(alpha), float(p_drop if drop else 0.0)
    a.call_id = int(call_id) & 0xffffffff
    a.seed = cabi.ptr(seed)
    a.a, a.a_stride = a2d.data_ptr(), a2d.stride(0)
    if b2d is not None:
        a.b, a.b_stride = b2d.data_ptr(), b2d.stride(0)
    if s is not None:
        a.s, a.s_stride = s.data_ptr(), s.stride(0)
    a.y, a.y_stride = y.data_ptr(), y.stride(0)
    if saved is not None:
        if saved.dtype == torch.uint8:
            a.mask = saved.data_ptr()
        else:
            a.key = saved.data_ptr()
    a.gamma, a.beta = cabi.ptr(weight), cabi.ptr(bias)
    a.mean, a.rstd = mean.data_ptr(), rstd.data_ptr()
    _call("cm_add_ln_fwd", lib.cm_add_ln_fwd, C.byref(a), cabi.stream_ptr())
    return s, y, mean, rstd, saved


def add_ln_backward(s2d, dy2d, ds2d, weight, mean, rstd, saved, alpha, p_drop, b_dtype, need_db=True, need_wgrad=True,
                    need_dbsum=False, defer=False):
    """cm_add_ln_bwd + deterministic reduction of the dgamma / dbeta partial rows.  ``saved`` as returned by
    ``add_ln_forward``.  Returns (da in s's dtype, db in b_dtype or None, dgamma fp32 (C,), dbeta fp32 (C,)); with
    ``need_dbsum`` a fifth value: the fp32 column sums of db (None when the geometry is outside the quad kernels)."""
    lib = cabi.lib()
    rows, Cn = s2d.shape
    dev = s2d.device
    if dy2d.stride(1) != 1 or dy2d.stride(0) % 2:
        dy2d = dy2d.contiguous()
    if ds2d is not None and (ds2d.stride(1) != 1 or ds2d.stride(0) % 2 or ds2d.dtype != s2d.dtype):
        ds2d = ds2d.to(s2d.dtype).contiguous()
    da = torch.empty((rows, Cn), dtype=s2d.dtype, device=dev)
    db = torch.empty((rows, Cn), dtype=b_dtype, device=dev) if need_db else None
    n_part = lib.cm_add_ln_num_part(rows, Cn)
    dg_part = torch.empty((n_part, Cn), dtype=torch.float32, device=dev)
    db_part = torch.empty((n_part, Cn), dtype=torch.float32, device=dev)
    a = cabi.AddLnArgs()
    a.rows, a.cols = rows, Cn
    a.a_dtype, a.b_dtype, a.y_dtype = cabi.dtype_code(s2d.dtype), cabi.dtype_code(b_dtype), cabi.dtype_code(dy2d.dtype)
    a.alpha, a.p_drop = float(alpha), float(p_drop if saved is not None else 0.0)
    a.s, a.s_stride = s2d.data_ptr(), s2d.stride(0)
    if saved is not None:
        if saved.dtype == torch.uint8:
            a.mask = saved.data_ptr()
        else:
            a.key = saved.data_ptr()
    a.gamma = cabi.ptr(weight)
    a.mean, a.rstd = mean.data_ptr(), rstd.data_ptr()
    a.dy, a.dy_stride = dy2d.data_ptr(), dy2d.stride(0)
    if ds2d is not None:
        a.ds, a.ds_stride = ds2d.data_ptr(), ds2d.stride(0)
    a.da, a.da_stride = da.data_ptr(), da.stride(0)
    if db is not None:
        a.db, a.db_stride = db.data_ptr(), db.stride(0)
    a.dgamma_part, a.dbeta_part = dg_part.data_ptr(), db_part.data_ptr()
    dbs_part = None
    if need_dbsum and db is not None:
        strides = [s2d.stride(0), dy2d.stride(0), da.stride(0), db.stride(0)] + ([ds2d.stride(0)] if ds2d is not None else [])
        ptrs_ok = all(t is None or t.data_ptr() % 16 == 0 for t in (s2d, dy2d, ds2d, da, db, weight))
        if ptrs_ok and lib.cm_add_ln_dbsum_supported(Cn, functools.reduce(lambda u, v: u | v, strides) & 3):
            dbs_part = torch.empty((n_part, Cn), dtype=torch.float32, device=dev)
            a.dbsum_part = dbs_part.data_ptr()
    _call("cm_add_ln_bwd", lib.cm_add_ln_bwd, C.byref(a), cabi.stream_ptr())
    jobs, dg, dbt, dbs = [], None, None, None
    if need_wgrad:
        dg = torch.empty((Cn,), dtype=torch.float32, device=dev)
        dbt = torch.empty((Cn,), dtype=torch.float32, device=dev)
        jobs += [(dg_part, dg), (db_part, dbt)]
    if dbs_part is not None:
        dbs = torch.empty((Cn,), dtype=torch.float32, device=dev)
        jobs.append((dbs_part, dbs))
    if jobs:
        reduce_many(jobs, defer=defer)
    if need_dbsum:
        return da, db, dg, dbt, dbs
    return da, db, dg, dbt


# ------------------------------------------------------------------------------------------------ GELU + dropout
def gelu_dropout_supported(x):
    return x.is_cuda and x.is_contiguous() and x.numel() % 8 == 0 and x.data_ptr() % 16 == 0 and \
        x.dtype in (torch.float32, torch.bfloat16, torch.float16)


def gelu_dropout_forward(x, p_drop, seed, call_id, store_mask=None):
    """cm_gelu_dropout_fwd_v2 on a contiguous tensor: returns (y, saved) where ``saved`` is what backward needs to rebuild the
    dropout mask - None (no dropout); by default the keep BITS, a uint8 tensor of numel / 8 bytes (one bit per element: the
    backward kernel is issue-bound and the four hashes per eight elements of a regenerated mask were a quarter of its
    instructions); a (1,) int32 key tensor with CM_DROPOUT_REGEN=1 (nothing stored, the mask is re-hashed); or the uint8
    byte mask (``store_mask=True`` or CM_DROPOUT_STORE_MASK=1)."""
    lib = cabi.lib()
    _require_cuda(x, "x")
    y = torch.empty_like(x)
    if store_mask is None:
        store_mask = os.environ.get("CM_DROPOUT_STORE_MASK", "0") == "1"
    a = cabi.ActArgs()
    a.x, a.y, a.n, a.dtype = x.data_ptr(), y.data_ptr(), x.numel(), cabi.dtype_code(x.dtype)
    a.p_drop, a.seed, a.call_id = float(p_drop), cabi.ptr(seed), int(call_id) & 0xffffffff
    saved = None
    if p_drop > 0.0:
        if store_mask:
            saved = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
            a.mask = saved.data_ptr()
        elif os.environ.get("CM_DROPOUT_REGEN", "0") == "1":
            saved = torch.empty((1,), dtype=torch.int32, device=x.device)
            a.key = saved.data_ptr()
        else:
            saved = torch.empty((x.numel() // 8,), dtype=torch.uint8, device=x.device)
            a.keep_bits = saved.data_ptr()
    _call("cm_gelu_dropout_fwd", lib.cm_gelu_dropout_fwd_v2, C.byref(a), cabi.stream_ptr())
    return y, saved


def gelu_dropout_backward(x, dy, saved, p_drop, colsum_cols=0, defer=False):
    """cm_gelu_dropout_bwd_v2.  ``saved`` as returned by ``gelu_dropout_forward``.  colsum_cols > 0: also returns the fp32
    column sums of dx viewed as (-1, colsum_cols) (None if that width is outside the fused envelope)."""
    lib = cabi.lib()
    if not dy.is_contiguous() or dy.dtype != x.dtype:
        dy = dy.to(x.dtype).contiguous()
    dx = torch.empty_like(x)
    a = cabi.ActArgs()
    a.x, a.dy, a.dx, a.n, a.dtype = x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel(), cabi.dtype_code(x.dtype)
    a.p_drop = float(p_drop) if saved is not None else 0.0
    if saved is not None:
        if saved.dtype == torch.uint8 and saved.numel() == x.numel():
            a.mask = saved.data_ptr()
        elif saved.dtype == torch.uint8:
            a.keep_bits = saved.data_ptr()
        else:
            a.key = saved.data_ptr()
    part = None
    if colsum_cols > 0 and lib.cm_act_colsum_supported(x.numel(), colsum_cols):
        part = torch.empty((lib.cm_act_num_part(x.numel()), colsum_cols), dtype=torch.float32, device=x.device)
        a.cols, a.colsum_part = colsum_cols, part.data_ptr()
    _call("cm_gelu_dropout_bwd", lib.cm_gelu_dropout_bwd_v2, C.byref(a), cabi.stream_ptr())
    if colsum_cols <= 0:
        return dx
    if part is None:
        return dx, None
    cs = torch.empty((colsum_cols,), dtype=torch.float32, device=x.device)
    reduce_many([(part, cs)], defer=defer)
    return dx, cs


# ------------------------------------------------------------------------------------------------ GLU (last dimension)
def glu_supported(h):
    """True when cm_glu_fwd / cm_glu_bwd take this tensor: CUDA, unit last stride, rows of 2C elements with C % 8 == 0 that
    collapse to one row stride, 16-byte aligned."""
    if not (h.is_cuda and h.dim() >= 2 and h.dtype in (torch.float32, torch.bfloat16, torch.float16)):
        return False
    Cn2 = h.shape[-1]
    return Cn2 % 16 == 0 and h.is_contiguous() and h.data_ptr() % 16 == 0 and h.numel() > 0


def glu_forward(h):
    """y = h[..., :C] * sigmoid(h[..., C:]) (cm_glu_fwd) for a contiguous tensor; see glu_supported."""
    lib = cabi.lib()
    Cn = h.shape[-1] // 2
    rows = h.numel() // (2 * Cn)
    y = torch.empty(h.shape[:-1] + (Cn,), dtype=h.dtype, device=h.device)
    _call("cm_glu_fwd", lib.cm_glu_fwd, h.data_ptr(), y.data_ptr(), rows, Cn, 2 * Cn, Cn, cabi.dtype_code(h.dtype), cabi.stream_ptr())
    return y


def glu_backward(h, dy):
    """Gradient of glu_forward with respect to h (cm_glu_bwd); dy contiguous, h's dtype."""
    lib = cabi.lib()
    Cn = h.shape[-1] // 2
    rows = h.numel() // (2 * Cn)
    _same_device(h, dy=dy)
    if dy.dtype != h.dtype:
        dy = dy.to(h.dtype)
    if not dy.is_contiguous() or dy.data_ptr() % 16:
        dy = dy.contiguous()
    dh = torch.empty_like(h)
    _call("cm_glu_bwd", lib.cm_glu_bwd, h.data_ptr(), dy.data_ptr(), dh.data_ptr(), rows, Cn, 2 * Cn, Cn, 2 * Cn,
          cabi.dtype_code(h.dtype), cabi.stream_ptr())
    return dh


# ------------------------------------------------------------------------------------------------ tall-skinny A^T B
def tsmm_supported(a2d, b2d):
    """cm_tsmm computes a2d^T @ b2d for 16-bit row-major operands with b2d at most 64 columns wide."""
    if not (a2d.is_cuda and a2d.dtype in (torch.bfloat16, torch.float16) and b2d.dtype == a2d.dtype):
        return False
    if a2d.dim() != 2 or b2d.dim() != 2 or a2d.shape[0] != b2d.shape[0] or a2d.shape[0] == 0:
        return False
    M, N = a2d.shape[1], b2d.shape[1]
    if N > 64 or N % 8 or M % 8 or a2d.stride(1) != 1 or b2d.stride(1) != 1:
        return False
    return a2d.stride(0) % 8 == 0 and b2d.stride(0) % 8 == 0 and a2d.data_ptr() % 16 == 0 and b2d.data_ptr() % 16 == 0


def tsmm(a2d, b2d, defer=False):
    """a2d^T @ b2d -> (M, N) fp32 through cm_tsmm (tensor-core partial blocks per 256-row chunk) and the deterministic
    reducer.  a2d (rows, M), b2d (rows, N <= 64), 16-bit, unit column stride."""
    lib = cabi.lib()
    rows, M = a2d.shape
    N = b2d.shape[1]
    n_part = lib.cm_tsmm_num_part(rows, M)
    part = torch.empty((n_part, M * N), dtype=torch.float32, device=a2d.device)
    _call("cm_tsmm", lib.cm_tsmm, a2d.data_ptr(), a2d.stride(0), b2d.data_ptr(), b2d.stride(0), part.data_ptr(), rows, M, N,
          cabi.dtype_code(a2d.dtype), cabi.stream_ptr())
    if n_part == 1:
        return part.view(M, N)
    out = torch.empty((M * N,), dtype=torch.float32, device=a2d.device)
    reduce_many([(part, out)], defer=defer)
    return out.view(M, N)


# ------------------------------------------------------------------------------------------------ depthwise conv1d
DWCONV_KSIZES = (3, 7, 15, 31)


def _dwconv_args(x_blc, ksize, pad_left):
    _require_cuda(x_blc, "x")
    if x_blc.dim() != 3:
        raise ValueError("dwconv: x must be (batch, time, channels)")
    if ksize not in DWCONV_KSIZES:
        raise NotImplementedError("dwconv: kernel_size must be one of %s" % (DWCONV_KSIZES,))
    Bt, L, Cn = x_blc.shape
    a = cabi.DwConvArgs()
    a.batch, a.dim, a.seqlen, a.ksize = Bt, Cn, L, ksize
    a.pad_left = int(pad_left)
    a.dtype = cabi.dtype_code(x_blc.dtype)
    return a


def dwconv_forward(x_blc, weight_ck, bias, pad_left, flip=False):
    """cm_dwconv_fwd on a channel-last (B, L, C) tensor; weight (C, K) fp32, bias (C,) fp32 or None.
    flip=True evaluates the backward-data form (taps reversed; pass pad_left' = K-1-pad_left, bias None)."""
    lib = cabi.lib()
    Cn, K = weight_ck.shape
    _same_device(x_blc, weight=weight_ck, bias=bias)
    a = _dwconv_args(x_blc, K, pad_left)
    y = torch.empty(x_blc.shape, dtype=x_blc.dtype, device=x_blc.device)
    a.flip = 1 if flip else 0
    a.x, a.y = cabi.t3(x_blc, "bld"), cabi.t3(y, "bld")
    a.weight, a.bias = weight_ck.data_ptr(), cabi.ptr(bias)
    _call("cm_dwconv_fwd", lib.cm_dwconv_fwd, C.byref(a), cabi.stream_ptr())
    return y


def dwconv_backward_weight(x_blc, dy_blc, ksize, pad_left, need_bias=True, defer=False):
    """cm_dwconv_bwd_weight + deterministic reduction: returns (dweight (C, K) fp32, dbias (C,) fp32 or None)."""
    lib = cabi.lib()
    Bt, L, Cn = x_blc.shape
    a = _dwconv_args(x_blc, ksize, pad_left)
    n_part = lib.cm_dwconv_num_part(Bt, L, ksize)
    dw_part = torch.empty((n_part, Cn * ksize), dtype=torch.float32, device=x_blc.device)
    db_part = torch.empty((n_part, Cn), dtype=torch.float32, device=x_blc.device) if need_bias else None
    a.x, a.dy = cabi.t3(x_blc, "bld"), cabi.t3(dy_blc, "bld")
    a.dweight_part, a.dbias_part = dw_part.data_ptr(), cabi.ptr(db_part)
    _call("cm_dwconv_bwd_weight", lib.cm_dwconv_bwd_weight, C.byref(a), cabi.stream_ptr())
    dw = torch.empty((Cn * ksize,), dtype=torch.float32, device=x_blc.device)
    jobs = [(dw_part, dw)]
    db = None
    if need_bias:
        db = torch.empty((Cn,), dtype=torch.float32, device=x_blc.device)
        jobs.append((db_part, db))
    reduce_many(jobs, defer=defer)
    return dw.view(Cn, ksize), db


# ------------------------------------------------------------------------------------------------ column sums (bias grads)
def colsum(x2d, defer=False):
    """fp32 column sums of a (rows, cols) CUDA matrix with unit column stride (cm_colsum + fixed-order reduction).
    Returns None when the shape is outside the kernel's envelope (odd cols / stride): the caller uses torch.sum."""
    lib = cabi.lib()
    _require_cuda(x2d, "x")
    rows, cols = x2d.shape
    if x2d.stride(1) != 1 or (cols & 1) or (x2d.stride(0) & 1) or x2d.data_ptr() % (2 * x2d.element_size()):
        return None
    n_part = lib.cm_colsum_num_part(rows)
    part = torch.empty((n_part, cols), dtype=torch.float32, device=x2d.device)
    _call("cm_colsum", lib.cm_colsum, x2d.data_ptr(), rows, cols, x2d.stride(0), cabi.dtype_code(x2d.dtype), part.data_ptr(),
          cabi.stream_ptr())
    out = torch.empty((cols,), dtype=torch.float32, device=x2d.device)
    reduce_many([(part, out)], defer=defer)
    return out


# ------------------------------------------------------------------------------------------------ optimizer step
def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, max_grad_norm=0.0, grad_scale=1.0, p_bf16=None,
               scratch=None, norm_out=None):
    """cm_sumsq_partial + cm_adamw_step on flat fp32 CUDA buffers (len a multiple of 4): clip to ``max_grad_norm`` (0 = off),
    AdamW with decoupled weight decay, ``step`` counted from 1.  ``grad_scale`` = 1 / world_size folds the DDP average in.
    ``scratch``: optional fp32 buffer of cm_optim_num_part(n) elements (allocated when None)."""
    lib = cabi.lib()
    _require_cuda(p, "p")
    n = p.numel()
    for t, name in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n:
            raise ValueError("adamw_step: %s must be a contiguous fp32 buffer of the parameters' length" % name)
    _same_device(p, g=g, m=m, v=v, p_bf16=p_bf16, scratch=scratch, norm_out=norm_out)
    a = cabi.AdamWArgs()
    a.p, a.g, a.m, a.v = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
    a.p_bf16 = cabi.ptr(p_bf16)
    a.norm_out = cabi.ptr(norm_out)
    a.n = n
    a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = float(lr), float(beta1), float(beta2), float(eps), float(weight_decay)
    a.bias_corr1, a.bias_corr2 = 1.0 - float(beta1) ** int(step), 1.0 - float(beta2) ** int(step)
    a.max_grad_norm, a.grad_scale = float(max_grad_norm), float(grad_scale)
    st = cabi.stream_ptr()
    if max_grad_norm and max_grad_norm > 0.0:
        n_part = lib.cm_optim_num_part(n)
        if scratch is None:
            scratch = torch.empty((n_part,), dtype=torch.float32, device=p.device)
        _call("cm_sumsq_partial", lib.cm_sumsq_partial, g.data_ptr(), n, scratch.data_ptr(), st)
        a.sumsq_part, a.n_part = scratch.data_ptr(), n_part
    _call("cm_adamw_step", lib.cm_adamw_step, C.byref(a), st)
