"""Channel-last Mamba inner block (conv -> x_proj -> dt_proj -> fused scan) as one autograd Function.

This is the B200-native body behind ``Mamba.forward`` (both BiMamba-v2 directions in one pass) and behind the
reference-signature wrappers in ``selective_scan_interface.py``.  It computes what the reference computes in
``MambaInnerFnNoOutProj.forward/backward`` (modules/mamba/selective_scan_interface.py:160-294) called twice with
a flip in between (modules/mamba/bimamba.py:223-253), but

  * activations stay time-major / channel-last ``(B, L, D)``: in_proj output, x_proj input and out_proj input
    need no transposes (the reference makes three transposing copies, :186-212), and every scan access is a
    coalesced row segment;
  * the two directions share one conv launch (one read of x), one scan launch (no ``flip``, one gate, one
    output) and one launch each in backward;
  * conv1d_out and delta are kept for backward instead of recomputed (the reference's checkpoint_lvl=1 saves
    memory that a 180 GB part does not need to save).

The dense projections are torch.matmul (cuBLAS tensor cores); everything else is the hand-written kernels.
"""
import os

import torch

from . import kernels as K
from .linear import cached_param


def _aligned_proj_weights(x_proj_w, dt_proj_w, R, N, act):
    """x_proj rows reordered to [B (N) | C (N) | dt (R) | zero pad] and dt_proj's K padded with zero columns so that
    both GEMMs have 8-element-aligned leading dimensions (tensor-core kernels instead of the align1 fallbacks at
    R = 9) and every B/C row of x_dbl starts on a 16-byte boundary.  The pad columns of x_dbl are exact zeros."""
    Rp = (R + 7) // 8 * 8
    xw = cached_param(x_proj_w, act)              # per-step bf16 copies (linear.ParamCache) when installed
    dw = cached_param(dt_proj_w, act)
    xw = x_proj_w.to(act) if xw is None else xw
    dw = dt_proj_w.to(act) if dw is None else dw
    parts = [xw[R:R + 2 * N], xw[:R]]
    if Rp != R:
        parts.append(xw.new_zeros((Rp - R, xw.shape[1])))
        dw = torch.cat([dw, dw.new_zeros((dw.shape[0], Rp - R))], dim=1)
    return torch.cat(parts, dim=0), dw


def _as_bdl(t_bld):
    """(B, L, D) memory -> logical (B, D, L) view."""
    return t_bld.transpose(1, 2)


_USE_TSMM = True      # A/B switch: CM_NO_TSMM=1 (or this flag) restores the cuBLAS bmm split
_TSMM_MAX_ELEMS = 1 << 22


def _wgrad(a, b, nsplit, defer=False, row_perm=None):
    """a^T @ b for tall-skinny operands (a: (K, M), b: (K, N), K = batch * L in the tens of thousands, M or N <= 64), fp32.
    16-bit operands go to the sm_100a kernel cm_tsmm (tensor-core partial blocks per row chunk + the deterministic
    reducer); otherwise one bmm over `nsplit` row blocks plus a sum (as a single GEMM cuBLAS picks a serial-K sm_75 CUTLASS
    kernel for these shapes on B200: 85 us per call at K = 32064, measured).
    ``defer``: the result goes straight to a parameter gradient (kernels.deferred_reductions).  ``row_perm`` = (r0, r1, r2):
    return rows [r1:r2] followed by rows [r0:r1] of the product (x_proj's rows back in the reference's [dt | B | C] order) -
    formed by the reducer itself where the split path runs, by a copy otherwise."""
    def permuted(full):
        if row_perm is None:
            return full
        r0, r1, r2 = row_perm
        return torch.cat([full[r1:r2], full[r0:r1]], dim=0)

    # measured in the bench step on B200: 16.15 vs 16.27 ms with cm_tsmm at 12032 x 288 (ConMamba-small), 54.9 vs 54.6 ms at
    # 32064 x 512 (ConMamba-large) - the kernel takes the shapes where it wins, the cuBLAS bmm split keeps the rest
    if _USE_TSMM and os.environ.get("CM_NO_TSMM") is None and a.shape[0] * max(a.shape[1], b.shape[1]) <= _TSMM_MAX_ELEMS:
        if K.tsmm_supported(a, b):
            return permuted(K.tsmm(a, b, defer=defer and row_perm is None))
        if K.tsmm_supported(b, a):
            return permuted(K.tsmm(b, a, defer=defer and row_perm is None).t())
    Kr = a.shape[0]
    od = {} if a.dtype == torch.float32 else {"out_dtype": torch.float32}   # fp32 straight out of the GEMM
    if nsplit <= 1 or Kr % nsplit != 0:
        return permuted(torch.mm(a.t(), b, **od))
    a3 = a.unflatten(0, (nsplit, Kr // nsplit))
    b3 = b.unflatten(0, (nsplit, Kr // nsplit))
    part = torch.bmm(a3.transpose(1, 2), b3, **od)
    if part.dtype != torch.float32 or not part.is_cuda or not part.is_contiguous():
        return permuted(part.sum(0))
    if row_perm is None:
        return K.sum_leading(part, defer=defer)
    r0, r1, r2 = row_perm
    M, Nc = part.shape[1], part.shape[2]
    out = torch.empty((r2 - r0, Nc), dtype=torch.float32, device=part.device)
    flat = out.view(-1)
    n1 = (r2 - r1) * Nc
    K.reduce_many([(part, flat[:n1], nsplit, n1, M * Nc, r1 * Nc),
                   (part, flat[n1:], nsplit, (r1 - r0) * Nc, M * Nc, r0 * Nc)], defer=defer)
    return out


def inner_forward(xz, ndir, out_scale, reverse0, params, need_grad=False, need_last_state=False):
    """conv -> x_proj -> dt_proj -> fused scan on channel-last buffers.  Returns (y (B, L, D), tensors to save for
    backward or None, per-direction time order, per-direction last states (B, D, N) fp32 or [])."""
    assert len(params) == ndir * MambaInnerCL.NPER
    if xz.stride(-1) != 1:
        xz = xz.contiguous()
    Bt, L, twoD = xz.shape
    D = twoD // 2
    act = xz.dtype
    x = _as_bdl(xz[..., :D])
    z = _as_bdl(xz[..., D:])
    P = [params[r * 7:(r + 1) * 7] for r in range(ndir)]
    rev = [bool(reverse0) ^ (r == 1) for r in range(ndir)]
    with torch.autocast("cuda", enabled=False):
        conv_dirs = [dict(weight=p[0][:, 0, :], bias=p[1], anticausal=rev[r]) for r, p in enumerate(P)]
        us = K.conv_forward(x, conv_dirs, silu=True)                     # logical (B, D, L), memory (B, L, D)
        scan_dirs, x_dbls, deltas, wx, wdt = [], [], [], [], []
        for r, p in enumerate(P):
            R = p[3].shape[1]
            N = p[4].shape[1]
            xw, dw = _aligned_proj_weights(p[2], p[3], R, N, act)        # rows [B | C | dt | 0-pad], K padded
            u_mem = us[r].transpose(1, 2)                                # (B, L, D) contiguous
            x_dbl = torch.mm(u_mem.reshape(Bt * L, D), xw.t())           # (B*L, 2N + Rp), 16-byte-aligned rows
            delta_mem = torch.mm(x_dbl[:, 2 * N:], dw.t()).view(Bt, L, D)
            xv = x_dbl.view(Bt, L, -1)
            scan_dirs.append(dict(u=us[r], delta=_as_bdl(delta_mem), A=p[4],
                                  B=_as_bdl(xv[..., :N]), C=_as_bdl(xv[..., N:2 * N]),
                                  D=p[5], delta_bias=p[6], reverse=rev[r]))
            x_dbls.append(x_dbl)
            deltas.append(delta_mem)
            wx.append(xw)
            wdt.append(dw)
        res = K.scan_forward(scan_dirs, z=z, out_scale=out_scale, delta_softplus=True,
                             need_ckpt=need_grad, need_out_pre=need_grad, need_last_state=need_last_state)
    y = res["out"].transpose(1, 2)                                       # (B, L, D) contiguous memory
    saved = None
    if need_grad:
        saved = (xz, res["out_pre"], *us, *deltas, *x_dbls, *res["ckpt"], *wx, *wdt, *params)
    return y, saved, rev, res["last_state"]


class MambaInnerCL(torch.autograd.Function):
    """y = MambaInnerCL.apply(xz, ndir, out_scale, reverse0, *params)

    xz: (B, L, 2*D) with unit last stride.  ``params`` holds, per direction,
    (conv_w (D,1,W), conv_b (D,)|None, x_proj_w (R+2N, D), dt_proj_w (D, R), A (D, N) fp32, Dskip (D,) fp32,
    dt_bias (D,) fp32).  Direction 0 runs in time order ``reverse0`` (False = causal), direction 1 the opposite.
    Returns y: (B, L, D) = out_scale * sum_dirs(scan) * silu(z).
    """

    NPER = 7

    @staticmethod
    def forward(ctx, xz, ndir, out_scale, reverse0, *params):
        need_grad = any(ctx.needs_input_grad)
        y, saved, rev, _ = inner_forward(xz, ndir, out_scale, reverse0, params, need_grad=need_grad)
        if need_grad:
            ctx.ndir, ctx.out_scale, ctx.rev = ndir, out_scale, rev
            # A made by bimamba._NegExpMany: that node runs the reduction queue before it reads dA
            ctx.defer_dA = all(getattr(params[r * 7 + 4], "_cm_batched_A", False) for r in range(ndir))
            ctx.save_for_backward(*saved)
        return y

    @staticmethod
    def backward(ctx, dy):
        ndir, rev = ctx.ndir, ctx.rev
        sv = ctx.saved_tensors
        xz, out_pre = sv[0], sv[1]
        k = 2
        us = sv[k:k + ndir]; k += ndir
        deltas = sv[k:k + ndir]; k += ndir
        x_dbls = sv[k:k + ndir]; k += ndir
        ckpts = sv[k:k + ndir]; k += ndir
        wx = sv[k:k + ndir]; k += ndir
        wdt = sv[k:k + ndir]; k += ndir
        params = sv[k:]
        P = [params[r * 7:(r + 1) * 7] for r in range(ndir)]
        Bt, L, twoD = xz.shape
        D = twoD // 2
        act = xz.dtype
        x = _as_bdl(xz[..., :D])
        z = _as_bdl(xz[..., D:])
        with torch.autocast("cuda", enabled=False):
            dxz = torch.empty_like(xz)
            scan_dirs, dx_dbls, dbc_like = [], [], []
            for r, p in enumerate(P):
                N = p[4].shape[1]
                xv = x_dbls[r].view(Bt, L, -1)
                scan_dirs.append(dict(u=us[r], delta=_as_bdl(deltas[r]), A=p[4],
                                      B=_as_bdl(xv[..., :N]), C=_as_bdl(xv[..., N:2 * N]),
                                      D=p[5], delta_bias=p[6], reverse=rev[r]))
                dxd = torch.empty_like(x_dbls[r])
                dv = dxd.view(Bt, L, -1)
                dbc_like.append((_as_bdl(dv[..., :N]), _as_bdl(dv[..., N:2 * N])))
                dx_dbls.append(dxd)
            g = K.scan_backward(scan_dirs, ckpts, _as_bdl(dy), z=z, out_pre=out_pre, out_scale=ctx.out_scale,
                                delta_softplus=True, dz_out=_as_bdl(dxz[..., D:]), dBC_like=dbc_like, defer=True,
                                defer_dA=ctx.defer_dA)
            grads = []
            conv_dirs, conv_douts = [], []
            for r, p in enumerate(P):
                R = p[3].shape[1]
                N = p[4].shape[1]
                u_mem = us[r].transpose(1, 2).reshape(Bt * L, D)
                ddelta = g["ddelta"][r].transpose(1, 2).reshape(Bt * L, D)
                du = g["du"][r].transpose(1, 2).reshape(Bt * L, D)
                dxd = dx_dbls[r]
                d_dtw = _wgrad(ddelta, x_dbls[r][:, 2 * N:], Bt, defer=True)[:, :R]   # (D, R)
                dxd[:, 2 * N:] = torch.mm(ddelta, wdt[r])                            # (B*L, Rp); pad columns get 0
                d_xw = _wgrad(dxd, u_mem, Bt, defer=True, row_perm=(0, 2 * N, 2 * N + R))   # (R + 2N, D): [dt | B | C] rows
                du.addmm_(dxd, wx[r])                                                # + x_proj back-prop (:282)
                conv_dirs.append(dict(weight=p[0][:, 0, :], bias=p[1], anticausal=rev[r]))
                conv_douts.append(g["du"][r])
                grads.append([None, None, K.grad_cast(d_xw, p[2].dtype), K.grad_cast(d_dtw, p[3].dtype),
                              K.grad_cast(g["dA"][r], p[4].dtype),
                              g["dD"][r], g["dbias"][r]])
            _, dws, dbs = K.conv_backward(x, conv_dirs, conv_douts, silu=True, dx_out=_as_bdl(dxz[..., :D]), defer=True)
            for r, p in enumerate(P):
                grads[r][0] = K.grad_cast(dws[r], p[0].dtype).unsqueeze(1)
                grads[r][1] = None if p[1] is None else K.grad_cast(dbs[r], p[1].dtype)
        flat = [t for gr in grads for t in gr]
        return (dxz, None, None, None, *flat)
