"""Shim reproducing the ``selective_scan_cuda`` pybind11 module of mamba-ssm 1.1.3.post1 on top of the B200 kernels.

Argument order and return lists follow the reference call sites
(modules/mamba/selective_scan_interface.py:42, :67-70, :218, :252-256), so the reference's
``SelectiveScanFn`` / ``MambaInnerFnNoOutProj`` / ``MambaInnerFn`` run unmodified on top of it.

``x`` (the "scan intermediates" handed from fwd to bwd) keeps the reference's contract
``x[:, :, -1, 1::2] == last_state`` (selective_scan_interface.py:45); the rows before the last one carry this
library's 8-step state checkpoints.
"""
import torch
import torch.nn.functional as F

from mamba_asr_b200 import kernels as K


def _bc(M, name):
    if M.dim() == 4:
        if M.shape[1] != 1:
            raise NotImplementedError("%s: n_groups > 1 is outside the ConMamba path" % name)
        return M[:, 0]
    return M


def _x_geometry(L, N):
    nck = K.num_ckpt(L, 1)
    rows = (nck * 16 + 2 * N - 1) // (2 * N) + 1
    return nck, rows


def _ckpt_view(x, N):
    """(B, D, nck, 16) strided view over the leading floats of each (b, d) row of x."""
    Bt, D, rows, twoN = x.shape
    nck = ((rows - 1) * twoN) // 16
    return x.as_strided((Bt, D, nck, 16), (D * rows * twoN, rows * twoN, 16, 1), x.storage_offset())


def fwd(u, delta, A, B, C, D_, z_, delta_bias_, delta_softplus):
    if A.is_complex():
        raise NotImplementedError("complex A is outside the ConMamba hot path")
    Bt, Dm, L = u.shape
    N = A.shape[1]
    _, rows = _x_geometry(L, N)
    x = torch.empty((Bt, Dm, rows, 2 * N), dtype=torch.float32, device=u.device)
    ck = _ckpt_view(x, N)
    lib_dir = dict(u=u, delta=delta, A=A.float(), B=_bc(B, "B"), C=_bc(C, "C"),
                   D=None if D_ is None else D_.float(),
                   delta_bias=None if delta_bias_ is None else delta_bias_.float(), reverse=False)
    res = _scan_fwd_into(lib_dir, z_, delta_softplus, ck, x[:, :, -1, 1::2])
    if z_ is None:
        return [res["out"], x]
    return [res["out_pre"], x, res["out"]]


def _scan_fwd_into(d, z, softplus, ck, last_view):
    """scan_forward writing checkpoints / last state into caller-provided strided views."""
    import ctypes as C
    from mamba_asr_b200 import _cabi as cabi
    lib = cabi.lib()
    Bt, Dm, L, N, const_bc = K._check_dirs([d])
    a = cabi.ScanFwdArgs()
    a.batch, a.dim, a.seqlen, a.dstate = Bt, Dm, L, N
    a.ndir, a.dtype = 1, cabi.dtype_code(d["u"].dtype)
    a.flags = cabi.CM_FLAG_DELTA_SOFTPLUS if softplus else 0
    a.out_scale = 1.0
    keep = []
    K._fill_scan_dir(a.dir[0], d, keep, const_bc)
    a.dir[0].ckpt = ck.data_ptr()
    a.dir[0].ckpt_sb, a.dir[0].ckpt_sd = ck.stride(0), ck.stride(1)
    a.dir[0].last_state = last_view.data_ptr()
    a.dir[0].ls_sb, a.dir[0].ls_sd, a.dir[0].ls_sn = last_view.stride()
    out = K.empty_like_bdl(d["u"])
    out_pre = K.empty_like_bdl(d["u"]) if z is not None else None
    a.z, a.out, a.out_pre = cabi.t3(z), cabi.t3(out), cabi.t3(out_pre)
    cabi.check(lib.cm_scan_fwd(C.byref(a), cabi.stream_ptr()), "cm_scan_fwd")
    return dict(out=out, out_pre=out_pre)


def bwd(u, delta, A, B, C, D_, z_, delta_bias_, dout, x_, out_, dz_, delta_softplus, recompute_out_z):
    N = A.shape[1]
    d = dict(u=u, delta=delta, A=A.float(), B=_bc(B, "B"), C=_bc(C, "C"),
             D=None if D_ is None else D_.float(),
             delta_bias=None if delta_bias_ is None else delta_bias_.float(), reverse=False)
    g = K.scan_backward([d], [_ckpt_view(x_, N)], dout, z=z_, out_pre=out_, out_scale=1.0,
                        delta_softplus=delta_softplus, dz_out=dz_)
    dB, dC = g["dB"][0], g["dC"][0]
    if B.dim() == 4:
        dB = dB.unsqueeze(1)
    if C.dim() == 4:
        dC = dC.unsqueeze(1)
    dD = g["dD"][0] if D_ is not None else torch.zeros(0, device=u.device)
    dbias = g["dbias"][0] if delta_bias_ is not None else torch.zeros(0, device=u.device)
    ret = [g["du"][0], g["ddelta"][0], g["dA"][0].contiguous(), dB, dC, dD, dbias]
    if z_ is not None:
        ret.append(g["dz"])
        if recompute_out_z:
            ret.append(out_ * F.silu(z_))
    return ret
