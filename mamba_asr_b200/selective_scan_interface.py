"""Drop-in for the reference's ``modules/mamba/selective_scan_interface.py`` - same names, same arguments,
same return conventions, backed by the sm_100a kernels (no CPU fallback; CUDA tensors only).

  selective_scan_fn          reference :82-88   (SelectiveScanFn :19-79)
  mamba_inner_fn_no_out_proj reference :632-638 (MambaInnerFnNoOutProj :160-294)
  mamba_inner_fn             reference :611-618 (MambaInnerFn :297-439)
  bimamba_inner_fn           reference :621-629 (BiMambaInnerFn :442-608, the shared-weight "v1" form)
  selective_scan_ref / mamba_inner_ref / bimamba_inner_ref (reference :91, :641, :678) are importable names that RAISE:
  the CPU reference lives in ``oracle/`` (test infrastructure); the product package has no CPU path.

Differences that are deliberate and invisible to callers:
  * inputs are consumed through their strides - nothing is ``.contiguous()``-copied (reference :24-35);
  * ``xz`` given in the reference's (B, 2D, L) layout is moved to channel-last once; results come back as
    (B, D, L) *views* of channel-last memory, so the caller's ``rearrange(out, "b d l -> b l d")`` is free;
  * gradients w.r.t. B/C are accumulated deterministically (no atomics).
"""
import torch
import torch.nn.functional as F

from . import kernels as K
from .causal_conv1d import causal_conv1d_fn  # noqa: F401  (re-exported like reference :14)
from .mamba_inner import MambaInnerCL


def _norm_bc(M, name):
    """(B, N, L) | (B, 1, N, L) | (D, N) -> kernel form, and a function restoring the caller's grad shape."""
    if M.dim() == 4:
        if M.shape[1] != 1:
            raise NotImplementedError("%s with n_groups > 1 is outside the ConMamba path (reference uses G = 1)" % name)
        return M[:, 0], (lambda g: g.unsqueeze(1))
    if M.dim() in (2, 3):
        return M, (lambda g: g)
    raise ValueError("%s must be (D, N), (B, N, L) or (B, 1, N, L)" % name)


class SelectiveScanFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                return_last_state=False):
        if A.is_complex():
            raise NotImplementedError("complex A is outside the ConMamba hot path")
        Bk, ctx.fixB = _norm_bc(B, "B")
        Ck, ctx.fixC = _norm_bc(C, "C")
        d = dict(u=u, delta=delta, A=A.float(), B=Bk, C=Ck, D=None if D is None else D.float(),
                 delta_bias=None if delta_bias is None else delta_bias.float(), reverse=False)
        need_grad = any(ctx.needs_input_grad)
        res = K.scan_forward([d], z=z, out_scale=1.0, delta_softplus=delta_softplus, need_ckpt=need_grad,
                             need_last_state=return_last_state, need_out_pre=need_grad and z is not None)
        ctx.delta_softplus = delta_softplus
        ctx.has_z = z is not None
        ctx.has_D = D is not None
        ctx.has_bias = delta_bias is not None
        if need_grad:
            ctx.save_for_backward(u, delta, d["A"], Bk, Ck, d["D"], z, d["delta_bias"], res["ckpt"][0], res["out_pre"])
        out = res["out"]
        if return_last_state:
            last = res["last_state"][0]
            ctx.mark_non_differentiable(last)
            return out, last
        return out

    @staticmethod
    def backward(ctx, dout, *args):
        u, delta, A, Bk, Ck, D, z, delta_bias, ckpt, out_pre = ctx.saved_tensors
        d = dict(u=u, delta=delta, A=A, B=Bk, C=Ck, D=D, delta_bias=delta_bias, reverse=False)
        g = K.scan_backward([d], [ckpt], dout, z=z, out_pre=out_pre, out_scale=1.0,
                            delta_softplus=ctx.delta_softplus)
        return (g["du"][0], g["ddelta"][0], g["dA"][0], ctx.fixB(g["dB"][0]), ctx.fixC(g["dC"][0]),
                g["dD"][0] if ctx.has_D else None, g["dz"] if ctx.has_z else None,
                g["dbias"][0] if ctx.has_bias else None, None, None)


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False):
    """if return_last_state is True, returns (out, last_state); last_state is (batch, dim, dstate) and its
    gradient is not considered in the backward pass (as in the reference, :84-87)."""
    return SelectiveScanFn.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)


def _check_inner_args(A, B, C, B_proj_bias, C_proj_bias, delta_softplus, checkpoint_lvl=1):
    assert checkpoint_lvl in [0, 1]          # reference :170
    if A.is_complex():
        raise NotImplementedError("complex A is outside the ConMamba hot path")
    if B is not None or C is not None:
        raise NotImplementedError("constant B / C in the fused inner function: the ConMamba configs always use "
                                  "input-dependent B and C (bimamba.py:230-231)")
    if B_proj_bias is not None or C_proj_bias is not None:
        raise NotImplementedError("B_proj_bias / C_proj_bias are never set by the reference models")
    if not delta_softplus:
        raise NotImplementedError("the fused inner function always applies softplus (reference default, :613)")


def _to_channel_last(xz):
    """(B, 2D, L) any strides -> (B, L, 2D) contiguous (one pass; free if xz already is a channel-last view)."""
    t = xz.transpose(1, 2)
    return t if t.is_contiguous() else t.contiguous()


def mamba_inner_fn_no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                               A, B=None, C=None, D=None, delta_bias=None, B_proj_bias=None,
                               C_proj_bias=None, delta_softplus=True, checkpoint_lvl=1):
    """xz: (batch, 2*dim, seqlen) -> out_z (batch, dim, seqlen).  Reference :632-638."""
    _check_inner_args(A, B, C, B_proj_bias, C_proj_bias, delta_softplus, checkpoint_lvl)
    y = MambaInnerCL.apply(_to_channel_last(xz), 1, 1.0, False, conv1d_weight, conv1d_bias, x_proj_weight,
                           delta_proj_weight, A.float(), None if D is None else D.float(),
                           None if delta_bias is None else delta_bias.float())
    return y.transpose(1, 2)


def mamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                   out_proj_weight, out_proj_bias,
                   A, B=None, C=None, D=None, delta_bias=None, B_proj_bias=None,
                   C_proj_bias=None, delta_softplus=True):
    """xz: (batch, 2*dim, seqlen) -> (batch, seqlen, d_model).  Reference :611-618 (out_proj fused in)."""
    _check_inner_args(A, B, C, B_proj_bias, C_proj_bias, delta_softplus)
    y = MambaInnerCL.apply(_to_channel_last(xz), 1, 1.0, False, conv1d_weight, conv1d_bias, x_proj_weight,
                           delta_proj_weight, A.float(), None if D is None else D.float(),
                           None if delta_bias is None else delta_bias.float())
    w = out_proj_weight.to(y.dtype)
    b = None if out_proj_bias is None else out_proj_bias.to(y.dtype)
    return F.linear(y, w, b)


def bimamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                     out_proj_weight, out_proj_bias,
                     A, A_b, B=None, C=None, D=None, delta_bias=None, B_proj_bias=None,
                     C_proj_bias=None, delta_softplus=True):
    """BiMamba "v1" (reference :621-629 / BiMambaInnerFn :442-608): ONE set of conv / projection weights, the
    scan run forward with A and backward (on the flipped sequence) with A_b, outputs summed, out_proj applied.
    In original index space the second pass is a time-reversed scan over the same u, delta, B, C."""
    _check_inner_args(A, B, C, B_proj_bias, C_proj_bias, delta_softplus)
    xz_cl = _to_channel_last(xz)
    Bt, L, twoD = xz_cl.shape
    Dn = twoD // 2
    act = xz_cl.dtype
    x = xz_cl[..., :Dn].transpose(1, 2)
    z = xz_cl[..., Dn:].transpose(1, 2)
    R = delta_proj_weight.shape[1]
    N = A.shape[1]
    u = causal_conv1d_fn(x, conv1d_weight[:, 0, :], conv1d_bias, activation="silu")     # (B, D, L) channel-last view
    x_dbl = F.linear(u.transpose(1, 2), x_proj_weight.to(act))                           # (B, L, R+2N)
    delta = F.linear(x_dbl[..., :R], delta_proj_weight.to(act)).transpose(1, 2)          # (B, D, L) view
    Bm = x_dbl[..., R:R + N].transpose(1, 2)
    Cm = x_dbl[..., R + N:R + 2 * N].transpose(1, 2)
    Df = None if D is None else D.float()
    bias = None if delta_bias is None else delta_bias.float()
    y_f = selective_scan_fn(u, delta, A.float(), Bm, Cm, Df, z, bias, True)
    fl = lambda t: t.flip(-1)
    y_b = selective_scan_fn(fl(u), fl(delta), A_b.float(), fl(Bm), fl(Cm), Df, fl(z), bias, True)
    y = y_f + fl(y_b)
    return F.linear(y.transpose(1, 2), out_proj_weight.to(act),
                    None if out_proj_bias is None else out_proj_bias.to(act))


def _cpu_reference_is_test_infrastructure(name, ref_line):
    def _raise(*args, **kwargs):
        raise NotImplementedError(
            "%s (reference selective_scan_interface.py:%s) is the CPU reference of this path: it is test infrastructure "
            "and lives in oracle/ (oracle.scan_ref.selective_scan_oracle, oracle.bimamba_ref); the product package has "
            "no CPU path.  Use %s on CUDA tensors instead." % (name, ref_line, name.replace("_ref", "_fn")))
    _raise.__name__ = name
    _raise.__doc__ = "Name kept for import compatibility with the reference (:%s); raises NotImplementedError." % ref_line
    return _raise


# names the reference module exports (:91, :641, :678): kept importable so `from ...selective_scan_interface import
# selective_scan_ref` does not fail at import time, but they refuse to run - the product path never touches a CPU oracle
selective_scan_ref = _cpu_reference_is_test_infrastructure("selective_scan_ref", "91-157")
mamba_inner_ref = _cpu_reference_is_test_infrastructure("mamba_inner_ref", "641-675")
bimamba_inner_ref = _cpu_reference_is_test_infrastructure("bimamba_inner_ref", "678-714")
