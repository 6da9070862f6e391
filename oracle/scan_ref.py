"""Selective-scan oracle (TEST INFRASTRUCTURE - see oracle/__init__.py).

Restates ``selective_scan_ref`` of the reference (modules/mamba/selective_scan_interface.py:91-157) for
real-valued ``A``.  Arithmetic order follows the reference line by line so that fp32 results are
bit-identical on the same torch build:

  :106-112  upcast u, delta to fp32; add delta_bias; softplus
  :113-123  B, C upcast
  :126      deltaA    = exp(delta[b,d,l] * A[d,n])
  :127-134  deltaB_u  = delta * B * u           (einsum 'bdl,bnl,bdl->bdln' for 3-D B)
  :138-151  h = deltaA[:, :, i] * h + deltaB_u[:, :, i];  y_i = <h, C_i>
  :153-156  y + u * D ; * silu(z) ; cast back to the input dtype

``compute_dtype=torch.float64`` gives the error-budget "truth" the reference itself cannot produce
(its einsum at :144 fails for an fp64 A).  Complex ``A`` (reference :114-118) is
never reached by the trainers and is not restated.

The function is differentiable (plain torch ops), so autograd through it is the backward oracle.
"""
import torch
import torch.nn.functional as F


def _expand_groups(M, dim):
    # reference :133 / :136  repeat(B, "B G N L -> B (G H) N L", H=dim // G)
    G = M.shape[1]
    return M.repeat_interleave(dim // G, dim=1)


def selective_scan_oracle(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                          return_last_state=False, reverse=False, compute_dtype=torch.float32):
    """Same arguments and return convention as the reference's ``selective_scan_ref``.

    ``reverse=True`` runs time from L-1 down to 0 (the backward direction of BiMamba-v2 expressed in
    original index space); it equals flip -> scan -> flip.
    """
    if A.is_complex():
        raise NotImplementedError("complex A is outside the ConMamba hot path")
    out_dtype = u.dtype
    cd = compute_dtype
    u_ = u.to(cd)
    dl = delta.to(cd)
    if delta_bias is not None:
        dl = dl + delta_bias[..., None].to(cd)
    if delta_softplus:
        dl = F.softplus(dl)
    Bt, Dm, L = u_.shape
    N = A.shape[1]
    A_ = A.to(cd)
    B_ = B.to(cd)
    C_ = C.to(cd)
    var_B = B_.dim() >= 3
    var_C = C_.dim() >= 3

    dA = torch.exp(torch.einsum("bdl,dn->bdln", dl, A_))
    if not var_B:
        dBu = torch.einsum("bdl,dn,bdl->bdln", dl, B_, u_)
    elif B_.dim() == 3:
        dBu = torch.einsum("bdl,bnl,bdl->bdln", dl, B_, u_)
    else:
        dBu = torch.einsum("bdl,bdnl,bdl->bdln", dl, _expand_groups(B_, Dm), u_)
    if var_C and C_.dim() == 4:
        C_ = _expand_groups(C_, Dm)

    h = A_.new_zeros((Bt, Dm, N))
    ys = [None] * L
    order = range(L - 1, -1, -1) if reverse else range(L)
    for i in order:
        h = dA[:, :, i] * h + dBu[:, :, i]
        if not var_C:
            ys[i] = torch.einsum("bdn,dn->bd", h, C_)
        elif C_.dim() == 3:
            ys[i] = torch.einsum("bdn,bn->bd", h, C_[:, :, i])
        else:
            ys[i] = torch.einsum("bdn,bdn->bd", h, C_[:, :, :, i])
    last_state = h
    y = torch.stack(ys, dim=2)
    out = y if D is None else y + u_ * D.to(cd)[:, None]
    if z is not None:
        # reference :155 applies silu in z's own dtype (no upcast); only the fp64 truth upcasts
        out = out * (F.silu(z) if cd == torch.float32 else F.silu(z.to(cd)))
    out = out.to(out_dtype)
    return (out, last_state) if return_last_state else out
