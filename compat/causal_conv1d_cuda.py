"""Shim reproducing the ``causal_conv1d_cuda`` pybind11 module of causal-conv1d 1.1.3.post1
(reference call sites modules/mamba/selective_scan_interface.py:182, :244, :286)."""
from mamba_asr_b200 import kernels as K


def causal_conv1d_fwd(x, weight, bias_, seq_idx_, silu):
    if seq_idx_ is not None:
        raise NotImplementedError("seq_idx is not used by the ConMamba path")
    return K.conv_forward(x, [dict(weight=weight, bias=bias_, anticausal=False)], silu=bool(silu))[0]


def causal_conv1d_bwd(x, weight, bias_, dout, seq_idx_, dx_, silu):
    if seq_idx_ is not None:
        raise NotImplementedError("seq_idx is not used by the ConMamba path")
    dx, dws, dbs = K.conv_backward(x, [dict(weight=weight, bias=bias_, anticausal=False)], [dout], silu=bool(silu),
                                   dx_out=dx_)
    dweight = dws[0].to(weight.dtype)
    dbias = dbs[0].to(bias_.dtype) if bias_ is not None else None
    return dx, dweight, dbias
