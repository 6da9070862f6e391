"""Drop-in for ``mamba_ssm.ops.triton.selective_state_update.selective_state_update`` as the reference calls it
(``modules/mamba/bimamba.py:28-31`` import, ``:359-362`` call): one decoding token of the selective SSM.

Upstream this is a Triton kernel; here it is the sm_100a ``cm_ssm_step`` kernel behind the C ABI.  CUDA tensors only,
no fallback.  Semantics = the torch expression the reference falls back to without the import (bimamba.py:350-357).
"""
from . import kernels as K


def selective_state_update(state, x, dt, A, B, C, D=None, z=None, dt_bias=None, dt_softplus=False):
    """state: (batch, dim, dstate), updated in place; x, dt, z: (batch, dim); A: (dim, dstate); B, C: (batch, dstate);
    D, dt_bias: (dim,).  Returns out: (batch, dim)."""
    if B.dim() != 2 or C.dim() != 2:
        raise NotImplementedError("grouped B / C (ngroups > 1) is not used by the ConMamba path")
    return K.ssm_step(state, x, dt, A, B, C, D=D, z=z, dt_bias=dt_bias, dt_softplus=dt_softplus)
