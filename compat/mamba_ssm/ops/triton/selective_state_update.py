"""Shim for ``from mamba_ssm.ops.triton.selective_state_update import selective_state_update``
(reference modules/mamba/bimamba.py:28-31): the sm_100a single-token kernel, not Triton."""
from mamba_asr_b200.selective_state_update import selective_state_update  # noqa: F401
