"""Shim for the ``causal_conv1d`` pip package (reference modules/mamba/bimamba.py:19)."""
from mamba_asr_b200.causal_conv1d import causal_conv1d_fn, causal_conv1d_update  # noqa: F401
