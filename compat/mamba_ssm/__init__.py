"""Shim for ``from mamba_ssm import Mamba`` (reference modules/Conmamba.py:124)."""
from mamba_asr_b200.bimamba import UniMamba as Mamba  # noqa: F401
