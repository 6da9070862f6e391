"""Shim for ``modules.mamba.bimamba`` (reference modules/mamba/bimamba.py)."""
from mamba_asr_b200.bimamba import Mamba  # noqa: F401
from mamba_asr_b200.causal_conv1d import causal_conv1d_fn, causal_conv1d_update  # noqa: F401
from mamba_asr_b200.selective_scan_interface import (bimamba_inner_fn, mamba_inner_fn,  # noqa: F401
                                                      mamba_inner_fn_no_out_proj, selective_scan_fn)
