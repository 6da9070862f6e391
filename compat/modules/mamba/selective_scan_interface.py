"""Shim for ``modules.mamba.selective_scan_interface`` (reference file of the same name)."""
from mamba_asr_b200.selective_scan_interface import *  # noqa: F401,F403
from mamba_asr_b200.selective_scan_interface import (SelectiveScanFn, bimamba_inner_fn, causal_conv1d_fn,  # noqa: F401
                                                      mamba_inner_fn, mamba_inner_fn_no_out_proj, selective_scan_fn,
                                                      selective_scan_ref, mamba_inner_ref, bimamba_inner_ref)
