"""mamba_asr_b200 - B200-native (sm_100a) ConMamba hot path: Fbank tail, fused bidirectional causal conv1d+SiLU,
fused bidirectional selective scan, behind the reference's Python call signatures.

    from mamba_asr_b200 import Mamba, UniMamba, Fbank
    from mamba_asr_b200.selective_scan_interface import selective_scan_fn, mamba_inner_fn_no_out_proj
    from mamba_asr_b200.causal_conv1d import causal_conv1d_fn

All arithmetic runs in hand-written CUDA kernels reached through the C ABI in include/conmamba_b200.h
(lib/libconmamba_b200.so, built by ``python -m mamba_asr_b200.build``).  There is no CPU, Triton or
multi-backend fallback: calling any op without the library or with CPU tensors raises.
"""
from .bimamba import Mamba, UniMamba  # noqa: F401
from .fbank import Fbank  # noqa: F401

__version__ = "0.1.0"
