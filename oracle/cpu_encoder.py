"""CPU reference path of the whole encoder (TEST INFRASTRUCTURE - see oracle/__init__.py).

Used by ``bench.py``'s ``cpu_baseline`` leg and ``bench.py --impl reference``: the same module tree the GPU arm
runs, with the two hot-path pieces swapped for the reference's own CPU arithmetic -
``selective_scan_ref`` + torch causal conv composed as modules/mamba/bimamba.py:223-253 (``bimamba_v2_oracle``)
and the SpeechBrain Fbank restatement (``fbank_oracle``) - exactly the "reference CPU path
(selective_scan_ref + torch conv + Fbank)" BASELINE.json asks to time beside the GPU numbers.
"""
import torch
import torch.nn as nn

from .bimamba_ref import bimamba_v2_oracle
from .fbank_ref import fbank_oracle


class OracleBiMamba(nn.Module):
    """Holds the parameters of a product ``Mamba`` module and evaluates them with the CPU oracle."""

    def __init__(self, mamba):
        super().__init__()
        self.inner = mamba                      # parameters stay registered under the same names
        self.if_devide_out = mamba.if_devide_out

    def forward(self, hidden):
        p = dict(self.inner.named_parameters())
        return bimamba_v2_oracle(hidden, p, if_devide_out=self.if_devide_out)


class OracleFbank(nn.Module):
    def __init__(self, n_fft, n_mels, win_length_ms):
        super().__init__()
        self.n_fft, self.n_mels, self.win_length_ms = n_fft, n_mels, win_length_ms

    @torch.no_grad()
    def forward(self, wav):
        return fbank_oracle(wav, n_fft=self.n_fft, n_mels=self.n_mels, win_length_ms=self.win_length_ms)


def to_cpu_reference(model, n_fft, n_mels, win_length_ms):
    """In-place: swap the sm_100a mixers / Fbank of a ``ConMambaCTC`` (built on CPU) for the oracle versions."""
    from mamba_asr_b200.layernorm import FusedLayerNorm
    for layer in model.encoder.layers:
        layer.mamba = OracleBiMamba(layer.mamba)
    for mod in model.modules():                  # the sm_100a LayerNorm has no CPU path: torch's own op on the CPU arm
        if isinstance(mod, FusedLayerNorm):
            mod.__class__ = nn.LayerNorm
        if hasattr(mod, "use_kernel"):           # ConvolutionModule: the reference's transpose -> nn.Conv1d chain
            mod.use_kernel = False
    model.compute_features = OracleFbank(n_fft, n_mels, win_length_ms)
    return model
