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


class OracleUniMamba(nn.Module):
    """Causal mixer (``mamba_ssm.Mamba`` as the decoder uses it, reference modules/Conmamba.py:854-862) on the CPU oracle:
    in_proj -> MambaInnerFn body (selective_scan_interface.py:297-370) -> out_proj."""

    def __init__(self, mamba):
        super().__init__()
        self.inner = mamba

    def forward(self, hidden, inference_params=None, keep_last=None):
        import torch.nn.functional as F
        from .bimamba_ref import mamba_inner_oracle
        p = dict(self.inner.named_parameters())
        Bt, L, d = hidden.shape
        xz = (p["in_proj.weight"] @ hidden.reshape(Bt * L, d).t()).reshape(-1, Bt, L).transpose(0, 1)
        y = mamba_inner_oracle(xz, p["conv1d.weight"], p["conv1d.bias"], p["x_proj.weight"], p["dt_proj.weight"],
                               -torch.exp(p["A_log"].float()), p["D"].float(), p["dt_proj.bias"].float())
        out = F.linear(y.transpose(1, 2), p["out_proj.weight"], None)
        return out if keep_last is None else out[:, -keep_last:]


class OracleFbank(nn.Module):
    def __init__(self, n_fft, n_mels, win_length_ms):
        super().__init__()
        self.n_fft, self.n_mels, self.win_length_ms = n_fft, n_mels, win_length_ms

    @torch.no_grad()
    def forward(self, wav):
        return fbank_oracle(wav, n_fft=self.n_fft, n_mels=self.n_mels, win_length_ms=self.win_length_ms)


def to_cpu_reference(model, n_fft, n_mels, win_length_ms):
    """In-place: swap the sm_100a mixers / Fbank of a ``ConMambaCTC`` / ``ConMambaS2S`` (built on CPU) for the oracle
    versions."""
    from mamba_asr_b200.bimamba import Mamba as BiMamba, UniMamba
    from mamba_asr_b200.layernorm import FusedLayerNorm
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            if isinstance(child, BiMamba):
                setattr(parent, name, OracleBiMamba(child))
            elif isinstance(child, UniMamba):
                setattr(parent, name, OracleUniMamba(child))
    for mod in model.modules():                  # the sm_100a LayerNorm has no CPU path: torch's own op on the CPU arm
        if isinstance(mod, FusedLayerNorm):
            mod.__class__ = nn.LayerNorm
        if hasattr(mod, "use_kernel"):           # ConvolutionModule: the reference's transpose -> nn.Conv1d chain
            mod.use_kernel = False
    model.compute_features = OracleFbank(n_fft, n_mels, win_length_ms)
    return model
