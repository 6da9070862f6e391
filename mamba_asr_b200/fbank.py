"""``Fbank`` - drop-in for ``speechbrain.lobes.features.Fbank`` as the reference instantiates it
(hparams/CTC/conmamba_large.yaml:322-326; call sites train_CTC.py:285, train_S2S.py:349).

For the transform sizes of the reference YAMLs (n_fft 400 / 512) the whole front-end is ONE sm_100a kernel,
``cm_fbank_wav_logmel``: windowed DFT (SpeechBrain's STFT arguments: hamming window zero-padded to n_fft, centre padding
with zeros, one-sided), power, mel projection, dB and the per-utterance max, followed by ``cm_fbank_floor`` for the top_db
floor - the complex STFT never exists in memory.  Other sizes (and ``CM_FBANK_CUFFT=1``, the A/B switch) take
``torch.stft`` (cuFFT) + ``cm_fbank_logmel``.  Output (B, T, n_mels) fp32, no grad (the reference builds Fbank with
``requires_grad=False`` and runs it under ``torch.no_grad`` semantics).
"""
import math
import os

import torch
import torch.nn as nn

from . import kernels as K


def _to_mel(hz):
    return 2595.0 * math.log10(1.0 + hz / 700.0)


def triangular_filterbank(n_fft, n_mels, sample_rate, f_min, f_max):
    """(n_fft//2+1, n_mels) fp32 - SpeechBrain 1.0.0 ``Filterbank`` with triangular filters, built with the same
    fp32 torch ops so the matrix is bit-identical to the one SpeechBrain multiplies by."""
    n_stft = n_fft // 2 + 1
    mel = torch.linspace(_to_mel(f_min), _to_mel(f_max), n_mels + 2)
    hz = 700.0 * (10.0 ** (mel / 2595.0) - 1.0)
    band = (hz[1:] - hz[:-1])[:-1]
    f_central = hz[1:-1]
    all_freqs = torch.linspace(0, sample_rate // 2, n_stft)
    slope = (all_freqs.unsqueeze(0) - f_central.unsqueeze(1)) / band.unsqueeze(1)      # (n_mels, n_stft)
    fb = torch.clamp(torch.minimum(slope + 1.0, -slope + 1.0), min=0.0)
    return fb.t().contiguous()


class Fbank(nn.Module):
    def __init__(self, deltas=False, context=False, requires_grad=False, sample_rate=16000, f_min=0, f_max=None,
                 n_fft=400, n_mels=40, filter_shape="triangular", param_change_factor=1.0, param_rand_factor=0.0,
                 left_frames=5, right_frames=5, win_length=25, hop_length=10):
        super().__init__()
        if deltas or context:
            raise NotImplementedError("deltas / context windows are off in every reference YAML")
        if requires_grad:
            raise NotImplementedError("learnable filterbanks are not used by the reference (requires_grad=False)")
        if filter_shape != "triangular":
            raise NotImplementedError("only triangular filters (reference default)")
        self.sample_rate = sample_rate
        self.n_fft = n_fft
        self.n_mels = n_mels
        self.win_length = int(round((sample_rate / 1000.0) * win_length))
        self.hop_length = int(round((sample_rate / 1000.0) * hop_length))
        f_max = sample_rate / 2 if f_max is None else f_max
        self.register_buffer("window", torch.hamming_window(self.win_length), persistent=False)
        self.register_buffer("fbank_matrix", triangular_filterbank(n_fft, n_mels, sample_rate, f_min, f_max),
                             persistent=False)
        # analysis window zero-padded (centred) to n_fft, as torch.stft pads it; support [lo, hi) of every mel filter
        left = (n_fft - self.win_length) // 2
        wpad = torch.zeros(n_fft)
        wpad[left:left + self.win_length] = self.window
        self.register_buffer("window_padded", wpad, persistent=False)
        nz = self.fbank_matrix != 0
        idx = torch.arange(nz.shape[0]).unsqueeze(1)
        lo = torch.where(nz, idx, torch.full_like(idx, nz.shape[0])).amin(0)
        hi = torch.where(nz, idx + 1, torch.zeros_like(idx)).amax(0)
        self.register_buffer("band", torch.stack([lo, hi], 1).to(torch.int32).contiguous(), persistent=False)
        self.top_db = 80.0
        self.amin = 1e-10
        self.multiplier = 10.0                       # power spectrogram
        self.db_offset = self.multiplier * math.log10(max(self.amin, 1.0))

    @torch.no_grad()
    def forward(self, wav):
        """wav: (B, n_samples) float -> (B, 1 + n_samples // hop, n_mels) fp32."""
        if wav.dim() != 2:
            raise NotImplementedError("multi-channel audio (B, T, C) is not used by the reference recipes")
        if not wav.is_cuda:
            raise RuntimeError("mamba_asr_b200.Fbank runs on CUDA only (no CPU fallback)")
        if self.win_length <= self.n_fft and K.fbank_wav_supported(self.n_fft) and os.environ.get("CM_FBANK_CUFFT") is None:
            dev = wav.device
            return K.fbank_wav_logmel(wav.float(), self.window_padded.to(dev), self.fbank_matrix.to(dev), self.band.to(dev),
                                      self.n_fft, self.hop_length, self.top_db, self.amin, self.multiplier, self.db_offset)
        with torch.autocast("cuda", enabled=False):
            stft = torch.stft(wav.float(), self.n_fft, self.hop_length, self.win_length, self.window.to(wav.device),
                              center=True, pad_mode="constant", normalized=False, onesided=True,
                              return_complex=True)                                     # (B, F, T) complex64
            return K.fbank_logmel(stft, self.fbank_matrix.to(wav.device), self.top_db, self.amin, self.multiplier, self.db_offset)
