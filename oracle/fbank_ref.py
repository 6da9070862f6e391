"""SpeechBrain Fbank oracle (TEST INFRASTRUCTURE - see oracle/__init__.py).  PARITY UNPINNED.

The reference instantiates ``speechbrain.lobes.features.Fbank`` from YAML
(hparams/CTC/conmamba_large.yaml:322-326: sample_rate 16000, n_fft 512, n_mels 80, win_length 32;
hparams/S2S/conmamba_small.yaml:469-472: n_fft 400, win_length default 25) and calls it at
train_CTC.py:285 / train_S2S.py:349.  The arithmetic lives in speechbrain==1.0.0
(requirement.txt:11), which is neither vendored in the reference nor installable here, so this
restatement follows the published 1.0.0 algorithm:

  STFT              torch.stft(x, n_fft, hop=round(sr/1000*hop_ms), win=round(sr/1000*win_ms),
                    window=hamming_window(win) (periodic), center=True, pad_mode="constant",
                    normalized=False, onesided=True)  -> (B, T, n_fft//2+1, 2)
  spectral_magnitude(power=1)   re^2 + im^2
  Filterbank        mel = linspace(mel(f_min), mel(f_max), n_mels+2); hz = 700*(10**(mel/2595)-1)
                    band = (hz[1:]-hz[:-1])[:-1]; f_central = hz[1:-1]
                    all_freqs = linspace(0, sr//2, n_fft//2+1)
                    slope = (all_freqs - f_central)/band
                    fbank = max(0, min(slope+1, -slope+1))            (triangular)
                    fbanks = spectrogram @ fbank
  amplitude_to_DB   x_db = 10*log10(clamp(x, 1e-10)) - 10*log10(max(1e-10, 1.0))
                    x_db = max(x_db, x_db.amax((-2,-1)) - 80)   per utterance (padding included)

Deltas and context windows are off in every reference YAML.  Output (B, T, n_mels) fp32,
T = 1 + n_samples // hop.
"""
import math

import numpy as np
import torch


def _to_mel(hz):
    return 2595.0 * math.log10(1.0 + hz / 700.0)


def mel_filterbank_matrix(n_fft=512, n_mels=80, sample_rate=16000, f_min=0.0, f_max=None):
    """(n_fft//2+1, n_mels) fp32 triangular filterbank, built with the same torch fp32 ops."""
    if f_max is None:
        f_max = sample_rate / 2
    n_stft = n_fft // 2 + 1
    mel = torch.linspace(_to_mel(f_min), _to_mel(f_max), n_mels + 2)
    hz = 700.0 * (10.0 ** (mel / 2595.0) - 1.0)
    band = (hz[1:] - hz[:-1])[:-1]
    f_central = hz[1:-1]
    all_freqs = torch.linspace(0, sample_rate // 2, n_stft)
    all_freqs_mat = all_freqs.repeat(f_central.shape[0], 1)                      # (n_mels, n_stft)
    f_central_mat = f_central.repeat(all_freqs_mat.shape[1], 1).transpose(0, 1)
    band_mat = band.repeat(all_freqs_mat.shape[1], 1).transpose(0, 1)
    slope = (all_freqs_mat - f_central_mat) / band_mat
    left_side = slope + 1.0
    right_side = -slope + 1.0
    fbank = torch.max(torch.zeros(1), torch.min(left_side, right_side)).transpose(0, 1)
    return fbank.contiguous()


def stft_power(wav, n_fft=512, win_length_ms=32, hop_length_ms=10, sample_rate=16000):
    """(B, n_samples) fp32 -> (B, T, n_fft//2+1) power spectrum."""
    win = int(round((sample_rate / 1000.0) * win_length_ms))
    hop = int(round((sample_rate / 1000.0) * hop_length_ms))
    window = torch.hamming_window(win)
    st = torch.stft(wav.float(), n_fft, hop, win, window, center=True, pad_mode="constant",
                    normalized=False, onesided=True, return_complex=True)
    st = torch.view_as_real(st).transpose(2, 1)                                  # (B, T, F, 2)
    return st.pow(2).sum(-1)


def power_to_logmel(power, fbank, top_db=80.0, amin=1e-10):
    fb = torch.matmul(power, fbank)
    x_db = 10.0 * torch.log10(torch.clamp(fb, min=amin))
    x_db = x_db - 10.0 * math.log10(max(amin, 1.0))
    new_max = x_db.amax(dim=(-2, -1)) - top_db
    return torch.max(x_db, new_max.view(x_db.shape[0], 1, 1))


def fbank_oracle(wav, n_fft=512, n_mels=80, win_length_ms=32, hop_length_ms=10, sample_rate=16000):
    power = stft_power(wav, n_fft, win_length_ms, hop_length_ms, sample_rate)
    return power_to_logmel(power, mel_filterbank_matrix(n_fft, n_mels, sample_rate))


def fbank_numpy_dft(wav, n_fft=512, n_mels=80, win_length_ms=32, hop_length_ms=10, sample_rate=16000):
    """Independent float64 cross-check: explicit framing + rfft, no torch.stft."""
    wav = np.asarray(wav, dtype=np.float64)
    win = int(round((sample_rate / 1000.0) * win_length_ms))
    hop = int(round((sample_rate / 1000.0) * hop_length_ms))
    n = np.arange(win)
    w = 0.54 - 0.46 * np.cos(2.0 * np.pi * n / win)            # periodic hamming
    wpad = np.zeros(n_fft)
    left = (n_fft - win) // 2
    wpad[left:left + win] = w                                   # torch.stft centre-pads the window
    B, ns = wav.shape
    T = 1 + ns // hop
    padded = np.pad(wav, ((0, 0), (n_fft // 2, n_fft // 2)))
    idx = np.arange(T)[:, None] * hop + np.arange(n_fft)[None, :]
    frames = padded[:, idx] * wpad                              # (B, T, n_fft)
    spec = np.fft.rfft(frames, n=n_fft, axis=-1)
    power = spec.real ** 2 + spec.imag ** 2
    fb = power @ mel_filterbank_matrix(n_fft, n_mels, sample_rate).double().numpy()
    x_db = 10.0 * np.log10(np.maximum(fb, 1e-10))
    floor = x_db.max(axis=(1, 2), keepdims=True) - 80.0
    return np.maximum(x_db, floor)
