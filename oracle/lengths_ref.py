"""Sequence-length / padding-mask integer oracle (TEST INFRASTRUCTURE - see oracle/__init__.py).

Restates, bit-exactly, the integer arithmetic around the hot path:

  * Fbank frame count         T = 1 + n_samples // hop                 (torch.stft, center=True)
  * ConvolutionFrontEnd       two stride-2 "same"-padded convs: L = ((T-1)//2 + 1 - 1)//2 + 1
                              (hparams/CTC/conmamba_large.yaml:187-194)
  * modules/TransformerASR.py:408-410
        abs_len = torch.round(wav_len * src.shape[1])                  (round-half-to-even, fp32)
        src_key_padding_mask = ~length_to_mask(abs_len).bool()
    with speechbrain.dataio.dataio.length_to_mask:  max_len = abs_len.max();
        mask[b, t] = t < abs_len[b]
"""
import numpy as np
import torch


def fbank_frames(n_samples, hop=160):
    return 1 + n_samples // hop


def encoder_frames(T):
    t1 = (T - 1) // 2 + 1
    return (t1 - 1) // 2 + 1


def abs_lengths(wav_len, L):
    """wav_len: float tensor in (0, 1]; returns fp32 tensor exactly as torch.round(wav_len * L)."""
    return torch.round(wav_len.float() * L)


def key_padding_mask(wav_len, L):
    """True where padded.  Shape (B, max(abs_len)) like the reference (not (B, L))."""
    abs_len = abs_lengths(wav_len, L)
    max_len = int(abs_len.max().long().item())
    ar = torch.arange(max_len).unsqueeze(0).expand(abs_len.shape[0], max_len)
    return ~(ar < abs_len.unsqueeze(1))


def abs_lengths_numpy(wav_len, L):
    """Independent restatement with numpy's round-half-to-even on fp32 products."""
    prod = np.asarray(wav_len, dtype=np.float32) * np.float32(L)
    return np.rint(prod).astype(np.float32)
