"""Sequence-length and padding-mask integers around the hot path, bit-exact with the reference.

  fbank_frames / encoder_frames   frame counts of the Fbank (hop 10 ms, centre padding) and of the two stride-2 convs
  abs_lengths                     ``torch.round(wav_len * L)``                    (modules/TransformerASR.py:409)
  key_padding_mask                ``~length_to_mask(abs_len).bool()``              (modules/TransformerASR.py:410)

ConMamba itself drops the mask (modules/Conmamba.py:631-635: padding is never masked inside the encoder); the lengths
still feed the CTC loss and input normalisation, so they must match the reference to the integer.
"""
import torch


def fbank_frames(n_samples: int, hop: int = 160) -> int:
    return 1 + n_samples // hop


def encoder_frames(n_frames: int) -> int:
    t1 = (n_frames - 1) // 2 + 1
    return (t1 - 1) // 2 + 1


def abs_lengths(wav_len: torch.Tensor, L: int) -> torch.Tensor:
    """fp32 tensor, round-half-to-even exactly as ``torch.round(wav_len * src.shape[1])``."""
    return torch.round(wav_len * L)


def length_to_mask(length: torch.Tensor, max_len=None) -> torch.Tensor:
    """speechbrain.dataio.dataio.length_to_mask: mask[b, t] = t < length[b]; width = max(length) unless given."""
    if max_len is None:
        max_len = int(length.max().long().item())
    ar = torch.arange(max_len, device=length.device, dtype=length.dtype)
    return (ar.unsqueeze(0) < length.unsqueeze(1))


def key_padding_mask(wav_len: torch.Tensor, L: int) -> torch.Tensor:
    """True where padded; shape (B, max(abs_len)) like the reference."""
    return ~length_to_mask(abs_lengths(wav_len, L)).bool()
