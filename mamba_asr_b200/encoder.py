"""SpeechBrain-free assembly of the reference's CTC encoder stack, for the parity harness and ``bench.py``.

Mirrors what ``train_CTC.py:ASR.compute_forward`` runs (reference train_CTC.py:285-302) with the objects the
YAML builds (hparams/CTC/conmamba_large.yaml:187-227, 322-326):

    Fbank -> InputNormalization(global) -> ConvolutionFrontEnd (2 x [conv 3x3 stride 2, LayerNorm, LeakyReLU])
          -> TransformerASR.custom_src_module (Linear 640 -> d_model, Dropout)   (TransformerASR.py:773)
          -> ConmambaEncoder (N x ConmambaEncoderLayer, final LayerNorm)
          -> ctc_lin -> log_softmax

Only Fbank and the Mamba mixers are hand-written kernels; the surrounding layers are torch (cuBLAS / cuDNN)
stand-ins for the SpeechBrain modules, which are not installable here.  Parameter
initialisation follows ``TransformerASR._init_params`` (xavier_normal_ on every >= 2-D parameter of the
Transformer block, TransformerASR.py:1051-1054).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import math
import os

from .layernorm import DropoutSeed, conv_ln_act_stem, conv_ln_act_stem_supported, layer_norm_leaky_relu
from .linear import BiasGradLinear, ParamCache, set_param_cache
from .bimamba import precomputed_A
from .conmamba import ConmambaEncoder, MambaDecoder
from .fbank import Fbank

# model blocks of the BASELINE.json configs
CONFIGS = {
    # hparams/S2S/conmamba_small.yaml:188-189,229-236 model block on the CTC recipe
    "conmamba_small_ctc": dict(d_model=144, d_ffn=1024, num_layers=12, n_fft=400, win_length=25, n_mels=80,
                               output_neurons=31, seed=7775),
    # hparams/CTC/conmamba_large.yaml:153-183; features :101-105 (n_fft 512 with a 25 ms window: zero-padded to the transform)
    "conmamba_large_ctc": dict(d_model=256, d_ffn=1024, num_layers=18, n_fft=512, win_length=25, n_mels=80,
                               output_neurons=31, seed=3402),
    # encoder of hparams/S2S/conmambamamba_large.yaml:251-287
    "conmamba_large_s2s_encoder": dict(d_model=512, d_ffn=2048, num_layers=12, n_fft=512, win_length=32, n_mels=80,
                                       output_neurons=5000, seed=3407),
    # hparams/S2S/conmambamamba_large.yaml:251-315: 12 ConMamba encoder layers + 6 Mamba decoder layers, d_model 512
    "conmambamamba_large_s2s": dict(d_model=512, d_ffn=2048, num_layers=12, num_decoder_layers=6, n_fft=512,
                                    win_length=32, n_mels=80, output_neurons=5000, seed=3407),
}


class InputNormalization(nn.Module):
    """Stand-in for speechbrain InputNormalization(norm_type="global") on its first batch: per-feature mean / std
    over the valid frames of the batch."""

    def forward(self, feats, wav_lens=None):
        Bt, T, Fd = feats.shape
        if wav_lens is None:
            mean = feats.mean(dim=(0, 1))
            std = feats.std(dim=(0, 1))
        else:
            n = torch.round(wav_lens.float() * T).clamp(min=1)
            mask = (torch.arange(T, device=feats.device)[None, :] < n[:, None]).unsqueeze(-1).to(feats.dtype)
            cnt = mask.sum(dim=1)
            mean_u = (feats * mask).sum(dim=1) / cnt
            var_u = (((feats - mean_u[:, None]) ** 2) * mask).sum(dim=1) / (cnt - 1).clamp(min=1)
            mean, std = mean_u.mean(0), var_u.sqrt().mean(0)
        return (feats - mean) / std.clamp(min=1e-10)


class ConvFrontEnd(nn.Module):
    """Stand-in for speechbrain ConvolutionFrontEnd(num_blocks=2, out_channels=(64, 32), kernel 3, stride 2)."""

    def __init__(self, n_mels=80, out_channels=(64, 32)):
        super().__init__()
        blocks, c_in, f = [], 1, n_mels
        self.norms = nn.ModuleList()
        self.convs = nn.ModuleList()
        for c_out in out_channels:
            self.convs.append(nn.Conv2d(c_in, c_out, kernel_size=3, stride=2, padding=1))
            f = (f - 1) // 2 + 1
            self.norms.append(nn.LayerNorm([f, c_out]))
            c_in = c_out
        self.out_features = f * c_in
        # LayerNorm([F', C]) + LeakyReLU of each block on the fused sm_100a kernel (cm_ln_act_*); the CPU reference arm
        # (oracle/cpu_encoder.py) clears the flag and gets the two torch ops
        self.use_kernel = True

    def forward(self, feats):
        # channels-last memory throughout: the (B, T', F', C) view the LayerNorm wants is then the conv output's own
        # memory order, and cuDNN's bf16 kernels take NHWC directly (as NCHW the two blocks cost four nchwToNhwc kernels
        # and five permute copies per step - 3 ms of the 61 ms ConMamba-large step on B200)
        feats = feats.contiguous()
        Bt, T, Fd = feats.shape
        # (B, 1, T, F) with explicit NHWC strides: with one channel the default strides are ambiguous and torch resolves
        # the ambiguity to NCHW, which makes cuDNN convert the 64-channel output back and forth
        x = feats.as_strided((Bt, 1, T, Fd), (T * Fd, 1, Fd, 1))
        fused = self.use_kernel and os.environ.get("CM_NO_FUSE_FRONTEND_LN") is None
        stem = (fused and os.environ.get("CM_NO_FUSE_STEM") is None
                and conv_ln_act_stem_supported(feats, self.convs[0], self.norms[0]))
        for i, (conv, norm) in enumerate(zip(self.convs, self.norms)):
            if i == 0 and stem:
                # block 1 (one input channel) as ONE kernel: the conv output is never written (cm_stem_fwd / cm_stem_bwd)
                x = conv_ln_act_stem(feats, conv, norm).permute(0, 3, 1, 2)
                continue
            # fused: the conv bias is added inside the norm kernel (as a separate torch add it is one more pass over the
            # largest activation of the step)
            x = F.conv2d(x, conv.weight.contiguous(memory_format=torch.channels_last), None if fused else conv.bias,
                         conv.stride, conv.padding)                # (B, C, T', F'), NHWC memory
            x = x.permute(0, 2, 3, 1)                              # (B, T', F', C), contiguous
            if fused:
                x = layer_norm_leaky_relu(x, norm, pre_bias=conv.bias)
            else:
                x = F.leaky_relu(norm(x))
            x = x.permute(0, 3, 1, 2)
        x = x.permute(0, 2, 3, 1)                                  # (B, L, F'', C)
        return x.reshape(x.shape[0], x.shape[1], -1)               # (B, L, 640)  (TransformerASR.py:760-762)


class ConMambaCTC(nn.Module):
    def __init__(self, d_model, d_ffn, num_layers, n_fft=512, win_length=32, n_mels=80, output_neurons=31,
                 dropout=0.1, seed=None, d_state=16, expand=2, d_conv=4, bidirectional=True):
        super().__init__()
        self.compute_features = Fbank(sample_rate=16000, n_fft=n_fft, n_mels=n_mels, win_length=win_length)
        self.normalize = InputNormalization()
        self.CNN = ConvFrontEnd(n_mels)
        self.custom_src_module = nn.Sequential(BiasGradLinear(self.CNN.out_features, d_model), nn.Dropout(dropout))
        mamba_config = dict(d_state=d_state, expand=expand, d_conv=d_conv, bidirectional=bidirectional)
        # Transformer.py:740-751: encoder activation is branchformer_activation (GELU), kernel 31, bias, non-causal
        self.encoder = ConmambaEncoder(num_layers=num_layers, d_model=d_model, d_ffn=d_ffn, kernel_size=31,
                                       activation=nn.GELU, bias=True, dropout=dropout, causal=False,
                                       mamba_config=mamba_config)
        self.ctc_lin = BiasGradLinear(d_model, output_neurons)
        for p in list(self.custom_src_module.parameters()) + list(self.encoder.parameters()):
            if p.dim() > 1:
                nn.init.xavier_normal_(p)                          # TransformerASR.py:1051-1054

    def enable_param_cache(self, dtype=torch.bfloat16):
        """Keep bf16 copies of all parameters in one flat buffer, refreshed by one multi-tensor copy at the start of every
        autocast forward (linear.ParamCache): removes the per-parameter cast kernels of bf16 autocast training.  Call
        after the model is on its device."""
        self._param_cache = ParamCache(self, dtype)
        set_param_cache(self._param_cache)
        return self

    def _refresh_param_cache(self):
        c = getattr(self, "_param_cache", None)
        if c is not None and torch.is_autocast_enabled("cuda"):      # training and evaluation alike: never a stale copy
            set_param_cache(c)
            c.refresh()

    def _advance_dropout_seed(self, device):
        """Fresh masks for the fused dropout kernels on every forward: an in-place add on the device seed, so a captured
        CUDA graph replays it (call ids and the seed POINTER are baked into the graph, the seed VALUE is not)."""
        if self.training and device.type == "cuda":
            DropoutSeed.advance(device)

    def features(self, wavs, wav_lens=None):
        feats = self.compute_features(wavs)                        # (B, T, 80) fp32, no grad
        return self.normalize(feats, wav_lens)

    def encode(self, feats):
        src = self.custom_src_module(self.CNN(feats))
        out, _ = self.encoder(src)
        return out

    def forward(self, wavs, wav_lens=None):
        """wavs: (B, n_samples) -> log-probs (B, L, output_neurons)"""
        self._refresh_param_cache()
        self._advance_dropout_seed(wavs.device)
        with precomputed_A(self):
            enc = self.encode(self.features(wavs, wav_lens))
        return F.log_softmax(self.ctc_lin(enc), dim=-1)


class NormalizedEmbedding(nn.Module):
    """speechbrain NormalizedEmbedding (TransformerASR.py:738-740): token embedding scaled by sqrt(d_model)."""

    def __init__(self, d_model, vocab):
        super().__init__()
        self.emb = nn.Embedding(vocab, d_model, padding_idx=0)
        self.d_model = d_model

    def forward(self, x):
        return self.emb(x) * math.sqrt(self.d_model)


class PositionalEncoding(nn.Module):
    """speechbrain fixed sinusoidal PositionalEncoding, added to the decoder input (TransformerASR.py:794)."""

    def __init__(self, d_model, max_len=2500):
        super().__init__()
        pe = torch.zeros(max_len, d_model)
        pos = torch.arange(0, max_len).unsqueeze(1).float()
        den = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(pos * den)
        pe[:, 1::2] = torch.cos(pos * den)
        self.register_buffer("pe", pe.unsqueeze(0), persistent=False)

    def forward(self, x):
        return self.pe[:, :x.size(1)].clone().detach()


def kldiv_loss(log_probs, targets, label_smoothing=0.1, pad_idx=0):
    """speechbrain.nnet.losses.kldiv_loss as the S2S recipe uses it (hparams/S2S/conmambamamba_large.yaml:414-415):
    KL divergence to the label-smoothed target distribution (smoothing mass spread over the other n_class - 1 tokens),
    padding positions masked, summed and divided by the batch size."""
    bz, _, n_class = log_probs.shape
    lp = log_probs.reshape(-1, n_class)
    tg = targets.reshape(-1).long()
    with torch.no_grad():
        dist = torch.full_like(lp, label_smoothing / (n_class - 1))
        ignore = tg == pad_idx
        dist.scatter_(1, tg.masked_fill(ignore, 0).unsqueeze(1), 1.0 - label_smoothing)
    loss = F.kl_div(lp, dist, reduction="none").masked_fill(ignore.unsqueeze(1), 0.0)
    return loss.sum() / bz


class ConMambaS2S(ConMambaCTC):
    """ConMamba encoder + Mamba decoder of ``train_S2S.py`` (compute_forward :344-361) with the objects
    ``hparams/S2S/conmambamamba_large.yaml:300-340`` builds: TransformerASR(encoder_module=conmamba,
    decoder_module=mamba, normalize_before=True) -> (enc_out, dec_out) -> ctc_lin / seq_lin -> log_softmax."""

    def __init__(self, d_model, d_ffn, num_layers, num_decoder_layers=6, output_neurons=5000, dropout=0.1, d_state=16,
                 expand=2, d_conv=4, **kw):
        super().__init__(d_model, d_ffn, num_layers, output_neurons=output_neurons, dropout=dropout, d_state=d_state,
                         expand=expand, d_conv=d_conv, **kw)
        self.custom_tgt_module = NormalizedEmbedding(d_model, output_neurons)
        self.positional_encoding_decoder = PositionalEncoding(d_model)
        mamba_config = dict(d_state=d_state, expand=expand, d_conv=d_conv, bidirectional=True)
        # TransformerASR.py:703-712: decoder activation nn.GELU from the YAML, normalize_before=True
        self.decoder = MambaDecoder(num_layers=num_decoder_layers, d_model=d_model, d_ffn=d_ffn, activation=nn.GELU,
                                    dropout=dropout, normalize_before=True, mamba_config=mamba_config)
        self.seq_lin = BiasGradLinear(d_model, output_neurons)
        for p in list(self.custom_tgt_module.parameters()) + list(self.decoder.parameters()):
            if p.dim() > 1:
                nn.init.xavier_normal_(p)                          # TransformerASR.py:1051-1054

    def forward(self, wavs, tokens_bos, wav_lens=None):
        """wavs (B, n_samples), tokens_bos (B, S) -> (p_ctc (B, L, V), p_seq (B, S, V)) log-probabilities"""
        self._refresh_param_cache()
        self._advance_dropout_seed(wavs.device)
        with precomputed_A(self):
            enc = self.encode(self.features(wavs, wav_lens))
            tgt = self.custom_tgt_module(tokens_bos)
            tgt = tgt + self.positional_encoding_decoder(tgt)
            dec, _, _ = self.decoder(tgt, enc)
        return F.log_softmax(self.ctc_lin(enc), dim=-1), F.log_softmax(self.seq_lin(dec), dim=-1)


def build_model(name, **overrides):
    cfg = dict(CONFIGS[name])
    cfg.update(overrides)
    seed = cfg.pop("seed", None)
    if seed is not None:
        torch.manual_seed(seed)
    if "num_decoder_layers" in cfg:
        return ConMambaS2S(**cfg)
    return ConMambaCTC(**cfg)
