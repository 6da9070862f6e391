"""ConMamba layer API - drop-in for the classes of the reference's ``modules/Conmamba.py`` that sit directly
around the hot path:

  ConvolutionModule      reference modules/Conmamba.py:182-454
  ConmambaEncoderLayer   :457-650   (calls ``self.mamba(x)`` at :642 - the hot path)
  ConmambaEncoder        :653-727
  MambaDecoderLayer      :730-953   (``self_mamba`` :920, ``cross_mamba`` on cat([memory, tgt]) :934)
  MambaDecoder           :956-1031

Same constructor arguments, forward signatures, return tuples and sub-module names (so state_dict keys line up
with reference checkpoints).  The SpeechBrain building blocks the reference imports (``LayerNorm``,
``PositionalwiseFeedForward``, ``Swish``; modules/Conmamba.py:112-121) are restated below in plain torch with
SpeechBrain's parameter names; they stay cuBLAS / cuDNN work and are not part of the hand-written kernel scope
(SURVEY.md section 2.1 row 3).  The Mamba mixers are the fused sm_100a ones from ``bimamba.py``.
"""
import os
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .bimamba import Mamba as BiMamba
from .bimamba import UniMamba as Mamba
from .dwconv import depthwise_conv1d
from .kernels import DWCONV_KSIZES, gelu_dropout_supported
from .layernorm import (FusedLayerNorm, _BiasGradRoute, add_dropout_layer_norm, gelu_dropout, glu, layer_norm_act,
                        layer_norm_gelu_supported)
from .linear import BiasGradLinear, linear as _linear

LAYER_NORM_EPS = 1e-6        # reference ConMambaConstants.LAYER_NORM_EPS (Conmamba.py:687)
FFN_RESIDUAL_SCALE = 0.5     # reference ConMambaConstants.FFN_RESIDUAL_SCALE (Conmamba.py:638,649)


class Swish(nn.Module):
    """speechbrain.nnet.activations.Swish: x * sigmoid(beta * x), beta = 1."""

    def __init__(self, beta: float = 1.0):
        super().__init__()
        self.beta = beta

    def forward(self, x):
        return x * torch.sigmoid(self.beta * x)


class LayerNorm(nn.Module):
    """speechbrain.nnet.normalization.LayerNorm over the last dimension (parameters under ``.norm``)."""

    def __init__(self, input_size, eps=1e-05, elementwise_affine=True, keep_dtype=False):
        super().__init__()
        self.norm = FusedLayerNorm(input_size, eps=eps, elementwise_affine=elementwise_affine, keep_dtype=keep_dtype)

    def forward(self, x):
        return self.norm(x)


class PositionalwiseFeedForward(nn.Module):
    """speechbrain.nnet.attention.PositionalwiseFeedForward: Linear -> act -> Dropout -> Linear (under ``.ffn``)."""

    def __init__(self, d_ffn, input_size, dropout=0.0, activation=nn.ReLU):
        super().__init__()
        self.ffn = nn.Sequential(BiasGradLinear(input_size, d_ffn), activation(), nn.Dropout(dropout),
                                 BiasGradLinear(d_ffn, input_size))

    def forward(self, x, out_bias_grad=True):
        """``out_bias_grad=False``: the caller takes the gradient of the last Linear's bias from the consumer of the output
        (``add_dropout_layer_norm(..., b_bias=self.ffn[3].bias)``)."""
        act, drop = self.ffn[1], self.ffn[2]
        lin2 = self.ffn[3]
        last = lin2 if out_bias_grad else (lambda h: _linear(h, lin2.weight, lin2.bias, bias_grad=False))
        if x.is_cuda and type(act) is nn.GELU and act.approximate == "none" and os.environ.get("CM_NO_FUSE_GELU") is None:
            # Linear -> [GELU + Dropout: one sm_100a kernel] -> Linear; the first Linear's bias gradient is the column sum of
            # the activation's input gradient and is formed inside that backward kernel
            lin1 = self.ffn[0]
            if (type(lin1) is BiasGradLinear and lin1.bias is not None and torch.is_grad_enabled() and lin1.bias.requires_grad
                    and os.environ.get("CM_NO_FUSE_BIAS_GRAD") is None):
                h = _linear(x, lin1.weight, lin1.bias, bias_grad=False)
                if gelu_dropout_supported(h):
                    return last(gelu_dropout(h, drop.p, self.training, bias_for_grad=lin1.bias))
                return last(gelu_dropout(_BiasGradRoute.apply(h, lin1.bias), drop.p, self.training))
            return last(gelu_dropout(lin1(x), drop.p, self.training))
        return last(drop(act(self.ffn[0](x))))


class ConvolutionModule(nn.Module):
    """LayerNorm -> pointwise conv (x2) + GLU -> depthwise conv (k=31) -> LayerNorm -> act -> Linear -> Dropout."""

    def __init__(self, input_size, kernel_size=31, bias=True, activation=Swish, dropout=0.0, causal=False,
                 dilation=1):
        super().__init__()
        self.kernel_size = kernel_size
        self.causal = causal
        self.dilation = dilation
        full = (kernel_size - 1) * 2 ** (dilation - 1)
        self.padding = full if causal else full // 2
        self.layer_norm = FusedLayerNorm(input_size)
        self.bottleneck = nn.Sequential(nn.Conv1d(input_size, 2 * input_size, kernel_size=1, stride=1, bias=bias),
                                        nn.GLU(dim=1))
        self.conv = nn.Conv1d(input_size, input_size, kernel_size=kernel_size, stride=1, padding=self.padding,
                              dilation=dilation, groups=input_size, bias=bias)
        self.after_conv = nn.Sequential(FusedLayerNorm(input_size), activation(), BiasGradLinear(input_size, input_size, bias=bias),
                                        nn.Dropout(dropout))
        # channel-last evaluation on the sm_100a depthwise kernel (no transposes); the CPU reference arm
        # (oracle/cpu_encoder.py) clears the flag and gets the reference's own torch op chain below
        self.use_kernel = dilation == 1 and kernel_size in DWCONV_KSIZES

    def forward(self, x, mask: Optional[torch.Tensor] = None, dynchunktrain_config=None):
        if dynchunktrain_config is not None:
            raise NotImplementedError("Dynamic Chunk Training convolution is never enabled by the ConMamba encoder "
                                      "(TransformerASR.py:783-788 passes no config)")
        if self.use_kernel:
            out = self.body(self.layer_norm(x))
            if mask is not None:
                out.masked_fill_(mask, 0.0)
            return out
        out = self.layer_norm(x).transpose(1, 2)
        out = self.conv(self.bottleneck(out))
        if self.causal:
            out = out[..., :-self.padding]
        out = self.after_conv(out.transpose(1, 2))
        if mask is not None:
            out.masked_fill_(mask, 0.0)
        return out


def _conv_body(self, normed, final_dropout=True, out_bias_grad=True):
    """Everything of the kernel path after ``layer_norm`` (channel-last, no transposes); ``final_dropout=False`` leaves the
    trailing Dropout to the caller (the encoder layer fuses it into the next add + LayerNorm); ``out_bias_grad=False``
    likewise leaves the gradient of the last Linear's bias to that kernel (``b_bias=self.after_conv[2].bias``)."""
    pw = self.bottleneck[0]                                           # pointwise conv = Linear over channel-last rows
    out = glu(_linear(normed, pw.weight.squeeze(-1), pw.bias))      # cm_glu_fwd / cm_glu_bwd
    out = depthwise_conv1d(out, self.conv.weight, self.conv.bias, pad_left=self.padding)
    norm, act = self.after_conv[0], self.after_conv[1]
    # LayerNorm -> GELU as one cm_ln_act kernel each way is opt-in (CM_FUSE_LN_GELU=1): at these narrow rows (144 / 256
    # columns, one warp per row) it measured 16.1 vs 12.9 us backward at 12032 x 144 and 36.6 vs 36.0 us at 32064 x 256
    # against cm_layernorm + torch's GELU (gpurun_out/s4d_step_*.log) - the wide front-end rows are where it pays
    if (os.environ.get("CM_FUSE_LN_GELU") is not None and type(act) is nn.GELU and act.approximate == "none"
            and norm.normalized_shape[0] % 4 == 0):
        out = layer_norm_act(out, norm, "gelu")
    elif (type(act) is nn.GELU and act.approximate == "none" and isinstance(norm, FusedLayerNorm)
          and layer_norm_gelu_supported(out)):
        out = norm(out, gelu=True)            # GELU as the epilogue of cm_layernorm_fwd / _bwd (A/B: CM_NO_LN_GELU_EPILOGUE=1)
    else:
        out = act(norm(out))
    lin = self.after_conv[2]
    out = lin(out) if out_bias_grad else _linear(out, lin.weight, lin.bias, bias_grad=False)
    return self.after_conv[3](out) if final_dropout else out


ConvolutionModule.body = _conv_body


def _make_mixer(d_model, mamba_config, bidirectional_ok):
    """Reference idiom (Conmamba.py:579-591): pop 'bidirectional', build, put the key back."""
    assert mamba_config is not None
    bidirectional = mamba_config.pop('bidirectional')
    try:
        if bidirectional_ok and bidirectional:
            return BiMamba(d_model=d_model, bimamba_type='v2', **mamba_config)
        return Mamba(d_model=d_model, **mamba_config)
    finally:
        mamba_config['bidirectional'] = bidirectional


class ConmambaEncoderLayer(nn.Module):
    def __init__(self, d_model, d_ffn, kernel_size=31, activation=Swish, bias=True, dropout=0.0, causal=False,
                 mamba_config=None):
        super().__init__()
        self.mamba = _make_mixer(d_model, mamba_config, bidirectional_ok=not causal)
        self.convolution_module = ConvolutionModule(d_model, kernel_size, bias, activation, dropout, causal=causal)
        self.ffn_module1 = nn.Sequential(FusedLayerNorm(d_model),
                                         PositionalwiseFeedForward(d_ffn=d_ffn, input_size=d_model, dropout=dropout,
                                                                   activation=activation),
                                         nn.Dropout(dropout))
        self.ffn_module2 = nn.Sequential(FusedLayerNorm(d_model),
                                         PositionalwiseFeedForward(d_ffn=d_ffn, input_size=d_model, dropout=dropout,
                                                                   activation=activation),
                                         nn.Dropout(dropout))
        self.norm1 = LayerNorm(d_model)
        self.norm2 = LayerNorm(d_model, keep_dtype=True)      # its output is the residual stream of the next layer
        self.drop = nn.Dropout(dropout)
        self.fuse_add_norm = os.environ.get("CM_NO_FUSE_ADD_NORM") is None   # evaluation strategy only (A/B switch)

    def forward(self, x, src_mask: Optional[torch.Tensor] = None, src_key_padding_mask: Optional[torch.Tensor] = None,
                pos_embs: torch.Tensor = None, dynchunktrain_config=None):
        # The reference builds a conv mask from src_key_padding_mask and then discards it (Conmamba.py:631-635):
        # padding is never masked inside ConMamba.  Reproduced, not "fixed".
        conv_mask = None
        cm = self.convolution_module
        if x.is_cuda and cm.use_kernel and dynchunktrain_config is None and self.fuse_add_norm:
            # Same arithmetic with every "residual add (+ dropout, + 0.5 scale) -> next LayerNorm" pair evaluated by one
            # sm_100a kernel (cm_add_ln_*); sub-module structure and state_dict are untouched.
            tr = self.training
            # the bias gradients of the Linears that end the three branches are the column sums of the branch gradients the
            # add + LayerNorm backward kernels write: formed there (b_bias=...), not by a pass of their own
            route = torch.is_grad_enabled() and os.environ.get("CM_NO_FUSE_BIAS_GRAD") is None
            bias1 = self.ffn_module1[1].ffn[3].bias if route else None
            bias2 = self.ffn_module2[1].ffn[3].bias if route else None
            biasc = cm.after_conv[2].bias if route else None
            f1 = self.ffn_module1[1](self.ffn_module1[0](x), out_bias_grad=bias1 is None)
            x, h = add_dropout_layer_norm(x, f1, self.norm1.norm, FFN_RESIDUAL_SCALE, self.ffn_module1[2].p, tr, b_bias=bias1)
            x, c_in = add_dropout_layer_norm(x, self.mamba(h), cm.layer_norm, 1.0, 0.0, tr)
            c_out = cm.body(c_in, final_dropout=False, out_bias_grad=biasc is None)
            x, f_in = add_dropout_layer_norm(x, c_out, self.ffn_module2[0], 1.0, cm.after_conv[3].p, tr, b_bias=biasc)
            f2 = self.ffn_module2[1](f_in, out_bias_grad=bias2 is None)
            _, out = add_dropout_layer_norm(x, f2, self.norm2.norm, FFN_RESIDUAL_SCALE, self.ffn_module2[2].p, tr, b_bias=bias2)
            return out
        x = x + FFN_RESIDUAL_SCALE * self.ffn_module1(x)
        skip = x
        x = self.norm1(x)
        x = self.mamba(x)
        x = x + skip
        x = x + self.convolution_module(x, conv_mask, dynchunktrain_config=dynchunktrain_config)
        x = self.norm2(x + FFN_RESIDUAL_SCALE * self.ffn_module2(x))
        return x


class ConmambaEncoder(nn.Module):
    def __init__(self, num_layers, d_model, d_ffn, kernel_size=31, activation=Swish, bias=True, dropout=0.0,
                 causal=False, mamba_config=None):
        super().__init__()
        self.layers = nn.ModuleList([
            ConmambaEncoderLayer(d_model=d_model, d_ffn=d_ffn, dropout=dropout, activation=activation,
                                 kernel_size=kernel_size, bias=bias, causal=causal, mamba_config=mamba_config)
            for _ in range(num_layers)])
        self.norm = LayerNorm(d_model, eps=LAYER_NORM_EPS, keep_dtype=True)

    def forward(self, src, src_mask: Optional[torch.Tensor] = None, src_key_padding_mask: Optional[torch.Tensor] = None,
                pos_embs: Optional[torch.Tensor] = None, dynchunktrain_config=None):
        output = src
        for enc_layer in self.layers:
            output = enc_layer(output, src_mask=src_mask, src_key_padding_mask=src_key_padding_mask,
                               pos_embs=pos_embs, dynchunktrain_config=dynchunktrain_config)
        return self.norm(output), None


class MambaDecoderLayer(nn.Module):
    def __init__(self, d_model, d_ffn, activation=nn.ReLU, dropout=0.0, normalize_before=False, mamba_config=None):
        super().__init__()
        assert mamba_config is not None
        bidirectional = mamba_config.pop('bidirectional')
        self.self_mamba = Mamba(d_model=d_model, **mamba_config)
        self.cross_mamba = Mamba(d_model=d_model, **mamba_config)
        mamba_config['bidirectional'] = bidirectional
        self.pos_ffn = PositionalwiseFeedForward(d_ffn=d_ffn, input_size=d_model, dropout=dropout,
                                                 activation=activation)
        self.norm1 = LayerNorm(d_model, eps=LAYER_NORM_EPS)
        self.norm2 = LayerNorm(d_model, eps=LAYER_NORM_EPS)
        self.norm3 = LayerNorm(d_model, eps=LAYER_NORM_EPS)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)
        self.normalize_before = normalize_before

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, pos_embs_tgt=None, pos_embs_src=None):
        pre = self.normalize_before
        t = self.norm1(tgt) if pre else tgt
        tgt = tgt + self.dropout1(self.self_mamba(t))
        if not pre:
            tgt = self.norm1(tgt)
        t = self.norm2(tgt) if pre else tgt
        # causal scan over [memory ; tgt]: every target position sees the whole encoder output first
        # (Conmamba.py:934); only the target rows go through out_proj
        cross = self.cross_mamba(torch.cat([memory, t], dim=1), keep_last=t.shape[1])
        tgt = tgt + self.dropout2(cross)
        if not pre:
            tgt = self.norm2(tgt)
        t = self.norm3(tgt) if pre else tgt
        tgt = tgt + self.dropout3(self.pos_ffn(t))
        if not pre:
            tgt = self.norm3(tgt)
        return tgt, None, None

    # ---- incremental decoding: the reference re-runs the whole decoder on the growing
    # prefix at every token (TransformerASR.py:822-866), i.e. every layer rescans [memory ; tgt] (Conmamba.py:934).  Both
    # mixers are causal, so the scan over `memory` can be done ONCE per utterance and every further token is a single-token
    # state update (cm_conv_update + cm_ssm_step) from cached states.
    def init_decode(self, memory):
        """Caches for one utterance batch: zero states for self_mamba; for cross_mamba the conv window and SSM state its scan
        over ``memory`` (B, L, d_model) ends in."""
        Bt = memory.shape[0]
        cache = dict(self=self.self_mamba.allocate_inference_cache(Bt, 0), cross=self.cross_mamba.allocate_inference_cache(Bt, 0))
        self.cross_mamba.prefill(memory, *cache["cross"], need_output=False)
        return cache

    @torch.no_grad()
    def decode_step(self, tgt_t, cache):
        """One target position: tgt_t (B, 1, d_model) -> (B, 1, d_model); equals row t of ``forward`` on the prefix tgt[:, :t+1]."""
        pre = self.normalize_before
        tgt = tgt_t
        t = self.norm1(tgt) if pre else tgt
        tgt = tgt + self.self_mamba.step(t, *cache["self"])[0]
        if not pre:
            tgt = self.norm1(tgt)
        t = self.norm2(tgt) if pre else tgt
        tgt = tgt + self.cross_mamba.step(t, *cache["cross"])[0]
        if not pre:
            tgt = self.norm2(tgt)
        t = self.norm3(tgt) if pre else tgt
        tgt = tgt + self.pos_ffn(t)
        if not pre:
            tgt = self.norm3(tgt)
        return tgt


class MambaDecoder(nn.Module):
    def __init__(self, num_layers, d_model, d_ffn, activation=nn.ReLU, dropout=0.0, normalize_before=False,
                 mamba_config=None):
        super().__init__()
        self.layers = nn.ModuleList([
            MambaDecoderLayer(d_model=d_model, d_ffn=d_ffn, activation=activation, dropout=dropout,
                              normalize_before=normalize_before, mamba_config=mamba_config)
            for _ in range(num_layers)])
        self.norm = LayerNorm(d_model, eps=LAYER_NORM_EPS)

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, pos_embs_tgt=None, pos_embs_src=None):
        output = tgt
        for dec_layer in self.layers:
            output, _, _ = dec_layer(output, memory, tgt_mask=tgt_mask, memory_mask=memory_mask,
                                     tgt_key_padding_mask=tgt_key_padding_mask,
                                     memory_key_padding_mask=memory_key_padding_mask,
                                     pos_embs_tgt=pos_embs_tgt, pos_embs_src=pos_embs_src)
        return self.norm(output), [None], [None]

    @torch.no_grad()
    def init_decode(self, memory):
        """Per-layer caches for incremental decoding: ``memory`` is scanned once per layer here, never again."""
        return [layer.init_decode(memory) for layer in self.layers]

    @torch.no_grad()
    def decode_step(self, tgt_t, caches):
        """tgt_t: (B, 1, d_model) embedding (+ positional encoding) of the newest target token -> (B, 1, d_model), the row the
        full forward would produce for it."""
        out = tgt_t
        for layer, cache in zip(self.layers, caches):
            out = layer.decode_step(out, cache)
        return self.norm(out)
