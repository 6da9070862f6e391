"""GPU parity of the incremental-decoding path: cm_ssm_step / cm_conv_update behind
``selective_state_update``, ``Mamba.step``, ``allocate_inference_cache`` and ``forward(inference_params=...)``
(reference modules/mamba/bimamba.py:176-186, 320-414)."""
import os

import numpy as np
import pytest
import torch

from cm_testutil import assert_close

pytestmark = pytest.mark.gpu


class InferenceParams:          # the two fields of mamba_ssm.utils.generation.InferenceParams the modules read
    def __init__(self):
        self.seqlen_offset = 0
        self.key_value_memory_dict = {}


def _load(golden_dir):
    z = np.load(os.path.join(golden_dir, "mamba_step.npz"))
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p_")}
    return z, sd


@pytest.mark.parametrize("dtype,state_dtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32),
                                               (torch.bfloat16, torch.bfloat16), (torch.float16, torch.float32)])
@pytest.mark.parametrize("shape", [(3, 64, 16), (2, 33, 16), (1, 1024, 16), (4, 96, 8), (2, 40, 5), (65, 32, 16)])
@pytest.mark.parametrize("full", [True, False])
def test_selective_state_update_matches_oracle(dtype, state_dtype, shape, full):
    from mamba_asr_b200.selective_state_update import selective_state_update
    from oracle.step_ref import ssm_step_oracle
    Bt, D, N = shape
    g = torch.Generator().manual_seed(Bt * 1000 + D + N)
    rn = lambda *s: torch.randn(*s, generator=g)
    state = rn(Bt, D, N).to(state_dtype)
    A = -torch.exp(0.5 * rn(D, N))
    Dp = rn(D) if full else None
    bias = (rn(D) - 3.0) if full else None
    sg = state.clone().cuda()
    sref = state.float().clone()
    for it in range(3):
        x, dt, zz = rn(Bt, D).to(dtype), (0.5 * rn(Bt, D)).to(dtype), rn(Bt, D).to(dtype)
        if not full:
            dt = dt.abs()
        Bm, Cm = rn(Bt, N).to(dtype), rn(Bt, N).to(dtype)
        ref = ssm_step_oracle(sref, x.float(), dt.float(), A, Bm.float(), Cm.float(), Dp, zz.float() if full else None,
                              bias, dt_softplus=full)
        out = selective_state_update(sg, x.cuda(), dt.cuda(), A.cuda(), Bm.cuda(), Cm.cuda(),
                                     None if Dp is None else Dp.cuda(), z=zz.cuda() if full else None,
                                     dt_bias=None if bias is None else bias.cuda(), dt_softplus=full)
        assert out.dtype == dtype and sg.dtype == state_dtype
        assert_close(out.float(), ref, dtype, floor="max", what="step %d out" % it)
        assert_close(sg.float(), sref, state_dtype, floor="max", what="step %d state" % it)
        if state_dtype != torch.float32:
            sref = sg.float().cpu().clone()        # follow the rounded trajectory


def test_mamba_step_from_zero_cache_matches_reference_golden(golden_dir):
    from mamba_asr_b200 import Mamba
    z, sd = _load(golden_dir)
    m = Mamba(d_model=sd["in_proj.weight"].shape[1], bimamba_type="v2", layer_idx=0).cuda()
    m.load_state_dict(sd, strict=True)
    tokens = torch.from_numpy(z["tokens"]).cuda()
    cs, ss = m.allocate_inference_cache(tokens.shape[0], 0)
    assert cs.shape == (tokens.shape[0], m.d_inner, m.d_conv) and ss.shape == (tokens.shape[0], m.d_inner, m.d_state)
    outs = []
    for t in range(tokens.shape[1]):
        o, cs, ss = m.step(tokens[:, t:t + 1], cs, ss)
        outs.append(o)
    assert_close(torch.cat(outs, 1), torch.from_numpy(z["out_steps_zero"]), what="step outputs", rtol_mul=2.0)
    assert_close(cs, torch.from_numpy(z["conv_zero_final"]), what="conv state")
    assert_close(ss, torch.from_numpy(z["ssm_zero_final"]), floor="max", what="ssm state")


def test_prefill_then_step_matches_reference_golden(golden_dir):
    from mamba_asr_b200 import Mamba
    z, sd = _load(golden_dir)
    m = Mamba(d_model=sd["in_proj.weight"].shape[1], bimamba_type="v2", layer_idx=0).cuda()
    m.load_state_dict(sd, strict=True)
    prompt, tokens = torch.from_numpy(z["prompt"]).cuda(), torch.from_numpy(z["tokens"]).cuda()
    ip = InferenceParams()
    out = m(prompt, inference_params=ip)
    assert_close(out, torch.from_numpy(z["out_prefill"]), what="prefill out", rtol_mul=2.0)
    cs, ss = ip.key_value_memory_dict[0]
    assert_close(cs, torch.from_numpy(z["conv_after_prefill"]), what="conv state after prefill")
    assert_close(ss, torch.from_numpy(z["ssm_after_prefill"]), floor="max", what="ssm state after prefill")
    outs = []
    for t in range(tokens.shape[1]):
        ip.seqlen_offset = prompt.shape[1] + t
        outs.append(m(tokens[:, t:t + 1], inference_params=ip))
    assert_close(torch.cat(outs, 1), torch.from_numpy(z["out_steps"]), what="step outputs", rtol_mul=2.0)
    assert_close(ip.key_value_memory_dict[0][1], torch.from_numpy(z["ssm_final"]), floor="max", what="final ssm state")


@pytest.mark.parametrize("autocast", [False, True])
def test_unimamba_decoding_equals_full_sequence_forward(autocast):
    """Size-independent property at the decoder's width (cfg 4: d_model 512): prefill of k tokens + single-token steps
    reproduce the training-path forward over the whole sequence; a short prompt (L < d_conv) pads the conv window."""
    from mamba_asr_b200 import UniMamba
    torch.manual_seed(9)
    m = UniMamba(d_model=512, layer_idx=3).cuda()
    for p in m.parameters():
        if p.dim() > 1:
            torch.nn.init.xavier_normal_(p)
    x = torch.randn(4, 40, 512, device="cuda")
    dt = torch.bfloat16 if autocast else torch.float32
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        full = m(x)
        for k in (2, 17):
            ip = InferenceParams()
            outs = [m(x[:, :k], inference_params=ip)]
            for t in range(k, x.shape[1]):
                ip.seqlen_offset = t
                outs.append(m(x[:, t:t + 1], inference_params=ip))
            got = torch.cat(outs, 1)
            assert got.shape == full.shape
            assert_close(got.float(), full.float(), dt, floor="max", what="prefill %d + steps" % k,
                         rtol_mul=2.0 if not autocast else 1.0)


def test_step_has_no_cpu_path_and_validates():
    from mamba_asr_b200 import UniMamba
    from mamba_asr_b200.selective_state_update import selective_state_update
    m = UniMamba(d_model=32, layer_idx=0)
    with pytest.raises(RuntimeError):
        m.step(torch.zeros(1, 1, 32), torch.zeros(1, 64, 4), torch.zeros(1, 64, 16))
    with pytest.raises(ValueError):
        selective_state_update(torch.zeros(2, 8, 16, device="cuda"), torch.zeros(2, 8, device="cuda"),
                               torch.zeros(2, 8, device="cuda"), torch.zeros(8, 16, device="cuda"),
                               torch.zeros(2, 4, device="cuda"), torch.zeros(2, 16, device="cuda"))


@pytest.mark.parametrize("autocast", [False, True])
def test_decoder_incremental_decoding_scans_memory_once(autocast):
    """MambaDecoder.init_decode + decode_step (the cross-Mamba scan over `memory` done once per utterance, every target token a
    single-token state update) reproduce the full decoder forward on the same (memory, tgt) - which is what the reference
    recomputes from scratch for every new token (TransformerASR.py:822-866, Conmamba.py:934).  Decoder of BASELINE config 4
    (d_model 512, 2 of its 6 layers), memory of 67 frames, 11 target positions."""
    import torch.nn as nn
    from mamba_asr_b200.conmamba import MambaDecoder
    torch.manual_seed(12)
    cfg = dict(d_state=16, expand=2, d_conv=4, bidirectional=True)
    dec = MambaDecoder(num_layers=2, d_model=512, d_ffn=2048, activation=nn.GELU, dropout=0.1, normalize_before=True,
                       mamba_config=cfg).cuda().eval()
    for p in dec.parameters():
        if p.dim() > 1:
            nn.init.xavier_normal_(p)
    memory = torch.randn(3, 67, 512, device="cuda")
    tgt = torch.randn(3, 11, 512, device="cuda")
    dt = torch.bfloat16 if autocast else torch.float32
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        full = dec(tgt, memory)[0]
        caches = dec.init_decode(memory)
        rows = [dec.decode_step(tgt[:, t:t + 1], caches) for t in range(tgt.shape[1])]
    got = torch.cat(rows, 1)
    assert got.shape == full.shape
    assert_close(got.float(), full.float(), dt, floor="max", what="incremental decoder", rtol_mul=2.0 if not autocast else 1.0)
    # the caches advanced: a second utterance needs fresh ones
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        again = dec.decode_step(tgt[:, :1], dec.init_decode(memory))
    assert_close(again.float(), full[:, :1].float(), dt, floor="max", what="fresh caches", rtol_mul=2.0 if not autocast else 1.0)
