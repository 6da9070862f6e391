"""CPU-only checks (no GPU, no compute calls): the C-ABI library loads and exports every symbol the header declares,
the ctypes structures match the library's struct sizes, host-side helpers agree with the oracle's integer math, the
module layer keeps the reference's state_dict layout, and the product path refuses to run without CUDA."""
import ctypes
import json
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()                          # nvcc cross-compiles sm_100a without a GPU
    from mamba_asr_b200 import _cabi
    return _cabi.lib()


def _declared_entry_points():
    src = open(os.path.join(ROOT, "include", "conmamba_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:int|int64_t)\s+(cm_\w+)\s*\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol(lib):
    from mamba_asr_b200 import _cabi
    names = _declared_entry_points()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libconmamba_b200.so does not export %s" % n
    assert sorted(_cabi.EXPORTS) == names, "ctypes binding and header disagree on the entry points"


def test_ctypes_structs_match_library_layout(lib):
    from mamba_asr_b200 import _cabi
    for i, st in enumerate(_cabi.ABI_STRUCTS):
        assert lib.cm_abi_sizeof(i) == ctypes.sizeof(st), st.__name__
    arch = ctypes.c_int32(0)
    assert lib.cm_version(ctypes.byref(arch)) == _cabi.CM_ABI_VERSION and arch.value == 100


def test_entry_points_validate_arguments_without_a_gpu(lib):
    from mamba_asr_b200 import _cabi
    a = _cabi.ScanFwdArgs()
    assert lib.cm_scan_fwd(ctypes.byref(a), None) == _cabi.CM_ERR_BAD_ARG          # zero sizes / null pointers
    a.batch, a.dim, a.seqlen, a.dstate, a.ndir, a.dtype = 1, 32, 8, 64, 1, _cabi.CM_BF16
    a.out.ptr = 1
    assert lib.cm_scan_fwd(ctypes.byref(a), None) == _cabi.CM_ERR_UNSUPPORTED      # dstate > 16
    c = _cabi.ConvArgs()
    c.batch, c.dim, c.seqlen, c.width, c.ndir, c.dtype = 1, 32, 8, 7, 1, _cabi.CM_F32
    c.x.ptr = 1
    assert lib.cm_conv_fwd(ctypes.byref(c), None) == _cabi.CM_ERR_UNSUPPORTED      # width > 4
    assert lib.cm_reduce_rows(None, 1, 1, None, None) == _cabi.CM_ERR_BAD_ARG


def test_checkpoint_and_partial_counts_follow_the_documented_formulas(lib):
    for L in (1, 7, 8, 9, 37, 251, 376, 501, 7501):
        assert lib.cm_scan_num_ckpt(L, 1) == -(-L // 8)
        m = (L + 1) // 2
        assert lib.cm_scan_num_ckpt(L, 2) == -(-m // 8) + -(-(L - m) // 8)
        assert lib.cm_conv_num_part(3, L) == 3 * -(-L // 64)
    assert [lib.cm_scan_slab_channels(x) for x in (1, 2, 4)] == [32, 16, 8]
    assert lib.cm_scan_pick_lanes(64, 512, 2) == 1 and lib.cm_scan_pick_lanes(4, 512, 2) == 4
    # backward: 32-channel dB/dC slabs (the state-parallel kernel's CTA width) at every shape
    assert lib.cm_scan_pick_lanes_bwd(64, 512, 2) == 1 and lib.cm_scan_pick_lanes_bwd(4, 512, 2) == 1
    assert lib.cm_layernorm_num_part(8) == 1 and lib.cm_layernorm_num_part(10 ** 6) == 148 * 4
    # v17: the LayerNorm backward negotiates its partial rows per row width - the grid of cm_add_ln_bwd's kernels where they
    # apply (cols % 4 == 0: 3 / 2 / 1 CTAs per SM by width), the pair kernels' grid otherwise
    for cols, per_sm in ((144, 3), (256, 3), (512, 2), (1024, 1)):
        assert lib.cm_layernorm_num_part2(10 ** 6, cols) == lib.cm_add_ln_num_part(10 ** 6, cols) == 148 * per_sm
    assert lib.cm_layernorm_num_part2(10 ** 6, 150) == lib.cm_layernorm_num_part(10 ** 6)
    assert lib.cm_layernorm_num_part2(8, 256) == 1


def test_mamba_module_keeps_reference_state_dict_layout(golden_dir):
    from mamba_asr_b200 import Mamba
    meta = json.load(open(os.path.join(golden_dir, "mamba_state_dict.json")))
    m = Mamba(d_model=meta["d_model"], bimamba_type="v2")
    sd = m.state_dict()
    assert list(sd.keys()) == sorted(meta["params"].keys(), key=list(sd.keys()).index)
    assert set(sd.keys()) == set(meta["params"].keys())
    for k, v in sd.items():
        assert list(v.shape) == meta["params"][k]["shape"], k
        assert str(v.dtype) == meta["params"][k]["dtype"], k
    for n, p in m.named_parameters():
        tags = sorted(a for a in ("_no_weight_decay", "_no_reinit") if hasattr(p, a))
        assert tags == meta["tags"][n], n
    assert m.d_inner == meta["d_inner"] and m.dt_rank == meta["dt_rank"]
    z = np.load(os.path.join(golden_dir, "bimamba_v2.npz"))
    m.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p_")}, strict=True)
    with pytest.raises(AssertionError):
        Mamba(d_model=32, bimamba_type="none")                                   # reference bimamba.py:75


def test_product_path_has_no_cpu_fallback(golden_dir):
    from mamba_asr_b200 import Fbank, Mamba, UniMamba
    from mamba_asr_b200.causal_conv1d import causal_conv1d_fn
    from mamba_asr_b200.selective_scan_interface import selective_scan_fn
    with pytest.raises(RuntimeError):
        Mamba(d_model=32, bimamba_type="v2")(torch.randn(1, 8, 32))
    with pytest.raises(RuntimeError):
        UniMamba(d_model=32)(torch.randn(1, 8, 32))
    with pytest.raises(RuntimeError):
        Fbank(n_fft=400, n_mels=80)(torch.randn(1, 1600))
    with pytest.raises(RuntimeError):
        causal_conv1d_fn(torch.randn(1, 32, 8), torch.randn(32, 4))
    with pytest.raises(RuntimeError):
        selective_scan_fn(torch.randn(1, 32, 8), torch.randn(1, 32, 8), -torch.ones(32, 16), torch.randn(1, 16, 8),
                          torch.randn(1, 16, 8))
    # nothing under the package imports the oracle
    pkg = os.path.join(ROOT, "mamba_asr_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+\.*oracle", src, flags=re.M), f


def test_fbank_filterbank_and_frame_math_match_the_oracle():
    from mamba_asr_b200.fbank import Fbank, triangular_filterbank
    from oracle.fbank_ref import mel_filterbank_matrix
    from oracle.lengths_ref import fbank_frames
    for n_fft in (400, 512):
        assert torch.equal(triangular_filterbank(n_fft, 80, 16000, 0, 8000.0), mel_filterbank_matrix(n_fft, 80))
    fb = Fbank(n_fft=512, n_mels=80, win_length=32)
    assert fb.win_length == 512 and fb.hop_length == 160 and fbank_frames(320000, fb.hop_length) == 2001
    with pytest.raises(NotImplementedError):
        Fbank(deltas=True)


def test_layer_api_mirrors_reference_conventions():
    from mamba_asr_b200.bimamba import Mamba as BiMamba, UniMamba
    from mamba_asr_b200.conmamba import ConmambaEncoder, MambaDecoder
    cfg = dict(d_state=16, expand=2, d_conv=4, bidirectional=True)
    enc = ConmambaEncoder(num_layers=2, d_model=32, d_ffn=64, mamba_config=cfg)
    assert cfg == dict(d_state=16, expand=2, d_conv=4, bidirectional=True)     # popped and restored (Conmamba.py:579-591)
    assert isinstance(enc.layers[0].mamba, BiMamba)
    causal = ConmambaEncoder(num_layers=1, d_model=32, d_ffn=64, causal=True, mamba_config=cfg)
    assert isinstance(causal.layers[0].mamba, UniMamba)
    dec = MambaDecoder(num_layers=1, d_model=32, d_ffn=64, mamba_config=cfg)
    assert isinstance(dec.layers[0].self_mamba, UniMamba) and isinstance(dec.layers[0].cross_mamba, UniMamba)
    keys = set(enc.state_dict().keys())
    for k in ("layers.0.mamba.A_b_log", "layers.0.ffn_module1.1.ffn.0.weight", "layers.0.norm1.norm.weight",
              "layers.0.convolution_module.bottleneck.0.weight", "norm.norm.bias"):
        assert k in keys, k


def test_bench_reference_arm_prints_the_contract_line():
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--workload", "tiny"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "audio-sec/sec"


def test_product_length_and_mask_integers_are_bit_exact_with_the_oracle():
    from mamba_asr_b200 import lengths as P
    from oracle import lengths_ref as O
    for n in (16000, 159999, 160000, 240000, 320000, 4800000):
        assert P.fbank_frames(n) == O.fbank_frames(n)
        assert P.encoder_frames(P.fbank_frames(n)) == O.encoder_frames(O.fbank_frames(n))
    g = torch.Generator().manual_seed(0)
    for L in (1, 2, 251, 376, 501, 7501):
        wl = torch.cat([torch.linspace(0.5, 1.0, 64), torch.rand(64, generator=g).clamp(min=0.01), torch.tensor([1.0])])
        a, b = P.abs_lengths(wl, L), O.abs_lengths(wl, L)
        assert torch.equal(a, b) and np.array_equal(a.numpy(), O.abs_lengths_numpy(wl.numpy(), L))
        assert torch.equal(P.key_padding_mask(wl, L), O.key_padding_mask(wl, L))
    assert float(P.abs_lengths(torch.tensor([0.5]), 501)) == 250.0      # half-to-even


def test_kldiv_loss_matches_a_hand_computed_label_smoothed_kl():
    """The S2S recipe's sequence loss (hparams/S2S/conmambamamba_large.yaml:414-415): KL to the label-smoothed target
    distribution, padding masked, sum / batch."""
    import math
    import torch
    from mamba_asr_b200.encoder import kldiv_loss
    torch.manual_seed(0)
    Bt, S, V, eps = 2, 3, 7, 0.1
    logp = torch.log_softmax(torch.randn(Bt, S, V), dim=-1)
    tgt = torch.tensor([[3, 5, 0], [1, 0, 0]])
    total = 0.0
    for b in range(Bt):
        for s_ in range(S):
            t = int(tgt[b, s_])
            if t == 0:
                continue
            for v in range(V):
                q = 1 - eps if v == t else eps / (V - 1)
                total += q * (math.log(q) - float(logp[b, s_, v]))
    assert abs(float(kldiv_loss(logp, tgt, label_smoothing=eps)) - total / Bt) < 1e-5


def test_s2s_model_runs_on_the_cpu_reference_path_and_backpropagates():
    """ConMambaS2S (encoder + Mamba decoder) swapped onto the CPU oracle mixers: shapes, finite loss, a gradient on every
    parameter - the path bench.py --impl reference times for the S2S workload."""
    import torch
    from mamba_asr_b200.encoder import CONFIGS, build_model, kldiv_loss
    from oracle.cpu_encoder import to_cpu_reference
    cfg = CONFIGS["conmambamamba_large_s2s"]
    m = build_model("conmambamamba_large_s2s", d_model=32, d_ffn=64, num_layers=1, num_decoder_layers=1,
                    output_neurons=20, dropout=0.0)
    m = to_cpu_reference(m, cfg["n_fft"], cfg["n_mels"], cfg["win_length"])
    wav = 0.1 * torch.randn(2, 6400)
    bos = torch.tensor([[1, 4, 5, 6], [1, 7, 8, 9]])
    p_ctc, p_seq = m(wav, bos)
    assert p_ctc.shape == (2, 11, 20) and p_seq.shape == (2, 4, 20)
    loss = kldiv_loss(p_seq, torch.tensor([[4, 5, 6, 2], [7, 8, 9, 2]])) + p_ctc[..., 0].mean()
    loss.backward()
    assert torch.isfinite(loss)
    missing = [n for n, p in m.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing


def test_decode_api_mirrors_the_reference_and_refuses_cpu():
    """step / allocate_inference_cache / _get_states_from_cache exist with the reference's signatures
    (modules/mamba/bimamba.py:320, 367, 381) and the product path has no CPU fallback for them."""
    import inspect
    import pytest
    import torch
    from mamba_asr_b200 import Mamba, UniMamba
    for cls in (Mamba, UniMamba):
        assert list(inspect.signature(cls.step).parameters)[1:] == ["hidden_states", "conv_state", "ssm_state"]
        assert list(inspect.signature(cls.allocate_inference_cache).parameters)[1:4] == ["batch_size", "max_seqlen", "dtype"]
    m = UniMamba(d_model=16, layer_idx=2)
    cs, ss = m.allocate_inference_cache(3, 0)
    assert cs.shape == (3, 32, 4) and ss.shape == (3, 32, 16) and cs.dtype == torch.float32

    class IP:
        seqlen_offset = 0
        key_value_memory_dict = {}
    ip = IP()
    a, b = m._get_states_from_cache(ip, 3)
    assert ip.key_value_memory_dict[2][0] is a and a.shape == cs.shape
    with pytest.raises(RuntimeError):
        m.step(torch.zeros(3, 1, 16), cs, ss)
    with pytest.raises(AssertionError):
        UniMamba(d_model=16)._get_states_from_cache(IP(), 1)       # layer_idx is required, as in the reference


def test_batched_neg_exp_matches_the_per_block_evaluation():
    """precomputed_A evaluates A = -exp(A_log) (reference bimamba.py:200,222) for all blocks with multi-tensor kernels;
    values and gradients must equal the per-block expression, unused outputs get no gradient."""
    import torch
    from mamba_asr_b200.bimamba import _NegExpMany
    torch.manual_seed(0)
    logs = [torch.randn(6, 16, requires_grad=True) for _ in range(3)]
    outs = _NegExpMany.apply(*logs)
    (outs[0].sum() * 2 + outs[2].pow(2).sum()).backward()
    refs = [x.detach().clone().requires_grad_(True) for x in logs]
    ro = [-torch.exp(x) for x in refs]
    (ro[0].sum() * 2 + ro[2].pow(2).sum()).backward()
    for o, r in zip(outs, ro):
        assert torch.equal(o, r)
    assert torch.allclose(logs[0].grad, refs[0].grad) and torch.allclose(logs[2].grad, refs[2].grad)
    assert logs[1].grad is None


def test_ln_act_entry_points_validate_arguments_without_a_gpu(lib):
    """cm_ln_act_* (include/conmamba_b200.h): null / empty -> BAD_ARG; rows outside the envelope (not a multiple of 4,
    > 2560 columns, a pre-norm bias whose period does not divide the row, unaligned base) -> UNSUPPORTED; partial-row
    count = one wave of the backward kernel, at least 1."""
    from mamba_asr_b200 import _cabi
    a = _cabi.LnActArgs()
    assert lib.cm_ln_act_fwd(ctypes.byref(a), None) == _cabi.CM_ERR_BAD_ARG
    assert lib.cm_ln_act_bwd(ctypes.byref(a), None) == _cabi.CM_ERR_BAD_ARG
    a.rows, a.cols, a.dtype, a.act = 4, 640, _cabi.CM_BF16, _cabi.CM_LN_ACT_LEAKY_RELU
    a.x = a.y = a.gamma = a.beta = a.mean = a.rstd = 4096        # non-null, 16-byte aligned (never dereferenced here)
    a.act = 7
    assert lib.cm_ln_act_fwd(ctypes.byref(a), None) == _cabi.CM_ERR_BAD_ARG          # unknown activation
    a.act = _cabi.CM_LN_ACT_GELU
    for bad_cols in (6, 2564, 4096):
        a.cols = bad_cols
        assert lib.cm_ln_act_fwd(ctypes.byref(a), None) == _cabi.CM_ERR_UNSUPPORTED
    a.cols = 640
    a.pre_bias, a.pre_bias_n = 4096, 48                            # 640 % 48 != 0
    assert lib.cm_ln_act_fwd(ctypes.byref(a), None) == _cabi.CM_ERR_UNSUPPORTED
    a.pre_bias_n = 6                                               # not a multiple of 4
    assert lib.cm_ln_act_fwd(ctypes.byref(a), None) == _cabi.CM_ERR_UNSUPPORTED
    a.pre_bias, a.pre_bias_n = None, 0
    a.x = 4098                                                     # unaligned base
    assert lib.cm_ln_act_fwd(ctypes.byref(a), None) == _cabi.CM_ERR_UNSUPPORTED
    a.x = 4096
    assert lib.cm_ln_act_bwd(ctypes.byref(a), None) == _cabi.CM_ERR_BAD_ARG          # dy / dx / partials missing
    assert lib.cm_ln_act_num_part(0, 640) == 1
    assert lib.cm_ln_act_num_part(10, 640) == 3                    # 4 rows per CTA at <= 640 columns
    assert lib.cm_ln_act_num_part(10, 2560) == 10                  # one row per CTA at 2560 columns
    assert lib.cm_ln_act_num_part(10 ** 6, 2560) == 148 * 4


def test_split_k_weight_gradient_row_blocks(monkeypatch):
    """linear._wgrad_nsplit: the measured choices at the BASELINE shapes (profiles/r01_wgrad_variants_session4.txt), row
    counts that do not divide fall back to a smaller power of two, short inputs and the A/B switch to one GEMM."""
    from mamba_asr_b200.linear import _wgrad_nsplit
    monkeypatch.delenv("CM_NO_WGRAD_SPLIT", raising=False)
    assert _wgrad_nsplit(12032, 1024, 144) == 4 and _wgrad_nsplit(12032, 144, 1024) == 4      # small FFN
    assert _wgrad_nsplit(12032, 576, 144) == 8 and _wgrad_nsplit(12032, 144, 144) == 8
    assert _wgrad_nsplit(32064, 1024, 256) == 8 and _wgrad_nsplit(32064, 256, 1024) == 8      # large FFN
    assert _wgrad_nsplit(32064, 256, 512) == 16 and _wgrad_nsplit(32064, 256, 256) == 16
    assert _wgrad_nsplit(5010, 64, 48) == 2                        # 5010 = 2 * 2505
    assert _wgrad_nsplit(5011, 64, 48) == 1 and _wgrad_nsplit(1503, 256, 1024) == 1
    monkeypatch.setenv("CM_NO_WGRAD_SPLIT", "1")
    assert _wgrad_nsplit(12032, 1024, 144) == 1


def test_reference_cpu_names_are_importable_but_refuse_to_run():
    """B2 namespace parity (reference selective_scan_interface.py:91, :641, :678): the *_ref names import through the
    package and through compat/modules, and raise - the product has no CPU path and never imports oracle/."""
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    try:
        from modules.mamba.selective_scan_interface import (bimamba_inner_ref, mamba_inner_ref,  # noqa: F401
                                                            selective_scan_fn, selective_scan_ref)
    finally:
        sys.path.remove(os.path.join(ROOT, "compat"))
    for fn in (selective_scan_ref, mamba_inner_ref, bimamba_inner_ref):
        with pytest.raises(NotImplementedError, match="oracle"):
            fn(torch.randn(1, 4, 8), torch.randn(1, 4, 8), -torch.ones(4, 16), torch.randn(1, 16, 8), torch.randn(1, 16, 8))


def test_param_cache_never_hands_out_a_stale_copy():
    """linear.ParamCache: a copy taken before an in-place parameter update (optimizer step, load_state_dict) must not be
    returned afterwards (eval after a training step silently used stale bf16 weights)."""
    from mamba_asr_b200.linear import ParamCache

    class _P(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.randn(4, 8))

    m = _P()
    c = ParamCache.__new__(ParamCache)                      # the constructor filters on .is_cuda: build the CPU equivalent
    c.params, c.dtype = [m.w], torch.bfloat16
    c.flat = torch.zeros(32, dtype=torch.bfloat16)
    c.views = [c.flat[:32].view(4, 8)]
    c.index, c.versions = {id(m.w): 0}, [-1]
    assert c.get(m.w, torch.bfloat16) is None               # never refreshed: no copy yet
    c.refresh()
    assert torch.equal(c.get(m.w, torch.bfloat16), m.w.detach().bfloat16())
    with torch.no_grad():
        m.w.add_(1.0)                                       # what optimizer.step() does
    assert c.get(m.w, torch.bfloat16) is None
    c.refresh()
    assert torch.equal(c.get(m.w, torch.bfloat16), m.w.detach().bfloat16())
    assert c.get(m.w, torch.float16) is None


def test_reduce_batch_and_glu_entry_points_validate_arguments_without_a_gpu(lib):
    """cm_reduce_batch / cm_glu_fwd / cm_glu_bwd (include/conmamba_b200.h): null or empty -> BAD_ARG, geometry outside the
    envelope -> UNSUPPORTED, before any launch."""
    from mamba_asr_b200 import _cabi
    arr = (_cabi.ReduceJob2 * 2)()
    assert lib.cm_reduce_batch(arr, 2, None) == _cabi.CM_ERR_BAD_ARG                  # null pointers
    assert lib.cm_reduce_batch(None, 1, None) == _cabi.CM_ERR_BAD_ARG
    arr[0].part, arr[0].out, arr[0].rows, arr[0].cols, arr[0].stride = 256, 512, 4, 8, 7
    assert lib.cm_reduce_batch(arr, 1, None) == _cabi.CM_ERR_BAD_ARG                  # stride < cols
    assert lib.cm_reduce_batch(arr, _cabi.CM_REDUCE_BATCH_MAX + 1, None) == _cabi.CM_ERR_BAD_ARG
    assert lib.cm_glu_fwd(None, 256, 4, 64, 128, 64, _cabi.CM_BF16, None) == _cabi.CM_ERR_BAD_ARG
    assert lib.cm_glu_fwd(256, 512, 0, 64, 128, 64, _cabi.CM_BF16, None) == _cabi.CM_ERR_BAD_ARG
    assert lib.cm_glu_fwd(256, 512, 4, 60, 120, 60, _cabi.CM_BF16, None) == _cabi.CM_ERR_UNSUPPORTED    # dim % 8
    assert lib.cm_glu_fwd(258, 512, 4, 64, 128, 64, _cabi.CM_BF16, None) == _cabi.CM_ERR_UNSUPPORTED    # alignment
    assert lib.cm_glu_fwd(256, 512, 4, 64, 64, 64, _cabi.CM_BF16, None) == _cabi.CM_ERR_UNSUPPORTED     # h rows narrower than 2 dim
    assert lib.cm_glu_bwd(256, None, 512, 4, 64, 128, 64, 128, _cabi.CM_F32, None) == _cabi.CM_ERR_BAD_ARG
    assert lib.cm_glu_bwd(256, 512, 768, 4, 64, 128, 64, 64, _cabi.CM_F32, None) == _cabi.CM_ERR_BAD_ARG  # dh rows narrower than 2 dim


def test_deferred_reduction_queue_host_logic():
    """kernels.deferred_reductions / reduce_many bookkeeping that needs no GPU: job geometry of both tuple forms, the context
    manager restores the previous mode and leaves nothing queued, grad_cast returns the tensor itself when no cast is needed
    (a queued gradient must not be read), and a deferred call outside a backward pass is not queued (it would never run)."""
    import torch
    from mamba_asr_b200 import kernels as K
    part = torch.zeros(6, 4, 10)
    out = torch.zeros(4, 10)
    j = K._norm_job((part, out))
    assert j[4:] == (6, 40, 40) and j[2] == part.data_ptr() and j[3] == out.data_ptr()
    j = K._norm_job((part, out.view(-1)[8:], 6, 16, 40, 24))
    assert j[4:] == (6, 16, 40) and j[2] == part.data_ptr() + 4 * 24 and j[3] == out.data_ptr() + 4 * 8
    assert K._DEFER is False
    with K.deferred_reductions():
        assert K._DEFER is True
        with K.deferred_reductions(False):
            assert K._DEFER is False
        assert K._DEFER is True
    assert K._DEFER is False and not K._PENDING
    g = torch.zeros(3)
    assert K.grad_cast(g, torch.float32) is g and K.grad_cast(None, torch.float32) is None
    assert K.grad_cast(g, torch.float64).dtype == torch.float64
    K.reduce_many([])                                            # nothing to do, no library call


def test_layernorm_backward_argument_validation(lib):
    """cm_layernorm_bwd (include/conmamba_b200.h, ABI v17): missing pointers, empty shapes and a negative partial-row count
    are BAD_ARG, more than 1024 columns and an unknown epilogue are refused - all before any launch (no GPU needed)."""
    from mamba_asr_b200 import _cabi
    a = _cabi.LayerNormArgs()
    assert lib.cm_layernorm_bwd(ctypes.byref(a), None) == _cabi.CM_ERR_BAD_ARG
    for f in ("x", "dy", "dx", "mean", "rstd", "dgamma_part", "dbeta_part"):
        setattr(a, f, 256)
    a.rows, a.cols = 16, 256
    a.x_dtype = a.y_dtype = _cabi.CM_BF16
    a.x_stride = a.dy_stride = a.dx_stride = 256
    a.n_part = -1
    assert lib.cm_layernorm_bwd(ctypes.byref(a), None) == _cabi.CM_ERR_BAD_ARG
    a.n_part = 0
    a.cols = 2048
    assert lib.cm_layernorm_bwd(ctypes.byref(a), None) == _cabi.CM_ERR_UNSUPPORTED
    a.cols, a.act = 256, 7
    assert lib.cm_layernorm_bwd(ctypes.byref(a), None) == _cabi.CM_ERR_BAD_ARG
