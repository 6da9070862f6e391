"""CPU-only checks (no GPU, no compute calls): the C-ABI library loads and exports every symbol the header declares,
the ctypes structures match the library's struct sizes, host-side helpers agree with the oracle's integer math, the
module layer keeps the reference's state_dict layout, and the product path refuses to run without CUDA."""
import ctypes
import json
import os
import re

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
