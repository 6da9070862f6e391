"""HyperPyYAML-subset loader (mamba_asr_b200/hparams.py): tag semantics on a self-written file, and - where the reference tree is
present (this container; not the GPU box) - the reference's own hparams files against the model blocks bench.py runs."""
import functools
import os

import pytest
import torch

from mamba_asr_b200 import hparams as H

YAML = """
seed: 11
__set_seed: !apply:torch.manual_seed [!ref <seed>]
experiment: demo
output_folder: !ref results/<experiment>/<seed>
save_folder: !ref <output_folder>/save
data_folder: !PLACEHOLDER
d_model: 64
d_ffn: !ref <d_model> * 4
half: !ref <d_model> // 2
n_fft: 400
n_mels: 80
lr_model: 0.002
max_grad_norm: 3.0
activation: !name:torch.nn.GELU
d_state: 16
mamba_config:
    d_state: !ref <d_state>
    expand: 2
    d_conv: 4
    bidirectional: True
CNN: !new:speechbrain.lobes.models.convolution.ConvolutionFrontEnd
    input_shape: (8, 10, 80)
    num_blocks: 2
    num_layers_per_block: 1
    out_channels: (64, 32)
    kernel_sizes: (3, 3)
    strides: (2, 2)
    residuals: (False, False)
Transformer: !new:modules.TransformerASR.TransformerASR
    input_size: 640
    tgt_vocab: 31
    d_model: !ref <d_model>
    num_encoder_layers: 2
    num_decoder_layers: 0
    d_ffn: !ref <d_ffn>
    dropout: 0.1
    activation: !ref <activation>
    encoder_module: conmamba
    mamba_config: !ref <mamba_config>
ctc_lin: !new:speechbrain.nnet.linear.Linear
    input_size: !ref <d_model>
    n_neurons: 31
log_softmax: !new:torch.nn.LogSoftmax
    dim: -1
modules:
    CNN: !ref <CNN>
    Transformer: !ref <Transformer>
model: !new:torch.nn.ModuleList
    - [!ref <CNN>, !ref <Transformer>, !ref <ctc_lin>]
model_opt_class: !name:torch.optim.AdamW
    lr: !ref <lr_model>
    betas: (0.9, 0.98)
    eps: 0.000000001
    weight_decay: 0.01
noam_annealing: !new:speechbrain.nnet.schedulers.NoamScheduler
    lr_initial: !ref <lr_model>
    n_warmup_steps: 100
augment: !new:speechbrain.augment.time_domain.SpeedPerturb
    orig_freq: 16000
    speeds: [95, 100, 105]
compute_features: !new:speechbrain.lobes.features.Fbank
    sample_rate: 16000
    n_fft: !ref <n_fft>
    n_mels: !ref <n_mels>
"""


def test_tags_refs_arithmetic_and_registry():
    hp = H.load_hparams(YAML, overrides={"data_folder": "/data"})
    assert hp["output_folder"] == "results/demo/11" and hp["save_folder"] == "results/demo/11/save"
    assert hp["data_folder"] == "/data"
    assert hp["d_ffn"] == 256 and hp["half"] == 32
    assert hp["activation"] is torch.nn.GELU
    assert isinstance(hp["log_softmax"], torch.nn.LogSoftmax)
    assert isinstance(hp["ctc_lin"], torch.nn.Linear) and hp["ctc_lin"].in_features == 64
    assert hp["modules"]["Transformer"] is hp["Transformer"]          # a reference is the object, not a copy
    assert hp["CNN"].kwargs["input_shape"] == (8, 10, 80) and hp["CNN"].kwargs["residuals"] == (False, False)
    assert isinstance(hp["model_opt_class"], functools.partial) and hp["model_opt_class"].keywords["betas"] == (0.9, 0.98)
    assert isinstance(hp["augment"], H.Unresolved) and hp["augment"].kwargs["speeds"] == [95, 100, 105]
    with pytest.raises(NotImplementedError):
        hp["augment"]()
    assert abs(hp["noam_annealing"](50) - 0.002 * 10.0 * 50 * 100 ** -1.5) < 1e-12
    assert hp["compute_features"].n_fft == 400 and hp["compute_features"].win_length == 400


def test_placeholder_without_override_and_cycles():
    hp = H.load_hparams(YAML)
    assert isinstance(hp["data_folder"], H.Placeholder)
    with pytest.raises(ValueError):
        H.load_hparams("a: !ref <b>\nb: !ref <a>\n")
    with pytest.raises(KeyError):
        H.load_hparams("a: !ref <nope>\n")


def test_model_and_optimizer_config_from_yaml():
    hp = H.load_hparams(YAML, overrides={"data_folder": "/data"})
    cfg = H.model_config(hp)
    assert cfg == dict(d_model=64, d_ffn=256, num_layers=2, n_fft=400, win_length=25, n_mels=80, output_neurons=31, dropout=0.1,
                       d_state=16, expand=2, d_conv=4, bidirectional=True, seed=11)
    oc = H.optimizer_config(hp)
    assert oc == dict(lr=0.002, weight_decay=0.01, max_grad_norm=3.0, n_warmup_steps=100, betas=(0.9, 0.98), eps=1e-9)
    model = H.build_model_from_hparams(hp)
    from mamba_asr_b200.encoder import ConMambaCTC
    assert isinstance(model, ConMambaCTC) and len(model.encoder.layers) == 2
    bad = YAML.replace("encoder_module: conmamba", "encoder_module: conformer")
    with pytest.raises(NotImplementedError):
        H.model_config(H.load_hparams(bad))


REF = "/root/reference/hparams"


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("yaml_file,config", [("CTC/conmamba_large.yaml", "conmamba_large_ctc"),
                                               ("S2S/conmambamamba_large.yaml", "conmambamamba_large_s2s")])
def test_reference_yaml_selects_the_benchmarked_model(yaml_file, config):
    """The reference's own files resolve to the model blocks `encoder.CONFIGS` restates (and bench.py runs), and to the
    optimizer settings of bench.OPT for the CTC recipe."""
    from mamba_asr_b200.encoder import CONFIGS
    hp = H.load_hparams(open(os.path.join(REF, yaml_file)), overrides={"data_folder": "/tmp/none"})
    cfg = H.model_config(hp)
    for k, v in CONFIGS[config].items():
        assert cfg[k] == v, (k, cfg[k], v)
    assert cfg["d_state"] == 16 and cfg["expand"] == 2 and cfg["d_conv"] == 4 and cfg["dropout"] == 0.1
    if config == "conmamba_large_ctc":
        import bench
        oc = H.optimizer_config(hp)
        for k, v in bench.OPT.items():
            assert oc[k] == v, (k, oc[k], v)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
def test_every_reference_yaml_loads():
    import glob
    files = sorted(glob.glob(os.path.join(REF, "*", "*.yaml")))
    assert len(files) >= 6
    for f in files:
        hp = H.load_hparams(open(f), overrides={"data_folder": "/tmp/none"})
        assert hp["sample_rate"] == 16000 and hp["n_mels"] == 80
