"""B200-native ``Mamba`` mixers with the reference's constructor signatures and state_dict layout.

  ``Mamba``     - drop-in for ``modules.mamba.bimamba.Mamba`` (reference modules/mamba/bimamba.py:39-318), the
                  bidirectional "v2" block used by ``ConmambaEncoderLayer`` (modules/Conmamba.py:586-590).
  ``UniMamba``  - drop-in for ``mamba_ssm.Mamba`` (mamba-ssm 1.1.3.post1), the causal mixer used by the decoder
                  and by causal encoders (modules/Conmamba.py:581-584, :854-862).

Parameter names, shapes, dtypes, creation order and the ``_no_weight_decay`` / ``_no_reinit`` tags match the
reference (pinned by tests/golden/mamba_state_dict.json), so checkpoints load either way.

Forward is one fused path: in_proj GEMM -> ``MambaInnerCL`` (fused bidirectional conv + scan kernels,
channel-last, no flips) -> out_proj GEMM.  There is no slow path and no CPU path.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .mamba_inner import MambaInnerCL


def _s4d_real_log(d_inner, d_state, device):
    # reference bimamba.py:123-129: A = repeat(arange(1, N+1), "n -> d n"); A_log = log(A), kept in fp32
    a = torch.arange(1, d_state + 1, dtype=torch.float32, device=device)
    return torch.log(a).unsqueeze(0).expand(d_inner, d_state).contiguous()


class _MambaBase(nn.Module):
    def _build(self, d_model, d_state, d_conv, expand, dt_rank, dt_min, dt_max, dt_init, dt_scale, dt_init_floor,
               conv_bias, bias, use_fast_path, layer_idx, device, dtype, bidirectional):
        fk = {"device": device, "dtype": dtype}
        self.d_model = d_model
        self.d_state = d_state
        self.d_conv = d_conv
        self.expand = expand
        self.d_inner = int(self.expand * self.d_model)                                  # bimamba.py:68
        self.dt_rank = math.ceil(self.d_model / 16) if dt_rank == "auto" else dt_rank    # bimamba.py:69
        self.use_fast_path = use_fast_path
        self.layer_idx = layer_idx
        D, R, N = self.d_inner, self.dt_rank, self.d_state

        self.in_proj = nn.Linear(d_model, 2 * D, bias=bias, **fk)
        self.conv1d = nn.Conv1d(D, D, kernel_size=d_conv, groups=D, padding=d_conv - 1, bias=conv_bias, **fk)
        self.activation = "silu"
        self.act = nn.SiLU()
        self.x_proj = nn.Linear(D, R + 2 * N, bias=False, **fk)
        self.dt_proj = nn.Linear(R, D, bias=True, **fk)

        # dt projection init (bimamba.py:101-120): weight scale R^-0.5, bias = softplus^-1(dt), dt log-uniform
        std = R ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(self.dt_proj.weight, std)
        elif dt_init == "random":
            nn.init.uniform_(self.dt_proj.weight, -std, std)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(D, **fk) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min))
        dt = dt.clamp(min=dt_init_floor)
        with torch.no_grad():
            self.dt_proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))
        self.dt_proj.bias._no_reinit = True

        self.A_log = nn.Parameter(_s4d_real_log(D, N, device))
        self.A_log._no_weight_decay = True
        self.D = nn.Parameter(torch.ones(D, device=device))
        self.D._no_weight_decay = True

        if bidirectional:
            # second direction: its own A, conv, x_proj, dt_proj, D (bimamba.py:146-172).  As in the reference,
            # dt_proj_b keeps the default nn.Linear init.
            self.A_b_log = nn.Parameter(_s4d_real_log(D, N, device))
            self.A_b_log._no_weight_decay = True
            self.conv1d_b = nn.Conv1d(D, D, kernel_size=d_conv, groups=D, padding=d_conv - 1, bias=conv_bias, **fk)
            self.x_proj_b = nn.Linear(D, R + 2 * N, bias=False, **fk)
            self.dt_proj_b = nn.Linear(R, D, bias=True, **fk)
            self.D_b = nn.Parameter(torch.ones(D, device=device))
            self.D_b._no_weight_decay = True

        self.out_proj = nn.Linear(D, d_model, bias=bias, **fk)

    def _dir_params(self, suffix=""):
        g = lambda n: getattr(self, n + suffix)
        A_log = self.A_b_log if suffix else self.A_log
        Dk = self.D_b if suffix else self.D
        A = -torch.exp(A_log.float())                                                   # bimamba.py:200,222
        return (g("conv1d").weight, g("conv1d").bias, g("x_proj").weight, g("dt_proj").weight, A, Dk.float(),
                g("dt_proj").bias.float())

    def _check(self, hidden_states, inference_params):
        if inference_params is not None:
            raise NotImplementedError(
                "incremental decoding (inference_params / step) is not part of the ConMamba training or "
                "evaluation path: the reference trainers never pass it (SURVEY.md section 2.1 row 2)")
        if not hidden_states.is_cuda:
            raise RuntimeError("mamba_asr_b200.Mamba runs on CUDA only (sm_100a kernels, no CPU fallback)")

    def step(self, hidden_states, conv_state, ssm_state):
        raise NotImplementedError("single-token step is outside the ConMamba hot path (never called by the trainers)")

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        raise NotImplementedError("inference cache is outside the ConMamba hot path (never called by the trainers)")


class Mamba(_MambaBase):
    """Bidirectional (v2) Mamba mixer; signature of reference modules/mamba/bimamba.py:40-61."""

    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False,
                 use_fast_path=True, layer_idx=None, device=None, dtype=None, bimamba_type="none",
                 if_devide_out=True, init_layer_scale=None):
        super().__init__()
        assert bimamba_type == 'v2'                                                      # bimamba.py:75
        self.bimamba_type = bimamba_type
        self.if_devide_out = if_devide_out
        self.init_layer_scale = init_layer_scale
        if init_layer_scale is not None:
            self.gamma = nn.Parameter(init_layer_scale * torch.ones((d_model)), requires_grad=True)
        self._build(d_model, d_state, d_conv, expand, dt_rank, dt_min, dt_max, dt_init, dt_scale, dt_init_floor,
                    conv_bias, bias, use_fast_path, layer_idx, device, dtype, bidirectional=True)

    def forward(self, hidden_states, inference_params=None):
        """hidden_states: (B, L, d_model) -> (B, L, d_model)"""
        self._check(hidden_states, inference_params)
        xz = F.linear(hidden_states, self.in_proj.weight, self.in_proj.bias)              # (B, L, 2D), time-major
        scale = 0.5 if self.if_devide_out else 1.0                                        # bimamba.py:250-253
        y = MambaInnerCL.apply(xz, 2, scale, False, *self._dir_params(""), *self._dir_params("_b"))
        out = F.linear(y, self.out_proj.weight, self.out_proj.bias)
        if self.init_layer_scale is not None:
            out = out * self.gamma
        return out


class UniMamba(_MambaBase):
    """Causal Mamba mixer; signature of ``mamba_ssm.Mamba`` (imported at reference modules/Conmamba.py:124)."""

    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False,
                 use_fast_path=True, layer_idx=None, device=None, dtype=None):
        super().__init__()
        self._build(d_model, d_state, d_conv, expand, dt_rank, dt_min, dt_max, dt_init, dt_scale, dt_init_floor,
                    conv_bias, bias, use_fast_path, layer_idx, device, dtype, bidirectional=False)

    def forward(self, hidden_states, inference_params=None):
        self._check(hidden_states, inference_params)
        xz = F.linear(hidden_states, self.in_proj.weight, self.in_proj.bias)
        y = MambaInnerCL.apply(xz, 1, 1.0, False, *self._dir_params(""))
        return F.linear(y, self.out_proj.weight, self.out_proj.bias)
