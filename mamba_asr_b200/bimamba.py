"""B200-native ``Mamba`` mixers with the reference's constructor signatures and state_dict layout.

  ``Mamba``     - drop-in for ``modules.mamba.bimamba.Mamba`` (reference modules/mamba/bimamba.py:39-318), the
                  bidirectional "v2" block used by ``ConmambaEncoderLayer`` (modules/Conmamba.py:586-590).
  ``UniMamba``  - drop-in for ``mamba_ssm.Mamba`` (mamba-ssm 1.1.3.post1), the causal mixer used by the decoder
                  and by causal encoders (modules/Conmamba.py:581-584, :854-862).

Parameter names, shapes, dtypes, creation order and the ``_no_weight_decay`` / ``_no_reinit`` tags match the
reference (pinned by tests/golden/mamba_state_dict.json), so checkpoints load either way.

Forward is one fused path: in_proj GEMM -> ``MambaInnerCL`` (fused bidirectional conv + scan kernels,
channel-last, no flips) -> out_proj GEMM.  Incremental decoding (``inference_params``, ``step``,
``allocate_inference_cache``; reference bimamba.py:320-414) runs on the single-token kernels cm_conv_update and
cm_ssm_step.  There is no slow path and no CPU path.
"""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kernels as K
from .causal_conv1d import causal_conv1d_update
from .linear import linear
from .mamba_inner import MambaInnerCL, inner_forward
from .selective_state_update import selective_state_update


def _inner(xz, ndir, scale, params):
    """The fused conv + scan block: through autograd when a gradient can be asked for, else the inference launch - no
    state checkpoints, no saved pre-gate sums, and the chunk-parallel scan over time windows for few long sequences.
    (Inside ``Function.forward`` grad mode is always off and ``needs_input_grad`` is true for parameters even under
    ``torch.no_grad()``, so the distinction has to be made here.)"""
    if torch.is_grad_enabled():
        return MambaInnerCL.apply(xz, ndir, scale, False, *params)
    return inner_forward(xz, ndir, scale, False, params, need_grad=False)[0]


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
        pre = getattr(self, "_A_pre", None)                                             # set by `precomputed_A`
        A = pre[suffix] if pre is not None else -torch.exp(A_log.float())               # bimamba.py:200,222
        return (g("conv1d").weight, g("conv1d").bias, g("x_proj").weight, g("dt_proj").weight, A, Dk.float(),
                g("dt_proj").bias.float())

    def _check(self, hidden_states):
        if not hidden_states.is_cuda:
            raise RuntimeError("mamba_asr_b200.Mamba runs on CUDA only (sm_100a kernels, no CPU fallback)")

    # ---- incremental decoding (reference bimamba.py:176-186, 320-414; same code in mamba_ssm.Mamba) ----------------
    def _cached(self, hidden_states, inference_params):
        """The reference's inference_params protocol: a single-token ``step`` once ``seqlen_offset > 0``, else a
        prefill that leaves the conv window and the last SSM state in the cache (bimamba.py:183-186, 277, 314-316).
        As in the reference, a bidirectional module decodes with its forward-direction parameters only."""
        conv_state, ssm_state = self._get_states_from_cache(inference_params, hidden_states.shape[0])
        if inference_params.seqlen_offset > 0:
            out, _, _ = self.step(hidden_states, conv_state, ssm_state)
            return out
        return self.prefill(hidden_states, conv_state, ssm_state)

    def prefill(self, hidden_states, conv_state, ssm_state, need_output=True):
        """Scan a whole prefix (B, L, d_model) from zero state with the fused kernels and leave the conv window and the last
        SSM state in ``conv_state`` / ``ssm_state`` - from there ``step`` continues token by token.  ``need_output=False``
        skips out_proj (the decoder's cross-Mamba only wants the state its scan over ``memory`` ends in)."""
        with torch.no_grad():
            xz = F.linear(hidden_states, self.in_proj.weight, self.in_proj.bias)          # (B, L, 2D)
            L, D = xz.shape[1], self.d_inner
            W = self.d_conv
            xt = xz[..., :D].transpose(1, 2)                                              # logical (B, D, L)
            conv_state.copy_(F.pad(xt, (W - L, 0)) if L < W else xt[..., L - W:])          # bimamba.py:277
            y, _, _, last = inner_forward(xz, 1, 1.0, False, self._dir_params(""), need_last_state=True)
            ssm_state.copy_(last[0])                                                      # bimamba.py:314-316
            return F.linear(y, self.out_proj.weight, self.out_proj.bias) if need_output else None

    def step(self, hidden_states, conv_state, ssm_state):
        """One token: hidden_states (B, 1, d_model); conv_state (B, D, W) and ssm_state (B, D, N) updated in place.
        Reference bimamba.py:320-365 with both optional kernels present (the sm_100a ones)."""
        self._check(hidden_states)
        assert hidden_states.shape[1] == 1, "Only support decoding with 1 token at a time for now"
        xz = self.in_proj(hidden_states.squeeze(1))                                       # (B, 2D)
        x, z = xz.chunk(2, dim=-1)
        x = causal_conv1d_update(x.to(conv_state.dtype), conv_state, self.conv1d.weight[:, 0, :], self.conv1d.bias,
                                 self.activation).to(xz.dtype)
        x_db = self.x_proj(x)                                                             # (B, R + 2N)
        dt, Bm, Cm = torch.split(x_db, [self.dt_rank, self.d_state, self.d_state], dim=-1)
        dt = F.linear(dt, self.dt_proj.weight)                                            # (B, D); bias joins in-kernel
        A = -torch.exp(self.A_log.float())
        y = selective_state_update(ssm_state, x, dt, A, Bm, Cm, self.D, z=z, dt_bias=self.dt_proj.bias,
                                   dt_softplus=True)
        out = self.out_proj(y)
        return out.unsqueeze(1), conv_state, ssm_state

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        device = self.out_proj.weight.device
        conv_dtype = self.conv1d.weight.dtype if dtype is None else dtype
        conv_state = torch.zeros(batch_size, self.d_model * self.expand, self.d_conv, device=device, dtype=conv_dtype)
        ssm_dtype = self.dt_proj.weight.dtype if dtype is None else dtype
        ssm_state = torch.zeros(batch_size, self.d_model * self.expand, self.d_state, device=device, dtype=ssm_dtype)
        return conv_state, ssm_state

    def _get_states_from_cache(self, inference_params, batch_size, initialize_states=False):
        assert self.layer_idx is not None
        if self.layer_idx not in inference_params.key_value_memory_dict:
            conv_state, ssm_state = self.allocate_inference_cache(batch_size, 0)
            inference_params.key_value_memory_dict[self.layer_idx] = (conv_state, ssm_state)
        else:
            conv_state, ssm_state = inference_params.key_value_memory_dict[self.layer_idx]
            if initialize_states:
                conv_state.zero_()
                ssm_state.zero_()
        return conv_state, ssm_state


class _NegExpMany(torch.autograd.Function):
    """A_i = -exp(A_log_i) for a list of parameters with multi-tensor kernels: two launches forward, one backward
    (dA_log_i = dA_i * A_i) instead of four small elementwise kernels per direction and layer."""

    @staticmethod
    def forward(ctx, *a_logs):
        outs = torch._foreach_exp([a.float() for a in a_logs])
        torch._foreach_neg_(outs)
        ctx.save_for_backward(*outs)
        ctx.set_materialize_grads(False)        # blocks that did not run in this forward get no gradient
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        As = ctx.saved_tensors
        K.flush_reductions()                    # the dA this node reads may still be queued (kernels.deferred_reductions)
        idx = [i for i, g in enumerate(grads) if g is not None]
        prods = torch._foreach_mul([grads[i] for i in idx], [As[i] for i in idx]) if idx else []
        out = [None] * len(As)
        for i, pr in zip(idx, prods):
            out[i] = pr
        return tuple(out)


class precomputed_A:
    """Context manager for a model forward: evaluates ``A = -exp(A_log)`` (reference bimamba.py:200,222) of every Mamba
    block under ``root`` at once and hands the results to the blocks for the duration of the forward pass.  The values
    are recomputed from the current parameters on every entry, so this is an evaluation strategy only."""

    def __init__(self, root):
        self.mods = [m for m in root.modules() if isinstance(m, _MambaBase)]

    def __enter__(self):
        logs, slots = [], []
        for m in self.mods:
            logs.append(m.A_log)
            slots.append((m, ""))
            if hasattr(m, "A_b_log"):
                logs.append(m.A_b_log)
                slots.append((m, "_b"))
        if logs and logs[0].is_cuda and os.environ.get("CM_NO_BATCHED_A") is None:
            As = _NegExpMany.apply(*logs)
            for m in self.mods:
                m._A_pre = {}
            for (m, suffix), A in zip(slots, As):
                A._cm_batched_A = True          # lets the inner block queue its dA reduction (see _NegExpMany.backward)
                m._A_pre[suffix] = A
        return self

    def __exit__(self, *exc):
        for m in self.mods:
            m._A_pre = None
        return False


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
        self._check(hidden_states)
        if inference_params is not None:
            return self._cached(hidden_states, inference_params)
        xz = linear(hidden_states, self.in_proj.weight, self.in_proj.bias)                # (B, L, 2D), time-major
        scale = 0.5 if self.if_devide_out else 1.0                                        # bimamba.py:250-253
        y = _inner(xz, 2, scale, (*self._dir_params(""), *self._dir_params("_b")))
        out = linear(y, self.out_proj.weight, self.out_proj.bias)
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

    def forward(self, hidden_states, inference_params=None, keep_last=None):
        """``keep_last=S`` (extension, default off): project only the last S positions through out_proj - what the
        decoder's cross-Mamba keeps of its scan over [memory ; tgt] (reference modules/Conmamba.py:934)."""
        self._check(hidden_states)
        if inference_params is not None:
            return self._cached(hidden_states, inference_params)
        xz = linear(hidden_states, self.in_proj.weight, self.in_proj.bias)
        y = _inner(xz, 1, 1.0, self._dir_params(""))
        if keep_last is not None:
            y = y[:, -keep_last:]
        return linear(y, self.out_proj.weight, self.out_proj.bias)
