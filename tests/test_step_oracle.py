"""CPU: the incremental-decoding oracle (oracle/step_ref.py) against the fixtures frozen from the reference's own
``Mamba.forward(inference_params=...)`` / ``Mamba.step`` (tests/golden/mamba_step.npz, reference bimamba.py:176-186,
320-414)."""
import os

import numpy as np
import torch


def _load(golden_dir):
    z = np.load(os.path.join(golden_dir, "mamba_step.npz"))
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p_")}
    return z, sd


def test_step_oracle_matches_reference_from_zero_cache(golden_dir):
    from oracle.step_ref import mamba_step_oracle
    z, sd = _load(golden_dir)
    tokens = torch.from_numpy(z["tokens"])
    Bt, T, _ = tokens.shape
    Dn, W, N = sd["conv1d.weight"].shape[0], sd["conv1d.weight"].shape[-1], sd["A_log"].shape[1]
    cs, ss = torch.zeros(Bt, Dn, W), torch.zeros(Bt, Dn, N)
    outs = [mamba_step_oracle(tokens[:, t:t + 1], cs, ss, sd) for t in range(T)]
    torch.testing.assert_close(torch.cat(outs, 1), torch.from_numpy(z["out_steps_zero"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(cs, torch.from_numpy(z["conv_zero_final"]), rtol=0, atol=0)
    torch.testing.assert_close(ss, torch.from_numpy(z["ssm_zero_final"]), rtol=1e-5, atol=1e-7)


def test_prefill_then_step_oracle_matches_reference(golden_dir):
    from oracle.step_ref import mamba_prefill_oracle, mamba_step_oracle
    z, sd = _load(golden_dir)
    prompt, tokens = torch.from_numpy(z["prompt"]), torch.from_numpy(z["tokens"])
    Bt, T, _ = tokens.shape
    Dn, W, N = sd["conv1d.weight"].shape[0], sd["conv1d.weight"].shape[-1], sd["A_log"].shape[1]
    cs, ss = torch.zeros(Bt, Dn, W), torch.zeros(Bt, Dn, N)
    out = mamba_prefill_oracle(prompt, cs, ss, sd)
    torch.testing.assert_close(out, torch.from_numpy(z["out_prefill"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(cs, torch.from_numpy(z["conv_after_prefill"]), rtol=0, atol=0)
    torch.testing.assert_close(ss, torch.from_numpy(z["ssm_after_prefill"]), rtol=1e-5, atol=1e-7)
    outs = [mamba_step_oracle(tokens[:, t:t + 1], cs, ss, sd) for t in range(T)]
    torch.testing.assert_close(torch.cat(outs, 1), torch.from_numpy(z["out_steps"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ss, torch.from_numpy(z["ssm_final"]), rtol=1e-5, atol=1e-7)


def test_stepping_equals_the_full_sequence_scan(golden_dir):
    """Property the domain offers: T single-token steps from a zero cache = the causal scan over the T tokens."""
    from oracle.step_ref import mamba_prefill_oracle, mamba_step_oracle
    z, sd = _load(golden_dir)
    tokens = torch.from_numpy(z["tokens"])
    Bt, T, _ = tokens.shape
    Dn, W, N = sd["conv1d.weight"].shape[0], sd["conv1d.weight"].shape[-1], sd["A_log"].shape[1]
    cs, ss = torch.zeros(Bt, Dn, W), torch.zeros(Bt, Dn, N)
    full = mamba_prefill_oracle(tokens, torch.zeros_like(cs), torch.zeros_like(ss), sd)
    outs = [mamba_step_oracle(tokens[:, t:t + 1], cs, ss, sd) for t in range(T)]
    torch.testing.assert_close(torch.cat(outs, 1), full, rtol=1e-4, atol=1e-6)
