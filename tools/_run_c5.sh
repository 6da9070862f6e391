timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_conv_mamba_fbank.py tests/test_gpu_step.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python - <<'PY'
import torch, time, sys
sys.path.insert(0, ".")
from mamba_asr_b200.encoder import build_model
model = build_model("conmamba_large_ctc").cuda().eval()
for secs, Bt in ((300, 4), (20, 64)):
    wav = 0.1 * torch.randn(Bt, 16000 * secs, device="cuda")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for _ in range(2): model(wav)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): model(wav)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("ConMamba-large encoder inference (eager, bf16) %d x %d s: %.2f ms -> %.0f audio-s/s" % (Bt, secs, ms, Bt * secs / ms * 1e3))
PY
