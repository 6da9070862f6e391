# round 2, call U: one-kernel Fbank front-end (parity + timing) and compute-sanitizer over the hand-synchronised kernels
set -x
timeout 600 python -m pytest tests/test_gpu_conv_mamba_fbank.py -m gpu -x -q -k fbank 2>&1 | tail -5
python - <<'PY'
import torch, os, time
from mamba_asr_b200 import Fbank
for n_fft, win in ((512, 25), (400, 25)):
    fb = Fbank(n_fft=n_fft, n_mels=80, win_length=win).cuda()
    wav = torch.randn(64, 320000, device="cuda")
    for route in ("dft", "cufft"):
        if route == "cufft": os.environ["CM_FBANK_CUFFT"] = "1"
        else: os.environ.pop("CM_FBANK_CUFFT", None)
        for _ in range(3): fb(wav)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(10):
            e0.record(); fb(wav); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print("fbank 64 x 20 s n_fft %d win %d ms route %-5s: best %.3f ms" % (n_fft, win, route, min(ts)))
PY
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/r2u_memcheck.log 2>&1; tail -4 gpurun_out/r2u_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/r2u_racecheck.log 2>&1; tail -4 gpurun_out/r2u_racecheck.log
