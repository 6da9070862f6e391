set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_mamba_fbank.py tests/test_gpu_fused_ln.py -x -q -m gpu -k "depthwise or convolution_module or conv_module" > gpurun_out/r2dl_tests.log 2>&1; tail -4 gpurun_out/r2dl_tests.log | cut -c1-220
timeout 300 python tools/prof_kernels.py --cfg 3 --only aux > gpurun_out/r2dl_aux.txt 2>&1; grep -i dwconv gpurun_out/r2dl_aux.txt
CM_DWCONV_NO_MMA=1 timeout 300 python tools/prof_kernels.py --cfg 3 --only aux 2>&1 | grep -i dwconv
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2dl_large.log 2> gpurun_out/r2dl_large.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2dl_large.log").read().strip().splitlines()[-1])
print(round(d["value"],1), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"))
PY
