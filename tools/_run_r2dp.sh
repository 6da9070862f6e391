set -x
mkdir -p gpurun_out
export CM_DWCONV_MMA=1
timeout 900 python -m pytest tests/test_gpu_conv_mamba_fbank.py tests/test_gpu_fused_ln.py -x -q -m gpu -k "depthwise or convolution_module or conv_module" > gpurun_out/r2dp_tests.log 2>&1; tail -3 gpurun_out/r2dp_tests.log | cut -c1-220
timeout 600 python tools/step_profile.py --graphed --top 30 > gpurun_out/r2dp_step_large.txt 2>&1; grep -i "dwconv\|total CUDA" gpurun_out/r2dp_step_large.txt | cut -c1-150
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2dp_large.log 2> gpurun_out/r2dp_large.err
unset CM_DWCONV_MMA
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2dp_large_ffma.log 2> gpurun_out/r2dp_large_ffma.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2dp_large*.log")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"],1), d.get("ms_per_step"))
PY
