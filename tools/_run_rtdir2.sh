for v in "1 1" "0 0"; do set -- $v
  export CM_NVCC_EXTRA="-DCM_BWDSP_RTDIR_IO=$1 -DCM_FWDSP_RTDIR_IO=$2"
  python mamba_asr_b200/build.py >/dev/null 2>&1 || echo BUILD FAIL
  echo "== BWD_RTDIR_IO=$1 FWD_RTDIR_IO=$2"
  timeout 300 python tools/prof_kernels.py --cfg 2,3,4,5_4k --only scan_fwd,scan_bwd,scan_fwd_infer 2>&1 | grep "scan_" | cut -c1-130
done
unset CM_NVCC_EXTRA
python mamba_asr_b200/build.py >/dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_conv_mamba_fbank.py -q -m gpu -x 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s3_rt_small.log 2> gpurun_out/s3_rt_small.err
timeout 600 python bench.py --steps 8 --warmup 3 --workload conmamba_large_ctc_fwdbwd_b64x20s --no-cpu-baseline > gpurun_out/s3_rt_large.log 2> gpurun_out/s3_rt_large.err
python - <<'PY'
import json
for f in ["s3_rt_small","s3_rt_large"]:
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), d["roofline"]["kernel"], round(d["roofline"]["frac"],4), round(d["roofline"]["avg_launch_ms"],4), d["loss"])
    except Exception as e: print(f, "ERR", e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
PY
