# round 2, call CA: dwconv forward tile kernel (single-warp CTAs, circular register window): parity + timing
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_mamba_fbank.py -x -q -m gpu -k "depthwise or conv_module or dwconv" > gpurun_out/r2ca_tests.log 2>&1; tail -3 gpurun_out/r2ca_tests.log
timeout 300 python tools/prof_kernels.py --cfg 3 --only aux > gpurun_out/r2ca_aux.txt 2>&1; grep -i "dwconv" gpurun_out/r2ca_aux.txt
