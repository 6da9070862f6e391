# round 2, call Y: smem-tiled depthwise conv kernels - parity + timing (tile vs round-1 kernels)
set -x
timeout 900 python -m pytest tests/test_gpu_conv_mamba_fbank.py tests/test_gpu_models.py -m gpu -x -q -k "dwconv or convolution_module or ctc or model" 2>&1 | tail -4
timeout 300 python tools/prof_kernels.py --cfg 2,3 --only aux 2>&1 | grep -i dwconv | cut -c1-130
CM_DWCONV_NO_TILE=1 timeout 300 python tools/prof_kernels.py --cfg 2,3 --only aux 2>&1 | grep -i dwconv | cut -c1-130
