# round 2, call W: sliding-window conv backward - parity + timing
set -x
timeout 900 python -m pytest tests/test_gpu_conv_mamba_fbank.py -m gpu -x -q -k "conv or mamba or bimamba" 2>&1 | tail -4
timeout 300 python tools/prof_kernels.py --cfg 2,3,4 --only conv_bwd 2>&1 | cut -c1-130
CM_CONV_NO_SW=1 timeout 300 python tools/prof_kernels.py --cfg 3 --only conv_bwd 2>&1 | cut -c1-130
