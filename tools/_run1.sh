timeout 600 python -m pytest tests/test_gpu_conv_mamba_fbank.py -q -m gpu -k "layernorm" 2>&1 | tail -3
timeout 300 python tools/prof_kernels.py --cfg 2,3 --only aux 2>&1 | grep "ln_"
