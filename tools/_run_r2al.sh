# round 2, call AL: sliding-window conv forward - parity + timing; ncu --set full of both sliding-window conv kernels
set -x
timeout 900 python -m pytest tests/test_gpu_conv_mamba_fbank.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/prof_kernels.py --cfg 2,3,4 --only conv_fwd,conv_bwd 2>&1 | cut -c1-130
CM_CONV_NO_SW=1 timeout 300 python tools/prof_kernels.py --cfg 3 --only conv_fwd 2>&1 | cut -c1-130
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_sw -s 2 -c 1 -o gpurun_out/r2al_conv_fwd_sw_cfg3 python tools/prof_kernels.py --cfg 3 --only conv_fwd --iters 3 > gpurun_out/r2al_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_bwd_sw -s 2 -c 1 -o gpurun_out/r2al_conv_bwd_sw_cfg3 python tools/prof_kernels.py --cfg 3 --only conv_bwd --iters 3 > gpurun_out/r2al_ncu2.log 2>&1
