#!/bin/bash
# Call X: the vectorised evaluation loop on the planner (toy environments), plus the host-noise CPU tests on the GPU box.
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_evaluate.py tests/test_host_noise_cpu.py tests/test_evaluate_cpu.py -x -q -s 2>&1 | tail -15 | tee gpurun_out/r2x_tests.txt
