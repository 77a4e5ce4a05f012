#!/bin/bash
# Call Z: one-environment act(): interleaved draws vs graph replay (bit-identity tests + latency breakdown on c1).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_edges.py tests/test_gpu_pixels.py tests/test_gpu_evaluate.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2z_tests.txt
timeout 200 python scripts/act_latency_breakdown.py c1 2>&1 | tail -12 | tee gpurun_out/r2z_act_latency_c1.txt
