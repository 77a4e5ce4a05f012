#!/bin/bash
# Call W: the three branch / knob goldens (no policy-prior trajectories, one-step horizon, non-default planner knobs) on the GPU.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py -x -q -m gpu -k "nopi or h1 or knobs" -s 2>&1 | tail -15 | tee gpurun_out/r2w_tests.txt
