#!/bin/bash
mkdir -p gpurun_out
TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 120 python scripts/profile_iter.py c2 37 2 > gpurun_out/layer_timeline.txt 2>&1; tail -3 gpurun_out/layer_timeline.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'plan_kernel|pick_kernel|init_state_kernel|distribution|bitonic|vectorized' -s 60 -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
grep -c plan_kernel gpurun_out/launches.csv
timeout 300 python -m pytest tests/test_gpu_edges.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
