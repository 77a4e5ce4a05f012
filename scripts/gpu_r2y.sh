#!/bin/bash
# Call Y: where one-environment act() time goes (c1, c3 at E = 1).
mkdir -p gpurun_out
timeout 200 python scripts/act_latency_breakdown.py c1 2>&1 | tail -12 | tee gpurun_out/r2y_act_latency_c1.txt
TDMPC2_B200_ENGINE=tcgen05x2 timeout 200 python scripts/act_latency_breakdown.py c1 2>&1 | tail -12 | tee gpurun_out/r2y_act_latency_c1_x2.txt
TDMPC2_B200_ENGINE=tcgen05 timeout 200 python scripts/act_latency_breakdown.py c1 2>&1 | tail -12 | tee gpurun_out/r2y_act_latency_c1_1cta.txt
