#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/profile_iter.py c2 37 3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 3 -c 1 -o gpurun_out/prof_iter -f python scripts/profile_iter.py c2 37 3 > gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log; ls -la gpurun_out/
