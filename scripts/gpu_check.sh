#!/bin/bash
# Runs the GPU parity suite in separate, time-limited processes (a trapped kernel in one
# group must not hide the results of the others).  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, timeout, pytest -k expr
  echo "=== $1" | tee -a gpurun_out/summary.txt
  timeout "$2" python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$3" -s -p no:cacheprovider > "gpurun_out/$1.log" 2>&1
  echo "exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 25 "gpurun_out/$1.log" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run simt_layer 600 "simt and layer"
run tc_layer 600 "tcgen05 and layer"
run simt_plan 900 "simt and not layer"
run tc_plan 900 "tcgen05 and not layer"
