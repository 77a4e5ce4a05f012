#!/bin/bash
# first measurement: smoke, bench (ours + reference arm), ncu launch list
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
grep -c plan_kernel gpurun_out/launches.csv
