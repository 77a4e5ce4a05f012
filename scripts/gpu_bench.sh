#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 1200 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
