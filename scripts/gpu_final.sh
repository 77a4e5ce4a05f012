#!/bin/bash
# Round-end measurement: full GPU test suite, both bench arms, timeline, ncu launch list + full capture of the dominant kernel.
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -4
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
echo "== bench reference arm"; timeout 400 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json | cut -c1-400; tail -3 gpurun_out/bench.err
echo "== timeline"; TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 120 python scripts/profile_iter.py c2 37 2 > gpurun_out/layer_timeline.txt 2>&1; tail -2 gpurun_out/layer_timeline.txt
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'plan_kernel|pick_kernel|init_state_kernel|distribution' -s 40 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; grep -c plan_kernel gpurun_out/launches.csv
echo "== ncu full (one CEM iteration, c2, E=256)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 4 -c 1 -o gpurun_out/prof_iter_c2 -f python scripts/profile_iter.py c2 256 3 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
