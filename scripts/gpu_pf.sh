#!/bin/bash
# W-prefetch engine (tcgen05x2pf): bit-identity test, A/B timing, bench, default-engine tests with it, ncu evidence.
# Ordered by importance: the call may be cut short by the remaining GPU budget.
mkdir -p gpurun_out
echo "== bit-identity"; timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "prefetch or cta_pair" -p no:cacheprovider 2>&1 | tail -3
echo "== A/B"; TDMPC2_ENGINE=tcgen05x2 timeout 60 python scripts/profile_iter.py c2 256 4 | tail -2
TDMPC2_ENGINE=tcgen05x2pf timeout 60 python scripts/profile_iter.py c2 256 4 | tail -2
echo "== bench pf"; timeout 120 python bench.py --engine tcgen05x2pf --steps 20 --warmup 3 > gpurun_out/bench_pf.json 2> gpurun_out/bench_pf.err; cut -c1-260 gpurun_out/bench_pf.json
echo "== default-engine tests under pf"; TDMPC2_B200_ENGINE=tcgen05x2pf timeout 200 python -m pytest tests/test_gpu_golden.py tests/test_gpu_edges.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
echo "== timeline pf"; TDMPC2_ENGINE=tcgen05x2pf TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 60 python scripts/profile_iter.py c2 37 2 > gpurun_out/layer_timeline_pf.txt 2>&1; tail -1 gpurun_out/layer_timeline_pf.txt
echo "== ncu full pf"; TDMPC2_ENGINE=tcgen05x2pf timeout 200 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 4 -c 1 -o gpurun_out/prof_iter_c2_pf -f python scripts/profile_iter.py c2 256 3 > gpurun_out/ncu_full_pf.log 2>&1; tail -1 gpurun_out/ncu_full_pf.log
timeout 60 python scripts/extract_traffic.py gpurun_out/prof_iter_c2_pf.ncu-rep gpurun_out/traffic_pf.json | cut -c1-200
echo "== ncu launch list pf"; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'plan_kernel|pick_kernel|init_state_kernel|distribution' -s 40 -c 60 --csv --log-file gpurun_out/launches_pf.csv python bench.py --engine tcgen05x2pf --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_pf.log 2>&1; grep -c plan_kernel gpurun_out/launches_pf.csv
