#!/bin/bash
# Round-end measurement: full GPU test suite, ncu capture of the dominant kernel (-> DRAM traffic), both bench arms,
# timelines, ncu launch list.  Every step runs under a hard timeout; outputs land in gpurun_out/ and are copied to
# profiles/ by hand afterwards.
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -4
echo "== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
if [ -n "$1" ]; then   # optional A/B against an experiment build (build.build_variant)
  echo "== A/B default"; timeout 100 python scripts/profile_iter.py c2 256 4 | tail -3
  echo "== A/B $1"; TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_$1.so timeout 100 python scripts/profile_iter.py c2 256 4 | tail -3
fi
echo "== ncu full (one CEM iteration, c2, E=256)"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 4 -c 1 -o gpurun_out/prof_iter_c2 -f python scripts/profile_iter.py c2 256 3 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
timeout 100 python scripts/extract_traffic.py gpurun_out/prof_iter_c2.ncu-rep profiles/r01_traffic.json gpurun_out/traffic.json | cut -c1-260
echo "== bench"; timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
echo "== timelines"; TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 100 python scripts/profile_iter.py c2 37 2 > gpurun_out/layer_timeline.txt 2>&1; tail -1 gpurun_out/layer_timeline.txt
TDMPC2_ENGINE=tcgen05pp TDMPC2_TRACE=pp TDMPC2_PHASE_PROF=1 timeout 100 python scripts/profile_iter.py c2 37 2 > gpurun_out/pp_timeline.txt 2>&1; tail -1 gpurun_out/pp_timeline.txt
echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'plan_kernel|pick_kernel|init_state_kernel|distribution' -s 40 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; grep -c plan_kernel gpurun_out/launches.csv
timeout 60 scripts/micro/tma_bw > gpurun_out/tma_bw.txt 2>&1; timeout 30 scripts/micro/mma_rate > gpurun_out/mma_rate.txt 2>&1; tail -3 gpurun_out/tma_bw.txt
