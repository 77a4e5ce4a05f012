#!/bin/bash
# Round 2, call G: evidence run on the lean kernels -- tests, ncu captures (c2 full batch, c4 one wave), launch list,
# timelines, bench lines of all four workloads.
mkdir -p gpurun_out; O=gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -rf 2>&1 | tail -6
echo "== ncu full c2 (E=256, one CEM iteration)"; timeout 500 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 4 -c 1 -o $O/r02_iter_c2 -f python scripts/profile_iter.py c2 256 3 > $O/ncu_c2.log 2>&1; tail -1 $O/ncu_c2.log
timeout 100 python scripts/extract_traffic.py $O/r02_iter_c2.ncu-rep c2 256 $O/r02_traffic_c2.json | cut -c1-300
echo "== ncu full c4 (E=256 is 165 ms x 40 replays: use E=37, one wave)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 4 -c 1 -o $O/r02_iter_c4 -f python scripts/profile_iter.py c4 37 3 > $O/ncu_c4.log 2>&1; tail -1 $O/ncu_c4.log
timeout 100 python scripts/extract_traffic.py $O/r02_iter_c4.ncu-rep c4 37 $O/r02_traffic_c4_e37.json | cut -c1-300
echo "== timeline c2 (prof build)"; TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_prof.so TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c2 37 2 > $O/r02_timeline_c2.txt 2>&1; sed -n 1,16p $O/r02_timeline_c2.txt
echo "== bench c2"; timeout 600 python bench.py --steps 20 --warmup 3 > $O/r02_bench_c2.json 2> $O/r02_bench_c2.err; echo "rc=$?"; cut -c1-260 $O/r02_bench_c2.json; tail -2 $O/r02_bench_c2.err
echo "== ncu launch list (c2 bench steps)"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $O/r02_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity > $O/ncu_bench.log 2>&1; grep -c plan_kernel $O/r02_launches_c2.csv
echo "== bench c3"; timeout 900 python bench.py --workload c3 --steps 3 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r02_bench_c3.json 2> $O/r02_bench_c3.err; echo "rc=$?"; cut -c1-260 $O/r02_bench_c3.json; tail -2 $O/r02_bench_c3.err
echo "== bench c4"; timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err; echo "rc=$?"; cut -c1-260 $O/r02_bench_c4.json; tail -2 $O/r02_bench_c4.err
echo "== bench c5"; timeout 1200 python bench.py --workload c5 --steps 2 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r02_bench_c5.json 2> $O/r02_bench_c5.err; echo "rc=$?"; cut -c1-260 $O/r02_bench_c5.json; tail -2 $O/r02_bench_c5.err
