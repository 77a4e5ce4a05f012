#!/bin/bash
# Round 2 evidence run on the final kernels: tests, smoke, ncu captures (c2 ping-pong iteration, c3 and c4 one wave) ->
# traffic JSONs, launch lists of one plan() step, in-kernel timelines, bench lines of all four workloads (+ the declared
# non-parity mode and the reference arm).  Outputs in gpurun_out/, copied to profiles/ by hand.
mkdir -p gpurun_out; O=gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -rf 2>&1 | tail -6
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== ncu full c2 (E=256, one CEM iteration, ping-pong kernel)"; timeout 500 ncu --set full --clock-control none --import-source on -k regex:plan_pp_kernel -s 1 -c 1 -o $O/r02_iter_c2 -f env TDMPC2_ENGINE=tcgen05pp python scripts/profile_iter.py c2 256 3 > $O/ncu_c2.log 2>&1; tail -1 $O/ncu_c2.log
TDMPC2_ENGINE=tcgen05pp timeout 100 python scripts/extract_traffic.py $O/r02_iter_c2.ncu-rep c2 256 $O/r02_traffic_c2.json | cut -c1-400
echo "== ncu full c4 (one wave, E=37)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 4 -c 1 -o $O/r02_iter_c4 -f python scripts/profile_iter.py c4 37 3 > $O/ncu_c4.log 2>&1; tail -1 $O/ncu_c4.log
timeout 100 python scripts/extract_traffic.py $O/r02_iter_c4.ncu-rep c4 37 $O/r02_traffic_c4_e37.json | cut -c1-400
echo "== ncu full c3 (one wave, E=37)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 4 -c 1 -o $O/r02_iter_c3 -f python scripts/profile_iter.py c3 37 3 > $O/ncu_c3.log 2>&1; tail -1 $O/ncu_c3.log
timeout 100 python scripts/extract_traffic.py $O/r02_iter_c3.ncu-rep c3 37 $O/r02_traffic_c3_e37.json | cut -c1-400
echo "== ncu launch lists: one plan() step"; for WL in c2 c3 c4; do timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r02_launches_$WL.csv python scripts/launch_list.py $WL > $O/ncu_ll_$WL.log 2>&1; python scripts/summarize_launches.py $O/r02_launches_$WL.csv | tee $O/r02_launches_$WL.txt | head -6; done
echo "== timelines"; P=/root/repo/tdmpc2_b200/libtdmpc2_b200_prof.so
TDMPC2_ENGINE=tcgen05pp TDMPC2_B200_LIB=$P TDMPC2_TRACE=pp TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c2 37 2 > $O/r02_timeline_c2_pp.txt 2>&1; sed -n 9,12p $O/r02_timeline_c2_pp.txt
TDMPC2_ENGINE=tcgen05x2 TDMPC2_B200_LIB=$P TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c2 37 2 > $O/r02_timeline_c2_x2.txt 2>&1; sed -n 9,12p $O/r02_timeline_c2_x2.txt
TDMPC2_B200_LIB=$P TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 300 python scripts/profile_iter.py c4 37 2 > $O/r02_timeline_c4.txt 2>&1; sed -n 8,10p $O/r02_timeline_c4.txt
TDMPC2_B200_LIB=$P TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 300 python scripts/profile_iter.py c3 37 2 > $O/r02_timeline_c3.txt 2>&1; sed -n 8,10p $O/r02_timeline_c3.txt
echo "== bench c2"; timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02_bench_c2.json 2> $O/r02_bench_c2.err; echo "rc=$?"; cut -c1-260 $O/r02_bench_c2.json; tail -2 $O/r02_bench_c2.err
echo "== bench c2 CTA-pair engine"; timeout 600 python bench.py --engine tcgen05x2 --steps 20 --warmup 3 --no-gpu-baseline --no-cpu-baseline --no-parity > $O/r02_bench_c2_x2.json 2>/dev/null; cut -c1-260 $O/r02_bench_c2_x2.json
echo "== bench c2 non-parity single-pass mode"; timeout 600 python bench.py --passes 1 --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r02_bench_c2_fastmode.json 2>/dev/null; cut -c1-260 $O/r02_bench_c2_fastmode.json
echo "== bench c3"; timeout 900 python bench.py --workload c3 --steps 3 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r02_bench_c3.json 2> $O/r02_bench_c3.err; echo "rc=$?"; cut -c1-260 $O/r02_bench_c3.json
echo "== bench c4"; timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err; echo "rc=$?"; cut -c1-260 $O/r02_bench_c4.json
echo "== bench c5"; timeout 1200 python bench.py --workload c5 --steps 2 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r02_bench_c5.json 2> $O/r02_bench_c5.err; echo "rc=$?"; cut -c1-260 $O/r02_bench_c5.json
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; echo "rc=$?"; cut -c1-400 $O/r02_bench_reference_arm.json
