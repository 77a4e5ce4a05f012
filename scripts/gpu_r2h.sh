#!/bin/bash
# Round 2, call H: shared-latent fold + fast mode -- tests, same-box A/B (fold on/off, L2 persistence on/off, passes 1 vs 3),
# fast-mode bench line, timeline with the fold.
mkdir -p gpurun_out; O=gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -rf 2>&1 | tail -8
it() { timeout 200 python scripts/profile_iter.py $1 $2 4 | tail -2 | tr '\n' ' '; echo; }
for rep in 1 2; do
  for WL_E in "c2 256" "c3 1024" "c4 256"; do
    set -- $WL_E
    echo -n "[$rep] $1 fold=1: "; it $1 $2
    echo -n "[$rep] $1 fold=0: "; TDMPC2_B200_ZFOLD=0 it $1 $2
  done
done
for WL_E in "c2 256" "c4 256"; do
  set -- $WL_E
  echo -n "$1 l2persist=1: "; TDMPC2_B200_L2_PERSIST=1 it $1 $2
  echo -n "$1 passes=1   : "; TDMPC2_B200_PASSES=1 it $1 $2
done
echo -n "c3 passes=1   : "; TDMPC2_B200_PASSES=1 it c3 1024
echo "== timeline c2 (prof build, fold on)"; TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_prof.so TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c2 37 2 > $O/r02_timeline_c2_fold.txt 2>&1; sed -n 1,16p $O/r02_timeline_c2_fold.txt
echo "== timeline c2 fast mode"; TDMPC2_B200_PASSES=1 TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_prof.so TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c2 37 2 > $O/r02_timeline_c2_fast.txt 2>&1; sed -n 1,16p $O/r02_timeline_c2_fast.txt
echo "== bench c2 fast mode"; timeout 600 python bench.py --passes 1 --steps 10 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r02_bench_c2_fast.json 2> $O/r02_bench_c2_fast.err; echo "rc=$?"; cut -c1-300 $O/r02_bench_c2_fast.json; tail -2 $O/r02_bench_c2_fast.err
echo "== bench c2 l2 persist (power / clocks)"; TDMPC2_B200_L2_PERSIST=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-baseline --no-cpu-baseline --no-parity > $O/r02_bench_c2_l2persist.json 2> $O/r02_bench_c2_l2persist.err; echo "rc=$?"; cut -c1-300 $O/r02_bench_c2_l2persist.json
echo "== bench c2"; timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r02_bench_c2_fold.json 2> $O/r02_bench_c2_fold.err; echo "rc=$?"; cut -c1-300 $O/r02_bench_c2_fold.json
echo "== ncu launch list: one plan() step of c2, c3, c4"; for WL in c2 c3 c4; do timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r02_launches_$WL.csv python scripts/launch_list.py $WL > $O/ncu_ll_$WL.log 2>&1; python scripts/summarize_launches.py $O/r02_launches_$WL.csv | tee $O/r02_launches_$WL.txt | head -12; done
