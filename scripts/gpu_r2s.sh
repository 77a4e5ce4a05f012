#!/bin/bash
# Round 2, call S: the default engine rule ("auto": ping-pong for single-trip batches, CTA pairs beyond) -- tests, ncu capture of
# the CTA-pair iteration at c2 (the kernel the default bench times), bench c2 with every leg.
mkdir -p gpurun_out; O=gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -rf 2>&1 | tail -5
echo "== ncu full c2 (E=256, CTA-pair kernel)"; timeout 500 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 3 -c 1 -o $O/r02_iter_c2 -f python scripts/profile_iter.py c2 256 3 > $O/ncu_c2.log 2>&1; tail -1 $O/ncu_c2.log
timeout 100 python scripts/extract_traffic.py $O/r02_iter_c2.ncu-rep c2 256 $O/r02_traffic_c2.json | cut -c1-400
echo "== bench c2 (default engine)"; timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02_bench_c2.json 2> $O/r02_bench_c2.err; echo "rc=$?"; cut -c1-300 $O/r02_bench_c2.json; tail -2 $O/r02_bench_c2.err
echo "== bench c2 ping-pong engine"; timeout 600 python bench.py --engine tcgen05pp --steps 20 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r02_bench_c2_pp.json 2>/dev/null; cut -c1-300 $O/r02_bench_c2_pp.json
echo "== launch list c2"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r02_launches_c2.csv python scripts/launch_list.py c2 > $O/ncu_ll_c2.log 2>&1; python scripts/summarize_launches.py $O/r02_launches_c2.csv | tee $O/r02_launches_c2.txt | head -6
