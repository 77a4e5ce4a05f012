#!/bin/bash
# Round 2, call U (last): evict-last store hint on by default -- tests, ncu captures of the c2 iteration on both engines ->
# traffic JSONs of the final library, bench c2 with every leg.
mkdir -p gpurun_out; O=gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -rf 2>&1 | tail -5
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
echo "== ncu full c2 (CTA-pair kernel)"; timeout 500 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 3 -c 1 -o $O/r02_iter_c2 -f env TDMPC2_ENGINE=tcgen05x2 python scripts/profile_iter.py c2 256 3 > $O/ncu_c2.log 2>&1; tail -1 $O/ncu_c2.log
timeout 100 python scripts/extract_traffic.py $O/r02_iter_c2.ncu-rep c2 256 $O/r02_traffic_c2.json | cut -c1-330
echo "== ncu full c2 (ping-pong kernel)"; timeout 500 ncu --set full --clock-control none --import-source on -k regex:plan_pp_kernel -s 1 -c 1 -o $O/r02_iter_c2_pp -f env TDMPC2_ENGINE=tcgen05pp python scripts/profile_iter.py c2 256 3 > $O/ncu_c2pp.log 2>&1; tail -1 $O/ncu_c2pp.log
timeout 100 python scripts/extract_traffic.py $O/r02_iter_c2_pp.ncu-rep c2 256 $O/r02_traffic_c2_pp.json | cut -c1-330
echo "== bench c2 (default)"; timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02_bench_c2.json 2> $O/r02_bench_c2.err; echo "rc=$?"; cut -c1-260 $O/r02_bench_c2.json; tail -2 $O/r02_bench_c2.err
echo "== bench c2 ping-pong"; timeout 600 python bench.py --engine tcgen05pp --steps 20 --warmup 3 --no-gpu-baseline --no-cpu-baseline --no-parity > $O/r02_bench_c2_pp_final.json 2>/dev/null; cut -c1-260 $O/r02_bench_c2_pp_final.json
echo "== c3 / c4 iteration"; for WL_E in "c3 1024" "c4 256"; do set -- $WL_E; timeout 300 python scripts/profile_iter.py $1 $2 3 | tail -1; done
