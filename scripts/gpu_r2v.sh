#!/bin/bash
# Round 2, call V (last): in-kernel noise mode (tdmpc2_plan_iter_rng) -- tests, timing, traffic re-captures on the final library,
# bench lines (default, and the declared non-parity modes).
mkdir -p gpurun_out; O=gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -rf 2>&1 | tail -6
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
it() { timeout 300 python scripts/profile_iter.py $1 $2 4 | tail -2 | tr '\n' ' '; echo; }
for eng in tcgen05x2 tcgen05pp; do echo -n "c2 $eng: "; TDMPC2_ENGINE=$eng it c2 256; done
echo -n "c3: "; it c3 1024; echo -n "c4: "; it c4 256
echo "== ncu full c2 (CTA-pair kernel)"; timeout 500 ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 3 -c 1 -o $O/r02_iter_c2 -f env TDMPC2_ENGINE=tcgen05x2 python scripts/profile_iter.py c2 256 3 > $O/ncu_c2.log 2>&1; tail -1 $O/ncu_c2.log
timeout 100 python scripts/extract_traffic.py $O/r02_iter_c2.ncu-rep c2 256 $O/r02_traffic_c2.json | cut -c1-300
echo "== ncu full c2 (ping-pong kernel)"; timeout 500 ncu --set full --clock-control none --import-source on -k regex:plan_pp_kernel -s 1 -c 1 -o $O/r02_iter_c2_pp -f env TDMPC2_ENGINE=tcgen05pp python scripts/profile_iter.py c2 256 3 > $O/ncu_c2pp.log 2>&1; tail -1 $O/ncu_c2pp.log
timeout 100 python scripts/extract_traffic.py $O/r02_iter_c2_pp.ncu-rep c2 256 $O/r02_traffic_c2_pp.json | cut -c1-300
echo "== bench c2 (default)"; timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r02_bench_c2_final.json 2> $O/r02_bench_c2_final.err; echo "rc=$?"; cut -c1-260 $O/r02_bench_c2_final.json; tail -2 $O/r02_bench_c2_final.err
echo "== bench c2 in-kernel noise (declared non-parity)"; timeout 600 python bench.py --rng philox --steps 20 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r02_bench_c2_philox.json 2> $O/r02_bench_c2_philox.err; echo "rc=$?"; cut -c1-260 $O/r02_bench_c2_philox.json; tail -2 $O/r02_bench_c2_philox.err
