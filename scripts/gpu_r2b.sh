#!/bin/bash
# Round 2, call B: the rewritten wide path (super-chunks + drain + single pass 2, CTA pairs): parity, K-segment accuracy
# sweep, first full-size numbers for c3 / c4 / c5 (per-GPU share), c2 regression check.
mkdir -p gpurun_out; O=gpurun_out
echo "host: $(nproc) cpus; cgroup cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; lscpu | grep -E "Model name|Socket|Core|Thread" | head -5
echo "== tests"; timeout 1700 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 -rf -x 2>&1 | tail -30
echo "== kseg sweep c3"; timeout 600 python scripts/kseg_error.py c3 3 0 896 512 256 2>&1 | tail -6
echo "== kseg sweep c4"; timeout 900 python scripts/kseg_error.py c4 3 0 2048 1024 512 256 2>&1 | tail -7
echo "== bench c2"; timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-baseline > $O/r2b_bench_c2.json 2> $O/r2b_bench_c2.err; echo "rc=$?"; cut -c1-300 $O/r2b_bench_c2.json; tail -3 $O/r2b_bench_c2.err
echo "== bench c3"; timeout 900 python bench.py --workload c3 --steps 3 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r2b_bench_c3.json 2> $O/r2b_bench_c3.err; echo "rc=$?"; cut -c1-300 $O/r2b_bench_c3.json; tail -3 $O/r2b_bench_c3.err
echo "== bench c4"; timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r2b_bench_c4.json 2> $O/r2b_bench_c4.err; echo "rc=$?"; cut -c1-300 $O/r2b_bench_c4.json; tail -3 $O/r2b_bench_c4.err
echo "== bench c5"; timeout 1200 python bench.py --workload c5 --steps 2 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r2b_bench_c5.json 2> $O/r2b_bench_c5.err; echo "rc=$?"; cut -c1-300 $O/r2b_bench_c5.json; tail -3 $O/r2b_bench_c5.err
