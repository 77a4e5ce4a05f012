#!/bin/bash
# Round 2, call A: new multi-trip parity tests, graph/no-graph, L2 persistence and sleep A/B on c2, first numbers for
# c3 / c4 on the round-1 wide path, GPU baselines, reference arm.  Every step under a hard timeout.
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L | head -2; echo "host cores: $(nproc)"
echo "== tests"; timeout 1700 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 900 -rf 2>&1 | tail -25
echo "== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
echo "== bench c2 (graph)"; timeout 600 python bench.py --steps 20 --warmup 3 > $O/r2a_bench_c2.json 2> $O/r2a_bench_c2.err; echo "rc=$?"; cut -c1-400 $O/r2a_bench_c2.json; tail -3 $O/r2a_bench_c2.err
echo "== bench c2 (eager chain)"; timeout 300 python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-baseline --no-parity > $O/r2a_bench_c2_nograph.json 2> $O/r2a_bench_c2_nograph.err; cut -c1-330 $O/r2a_bench_c2_nograph.json; tail -2 $O/r2a_bench_c2_nograph.err
echo "== bench c2 (L2 persist)"; TDMPC2_B200_L2_PERSIST=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-parity > $O/r2a_bench_c2_l2p.json 2> $O/r2a_bench_c2_l2p.err; cut -c1-330 $O/r2a_bench_c2_l2p.json; tail -2 $O/r2a_bench_c2_l2p.err
echo "== A/B default / sleep200 / l2persist (per-launch)"; timeout 100 python scripts/profile_iter.py c2 256 4 | tail -3
TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_sleep200.so timeout 100 python scripts/profile_iter.py c2 256 4 | tail -3
TDMPC2_B200_L2_PERSIST=1 timeout 100 python scripts/profile_iter.py c2 256 4 | tail -3
echo "== bench c3 (round-1 wide path), 256 envs"; timeout 900 python bench.py --workload c3 --envs 256 --steps 2 --warmup 3 > $O/r2a_bench_c3_e256.json 2> $O/r2a_bench_c3_e256.err; echo "rc=$?"; cut -c1-400 $O/r2a_bench_c3_e256.json; tail -3 $O/r2a_bench_c3_e256.err
echo "== bench c4 (round-1 wide path), 64 envs"; timeout 900 python bench.py --workload c4 --envs 64 --steps 2 --warmup 3 --no-cpu-baseline > $O/r2a_bench_c4_e64.json 2> $O/r2a_bench_c4_e64.err; echo "rc=$?"; cut -c1-400 $O/r2a_bench_c4_e64.json; tail -3 $O/r2a_bench_c4_e64.err
echo "== reference arm"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2a_bench_ref.json 2> $O/r2a_bench_ref.err; tail -c 600 $O/r2a_bench_ref.json; tail -3 $O/r2a_bench_ref.err
