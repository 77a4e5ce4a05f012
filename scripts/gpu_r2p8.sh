#!/bin/bash
# Round 2, 8-GPU call: BASELINE config c4 (mt80 317M, 2048 envs over 8 GPUs) as stated, the on-hardware sharded == unsharded
# test, and a 2-rank c2 line (one all-gather per plan).
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== multi-GPU correctness test"; timeout 600 python -m pytest tests/test_gpu_multigpu.py -q -m gpu -p no:cacheprovider --timeout 500 -rs 2>&1 | tail -4
echo "== c4 on 8 GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload c4 --steps 3 --warmup 3 --no-gpu-baseline --no-cpu-baseline > $O/r02_bench_c4_n8.json 2> $O/r02_bench_c4_n8.err; echo "rc=$?"; cut -c1-400 $O/r02_bench_c4_n8.json; tail -2 $O/r02_bench_c4_n8.err
echo "== c2 on 8 GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r02_bench_c2_n8.json 2> $O/r02_bench_c2_n8.err; echo "rc=$?"; cut -c1-300 $O/r02_bench_c2_n8.json
