#!/bin/bash
# A/B of an experiment build (tdmpc2_b200/libtdmpc2_b200_<name>.so, see build.build_variant) against the default library:
# parity tests on the variant, short un-capped iteration timings, then the sustained (power-capped) 20-step bench of both.
name=${1:-poll1}
EXP=/root/repo/tdmpc2_b200/libtdmpc2_b200_${name}.so
mkdir -p gpurun_out
echo "== parity with $EXP"; TDMPC2_B200_LIB=$EXP timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -m gpu -x -k "not simt" -p no:cacheprovider 2>&1 | tail -3
for eng in tcgen05x2 tcgen05pp; do
  echo "== $eng default lib"; TDMPC2_ENGINE=$eng timeout 100 python scripts/profile_iter.py c2 256 4 | tail -3
  echo "== $eng $name lib"; TDMPC2_B200_LIB=$EXP TDMPC2_ENGINE=$eng timeout 100 python scripts/profile_iter.py c2 256 4 | tail -3
done
summ='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["clocks"])'
echo "== sustained bench, default lib"; timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$summ"
echo "== sustained bench, $name lib"; TDMPC2_B200_LIB=$EXP timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$summ"
