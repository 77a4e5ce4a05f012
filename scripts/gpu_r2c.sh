#!/bin/bash
# Round 2, call C: tensor-core rounding probe, wide-layer timelines, sleep A/B on the wide presets, full GPU tests.
mkdir -p gpurun_out; O=gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -rf 2>&1 | tail -12
echo "== mma rounding"; timeout 300 python scripts/micro/mma_rounding.py > $O/r2c_mma_rounding.txt 2>&1; tail -40 $O/r2c_mma_rounding.txt
echo "== timeline c4 (E=37: one wave)"; TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c4 37 2 > $O/r2c_timeline_c4.txt 2>&1; tail -34 $O/r2c_timeline_c4.txt
echo "== timeline c3 (E=37)"; TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c3 37 2 > $O/r2c_timeline_c3.txt 2>&1; tail -34 $O/r2c_timeline_c3.txt
for ns in 0 200 1000; do
  echo "== c4 E=256 wide_sleep_ns=$ns"; TDMPC2_B200_WIDE_SLEEP_NS=$ns timeout 200 python scripts/profile_iter.py c4 256 3 | tail -2
done
for ns in 0 1000; do
  echo "== c3 E=1024 wide_sleep_ns=$ns"; TDMPC2_B200_WIDE_SLEEP_NS=$ns timeout 200 python scripts/profile_iter.py c3 1024 3 | tail -2
done
