#!/bin/bash
# kseg trade-off (accuracy vs time) of the wide presets on the lean kernels.
echo "== tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -rf 2>&1 | tail -6
for ks in 0 2048 1024; do
  echo "== c4 kseg=$ks"; TDMPC2_B200_KSEG=$ks timeout 300 python scripts/diag_wide_error.py c4 tcgen05x2 2>&1 | tail -1 | cut -c1-420
  TDMPC2_B200_KSEG=$ks timeout 200 python scripts/profile_iter.py c4 256 3 | tail -1
done
for ks in 0 1024 512; do
  echo "== c3 kseg=$ks"; TDMPC2_B200_KSEG=$ks timeout 300 python scripts/diag_wide_error.py c3 tcgen05x2 2>&1 | tail -1 | cut -c1-420
  TDMPC2_B200_KSEG=$ks timeout 200 python scripts/profile_iter.py c3 1024 3 | tail -1
done
