#!/bin/bash
# Round 2, call D: K-segmented heads (accuracy), TMA-prefetched normalise pass of the wide layers, WIDE template split.
mkdir -p gpurun_out; O=gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -rf 2>&1 | tail -12
echo "== error breakdown c4 / c3 (head_kseg default 512)"; timeout 300 python scripts/diag_wide_error.py c4 tcgen05x2 simt 2>&1 | tail -2; timeout 300 python scripts/diag_wide_error.py c3 tcgen05x2 2>&1 | tail -1
echo "== head_kseg = 0 (whole K) / 256"; TDMPC2_B200_HEAD_KSEG=0 timeout 300 python scripts/diag_wide_error.py c4 tcgen05x2 2>&1 | tail -1; TDMPC2_B200_HEAD_KSEG=256 timeout 300 python scripts/diag_wide_error.py c4 tcgen05x2 2>&1 | tail -1
echo "== head_kseg 512 + wide kseg 1024"; TDMPC2_B200_KSEG=1024 timeout 300 python scripts/diag_wide_error.py c4 tcgen05x2 2>&1 | tail -1
echo "== timeline c4 (E=37)"; TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c4 37 2 > $O/r2d_timeline_c4.txt 2>&1; sed -n 1,14p $O/r2d_timeline_c4.txt
echo "== timeline c3 (E=37)"; TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c3 37 2 > $O/r2d_timeline_c3.txt 2>&1; sed -n 1,14p $O/r2d_timeline_c3.txt
echo "== c4 E=256"; timeout 200 python scripts/profile_iter.py c4 256 3 | tail -2
echo "== c3 E=1024"; timeout 200 python scripts/profile_iter.py c3 1024 3 | tail -2
echo "== c2 E=256"; timeout 200 python scripts/profile_iter.py c2 256 4 | tail -3
