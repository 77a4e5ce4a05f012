#!/bin/bash
for WL in c4 c3; do
echo "== $WL a847fef"; ( cd _ab/a847fef && TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py $WL 37 2 2>&1 | sed -n 1,7p )
echo "== $WL eef2a6e"; ( cd _ab/eef2a6e && TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py $WL 37 2 2>&1 | sed -n 1,7p )
echo "== $WL current (prof lib)"; TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_prof.so TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py $WL 37 2 2>&1 | sed -n 1,7p
echo "== $WL current (product lib, time only)"; timeout 200 python scripts/profile_iter.py $WL 37 3 2>&1 | tail -2
done
