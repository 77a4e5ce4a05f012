#!/bin/bash
# Round 2, call N: whole-warp role loops with elected issue (x2 + pp) -- full tests, timing of every workload, timelines.
mkdir -p gpurun_out; O=gpurun_out
echo "== tests (default engine)"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -rf 2>&1 | tail -6
echo "== tests on tcgen05pp"; TDMPC2_B200_ENGINE=tcgen05pp timeout 900 python -m pytest tests/test_gpu_multitrip.py tests/test_gpu_golden.py tests/test_gpu_edges.py -q -m gpu -p no:cacheprovider --timeout 600 -rf 2>&1 | tail -4
it() { timeout 300 python scripts/profile_iter.py $1 $2 4 | tail -2 | tr '\n' ' '; echo; }
for rep in 1 2; do
  for eng in tcgen05x2 tcgen05pp; do echo -n "[$rep] c2 $eng: "; TDMPC2_ENGINE=$eng it c2 256; done
  echo -n "[$rep] c3: "; it c3 1024
  echo -n "[$rep] c4: "; it c4 256
done
echo -n "c2 fast: "; TDMPC2_B200_PASSES=1 it c2 256
echo -n "c4 fast: "; TDMPC2_B200_PASSES=1 it c4 256
echo "== timeline c2 x2"; TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_prof.so TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c2 37 2 > $O/r02_timeline_c2_x2.txt 2>&1; sed -n 1,2p $O/r02_timeline_c2_x2.txt; sed -n 8,17p $O/r02_timeline_c2_x2.txt
echo "== timeline c4"; TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_prof.so TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 300 python scripts/profile_iter.py c4 37 2 > $O/r02_timeline_c4.txt 2>&1; sed -n 1,2p $O/r02_timeline_c4.txt; sed -n 8,13p $O/r02_timeline_c4.txt
