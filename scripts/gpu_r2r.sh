#!/bin/bash
# Round 2, call R: ping-pong engine, warp-private staging; weight prefetch during act_ready waits (non-blocking polls) vs none.
mkdir -p gpurun_out; O=gpurun_out
echo "== pp tests"; timeout 900 python -m pytest tests/test_gpu_multitrip.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 600 -k "ping_pong or many_trip" -rf 2>&1 | tail -4
echo "== pp tests (prefetch build)"; TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_pf.so timeout 900 python -m pytest tests/test_gpu_multitrip.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 600 -k "ping_pong or many_trip" -rf 2>&1 | tail -4
it() { timeout 300 python scripts/profile_iter.py $1 $2 4 | tail -2 | tr '\n' ' '; echo; }
for rep in 1 2; do
  echo -n "[$rep] c2 pp default  : "; TDMPC2_ENGINE=tcgen05pp it c2 256
  echo -n "[$rep] c2 pp +prefetch : "; TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_pf.so TDMPC2_ENGINE=tcgen05pp it c2 256
  echo -n "[$rep] c2 x2: "; TDMPC2_ENGINE=tcgen05x2 it c2 256
done
echo "== pp timeline (prefetch)"; TDMPC2_ENGINE=tcgen05pp TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_prof.so TDMPC2_TRACE=pp TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c2 37 2 > $O/r02_timeline_pp.txt 2>&1; sed -n 1,2p $O/r02_timeline_pp.txt; sed -n 8,14p $O/r02_timeline_pp.txt
