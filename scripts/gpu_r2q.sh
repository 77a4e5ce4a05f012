#!/bin/bash
# Round 2, call Q: ping-pong engine with warp-private staging + weight prefetch during act_ready waits; pixel-observation path.
mkdir -p gpurun_out; O=gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -rf 2>&1 | tail -8
it() { timeout 300 python scripts/profile_iter.py $1 $2 4 | tail -2 | tr '\n' ' '; echo; }
for rep in 1 2; do for eng in tcgen05pp tcgen05x2; do echo -n "[$rep] c2 $eng: "; TDMPC2_ENGINE=$eng it c2 256; done; done
echo "== pp timeline"; TDMPC2_ENGINE=tcgen05pp TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_prof.so TDMPC2_TRACE=pp TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c2 37 2 > $O/r02_timeline_pp.txt 2>&1; sed -n 1,2p $O/r02_timeline_pp.txt; sed -n 8,20p $O/r02_timeline_pp.txt
