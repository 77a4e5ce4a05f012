#!/bin/bash
# Round 2, call M: ping-pong engine with deeper weight ring -- parity (single trip + many trips), timing, timelines.
mkdir -p gpurun_out; O=gpurun_out
echo "== pp parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider --timeout 300 -k "ping_pong" -rf 2>&1 | tail -5
echo "== many-trip + edge + golden tests on tcgen05pp"; TDMPC2_B200_ENGINE=tcgen05pp timeout 900 python -m pytest tests/test_gpu_multitrip.py tests/test_gpu_golden.py tests/test_gpu_edges.py -q -m gpu -p no:cacheprovider --timeout 600 -rf -x 2>&1 | tail -8
echo "== pp vs x2 per iteration (E=256)"; for rep in 1 2; do for eng in tcgen05pp tcgen05x2; do echo -n "$eng: "; TDMPC2_ENGINE=$eng timeout 200 python scripts/profile_iter.py c2 256 4 | tail -2 | tr '\n' ' '; echo; done; done
echo "== pp timeline"; TDMPC2_ENGINE=tcgen05pp TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_prof.so TDMPC2_TRACE=pp TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c2 37 2 > $O/r02_timeline_pp.txt 2>&1; sed -n 1,2p $O/r02_timeline_pp.txt; sed -n 8,20p $O/r02_timeline_pp.txt
echo "== pp no-MMA timeline"; TDMPC2_ENGINE=tcgen05pp TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_nomma.so TDMPC2_TRACE=pp TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c2 37 2 > $O/r02_nomma_pp.txt 2>&1; sed -n 8,14p $O/r02_nomma_pp.txt
