#!/bin/bash
# Round 2, call L: how fast can the kernels' own TMA request streams ingest operands?  Build without MMAs (the commits
# free the slots at once): the GEMM phases then last exactly as long as the ingest.  Timelines of CTA 0, one wave.
mkdir -p gpurun_out; O=gpurun_out
L=/root/repo/tdmpc2_b200/libtdmpc2_b200_nomma.so
for eng in tcgen05x2 tcgen05 ; do
  echo "== no-MMA build, engine $eng"; TDMPC2_ENGINE=$eng TDMPC2_B200_LIB=$L TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c2 37 2 > $O/r02_nomma_$eng.txt 2>&1; sed -n 1,14p $O/r02_nomma_$eng.txt
done
echo "== no-MMA build, engine tcgen05pp"; TDMPC2_ENGINE=tcgen05pp TDMPC2_B200_LIB=$L TDMPC2_TRACE=pp TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c2 37 2 > $O/r02_nomma_pp.txt 2>&1; sed -n 1,14p $O/r02_nomma_pp.txt
echo "== regular prof build, engine tcgen05pp"; TDMPC2_ENGINE=tcgen05pp TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_prof.so TDMPC2_TRACE=pp TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c2 37 2 > $O/r02_timeline_pp.txt 2>&1; sed -n 1,12p $O/r02_timeline_pp.txt
echo "== pp vs x2 per iteration (E=256)"; for eng in tcgen05pp tcgen05x2; do echo -n "$eng: "; TDMPC2_ENGINE=$eng timeout 200 python scripts/profile_iter.py c2 256 4 | tail -2 | tr '\n' ' '; echo; done
echo "== no-MMA c4 (wide), one wave"; TDMPC2_B200_LIB=$L TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 300 python scripts/profile_iter.py c4 37 2 > $O/r02_nomma_c4.txt 2>&1; sed -n 1,12p $O/r02_nomma_c4.txt
echo "== regular c4 timeline"; TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_prof.so TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 300 python scripts/profile_iter.py c4 37 2 > $O/r02_timeline_c4.txt 2>&1; sed -n 1,12p $O/r02_timeline_c4.txt
