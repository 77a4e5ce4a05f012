#!/bin/bash
# Same-box A/B of library versions staged under _ab/<commit>/ against the working tree: per-iteration time of the wide presets.
run() { ( cd $1 && shift && "$@" timeout 200 python scripts/profile_iter.py $WL $E 4 | tail -2 | tr '\n' ' '; echo ); }
echo "== tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -rf 2>&1 | tail -8
for rep in 1 2; do
for WL_E in "c3 1024" "c4 256" "c2 256"; do
  set -- $WL_E; WL=$1; E=$2
  for v in _ab/a847fef .; do echo -n "[$rep] $WL $v: "; run $v env; done
done
done
echo "== timeline c4 (prof build)"; TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_prof.so TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c4 37 2 > gpurun_out/r2e_timeline_c4.txt 2>&1; sed -n 1,12p gpurun_out/r2e_timeline_c4.txt
echo "== timeline c3 (prof build)"; TDMPC2_B200_LIB=/root/repo/tdmpc2_b200/libtdmpc2_b200_prof.so TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 200 python scripts/profile_iter.py c3 37 2 > gpurun_out/r2e_timeline_c3.txt 2>&1; sed -n 1,12p gpurun_out/r2e_timeline_c3.txt
