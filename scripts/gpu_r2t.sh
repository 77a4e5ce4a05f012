#!/bin/bash
# Round 2, call T: evict-last hint on the fused epilogue's activation-plane TMA stores (TDMPC2_B200_L2HINT=2): DRAM write-back,
# iteration time, sustained bench -- same box A/B on the CTA-pair engine at c2.
mkdir -p gpurun_out; O=gpurun_out
echo "== parity with the hint"; TDMPC2_B200_L2HINT=2 timeout 600 python -m pytest tests/test_gpu_multitrip.py -q -m gpu -p no:cacheprovider --timeout 600 -k "many_trip and x2" 2>&1 | tail -2
it() { timeout 300 python scripts/profile_iter.py $1 $2 4 | tail -2 | tr '\n' ' '; echo; }
for rep in 1 2; do for h in 0 2; do echo -n "[$rep] c2 x2 l2hint=$h: "; TDMPC2_ENGINE=tcgen05x2 TDMPC2_B200_L2HINT=$h it c2 256; done; done
for h in 0 2; do
  echo "== dram bytes, l2hint=$h"; TDMPC2_ENGINE=tcgen05x2 TDMPC2_B200_L2HINT=$h timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:plan_kernel -s 3 -c 1 python scripts/profile_iter.py c2 256 3 2>&1 | grep -E "dram__bytes|gpu__time" 
done
for h in 0 2 0 2; do
  TDMPC2_B200_L2HINT=$h timeout 600 python bench.py --engine tcgen05x2 --steps 20 --warmup 3 --no-gpu-baseline --no-cpu-baseline --no-parity > $O/r02_bench_c2_hint$h.json 2>/dev/null
  python - <<PY
import json; d=json.load(open("$O/r02_bench_c2_hint$h.json")); print("bench l2hint=$h:", round(d["ms_per_step"],3), "ms/plan, iter", round(d["roofline"]["ms_per_launch"],3), d["clocks"])
PY
done
