#!/bin/bash
# Round 2, call I: de-phasing experiment (TDMPC2_B200_STAGGER), partial-grid TMA delivery rate, pair-mode prior rollouts.
mkdir -p gpurun_out; O=gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -rf 2>&1 | tail -6
echo "== tma_bw (partial grids)"; timeout 120 scripts/micro/tma_bw 2>&1 | tee $O/r02_micro_tma_bw.txt | tail -14
it() { timeout 200 python scripts/profile_iter.py $1 $2 5 | tail -3 | tr '\n' ' '; echo; }
for rep in 1 2; do
  for st in 0 12000 24000 36000 48000; do echo -n "[$rep] c2 stagger=$st: "; TDMPC2_B200_STAGGER=$st it c2 256; done
done
for st in 0 24000; do echo -n "c2 fast stagger=$st: "; TDMPC2_B200_PASSES=1 TDMPC2_B200_STAGGER=$st it c2 256; done
for st in 0 100000 300000; do echo -n "c4 stagger=$st: "; TDMPC2_B200_STAGGER=$st it c4 256; done
echo "== bench c2 stagger 0 / 24000 (clocks, power)"
for st in 0 24000; do TDMPC2_B200_STAGGER=$st timeout 600 python bench.py --steps 20 --warmup 3 --no-gpu-baseline --no-cpu-baseline --no-parity > $O/r02_bench_c2_stagger$st.json 2>/dev/null; python - <<PY
import json; d=json.load(open("$O/r02_bench_c2_stagger$st.json")); print("stagger $st:", round(d["ms_per_step"],3), "ms/plan, iter", round(d["roofline"]["ms_per_launch"],3), "outside", round(d["config"]["ms_outside_iter_kernels"],2), d["clocks"])
PY
done
