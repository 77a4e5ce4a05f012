#!/bin/bash
# Round 2, call O: L2 eviction hints on the wide models' operand loads (A planes evict-first, weights evict-last): same-box A/B.
it() { timeout 300 python scripts/profile_iter.py $1 $2 4 | tail -2 | tr '\n' ' '; echo; }
echo "== wide-model parity with hints"; TDMPC2_B200_L2HINT=1 timeout 600 python -m pytest tests/test_gpu_multitrip.py tests/test_gpu_golden.py -q -m gpu -p no:cacheprovider --timeout 600 -k "wide or 48 or 317 or c3 or c4 or golden" 2>&1 | tail -3
for rep in 1 2; do
  for h in 0 1; do echo -n "[$rep] c4 l2hint=$h: "; TDMPC2_B200_L2HINT=$h it c4 256; done
  for h in 0 1; do echo -n "[$rep] c3 l2hint=$h: "; TDMPC2_B200_L2HINT=$h it c3 1024; done
done
