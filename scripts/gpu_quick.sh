#!/bin/bash
# tcgen05-engine parity tests + iteration timing + phase profile (every step under a hard timeout)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tcgen05" -p no:cacheprovider 2>&1 | tail -15
TDMPC2_TRACE=1 TDMPC2_PHASE_PROF=1 timeout 120 python scripts/profile_iter.py c2 37 2 | tail -14
timeout 120 python scripts/profile_iter.py c2 256 3 | tail -2
