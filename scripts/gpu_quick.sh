#!/bin/bash
# tcgen05-engine parity tests + iteration timing + phase profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tcgen05" -p no:cacheprovider 2>&1 | tail -15
TDMPC2_PHASE_PROF=1 timeout 300 python scripts/profile_iter.py c2 37 2
timeout 300 python scripts/profile_iter.py c2 256 3
