#!/bin/bash
# ping-pong engine: parity test, then timing against the CTA-pair engine (every step under a hard timeout)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ping_pong or cta_pair" -x -p no:cacheprovider 2>&1 | tail -15
TDMPC2_ENGINE=tcgen05pp timeout 120 python scripts/profile_iter.py c2 256 3 | tail -2
TDMPC2_ENGINE=tcgen05x2 timeout 120 python scripts/profile_iter.py c2 256 3 | tail -2
