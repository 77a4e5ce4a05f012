#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "cta_pair" -p no:cacheprovider -x 2>&1 | tail -15
