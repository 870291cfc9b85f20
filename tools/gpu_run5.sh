#!/bin/bash
python -m pytest tests/test_gpu_pbr.py -x -q -m gpu 2>&1 | tail -2
B2_PIPE_DEBUG=1 python tools/e2e_probe.py 2>&1 | grep -E "pipe|e2e" | tail -6
python tools/e2e_probe.py 2>&1 | grep -E "e2e"
