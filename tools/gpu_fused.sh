#!/bin/bash
# scratch script for one gpurun call (rewritten per experiment): GPU tests, smoke, one bench line
set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 40 --warmup 5 --no-kernels --no-weak --no-cpu 2>/dev/null | tail -1 | cut -c1-300
