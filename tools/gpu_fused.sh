#!/bin/bash
# scratch script for one gpurun call (rewritten per experiment)
set -x
timeout 600 python -m pytest tests/test_buffer.py tests/test_plugin_gpu.py -m gpu -q 2>&1 | grep -E "FAILED|passed|failed" | tail -5
timeout 600 python bench.py --no-cpu --no-kernels 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('FULL', d['ms_per_step'], d['e2e']['ms_per_step'], d.get('e2e_device_buffer'))"
