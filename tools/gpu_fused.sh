#!/bin/bash
# scratch script for one gpurun call (rewritten per experiment): final sanity of the committed state
set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|passed|failed" | cut -c1-160 | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('FINAL', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['e2e_device_buffer']['ms_per_step'], d['cpu_baseline']['value'])"
