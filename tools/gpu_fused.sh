#!/bin/bash
# scratch script for one gpurun call (rewritten per experiment)
set -x
for t in 2 1; do
RIFT_B200_WGRAD_TERMS=$t timeout 900 python -m pytest tests/test_model_gpu.py tests/test_plugin_gpu.py -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|norm dev" | cut -c1-330 | tail -16
done
b() { timeout 300 python bench.py --steps 40 --warmup 5 --no-kernels --no-weak --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
b T3
RIFT_B200_WGRAD_TERMS=2 b T2
RIFT_B200_WGRAD_TERMS=1 b T1
b T3_again
