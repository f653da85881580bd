#!/bin/bash
# scratch script for one gpurun call (rewritten per experiment)
set -x
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_plugin_gpu.py -m gpu -q -x 2>&1 | tail -4
b() { timeout 300 python bench.py --steps 40 --warmup 5 --no-kernels --no-weak --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
b Q2
RIFT_B200_WGRAD_GROUP_SIZE=6 b Q2_G6
RIFT_B200_WGRAD_GROUP_SIZE=8 b Q2_G8
b Q2_again
