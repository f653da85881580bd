#!/bin/bash
# scratch script for one gpurun call (rewritten per experiment)
set -x
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q 2>&1 | grep -E "FAILED|passed|failed" | tail -6
python tools/attn_probe.py
RIFT_B200_ATTN_SMALL=0 python tools/attn_probe.py 384 12 8 32
RIFT_B200_ATTN_SMALL=0 python tools/attn_probe.py 768 6 8 32
b() { timeout 300 python bench.py --steps 40 --warmup 5 --no-kernels --no-weak --no-cpu --no-devbuf 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
b SMALL
RIFT_B200_ATTN_SMALL=0 b NOSMALL
b SMALL_again
RIFT_B200_ATTN_SMALL=0 b NOSMALL_again
