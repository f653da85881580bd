set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "wgrad" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -5
b() { timeout 300 python bench.py --steps 40 --warmup 5 --no-kernels --no-weak --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
b GROUP
RIFT_B200_WGRAD_GROUP=0 b NOGROUP
RIFT_B200_WGRAD_GROUP_ITEMS=2 b GROUP_I2
RIFT_B200_WGRAD_GROUP_ITEMS=8 b GROUP_I8
b GROUP_again
