set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "wgrad" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -5
b() { timeout 300 python bench.py --steps 40 --warmup 5 --no-kernels --no-weak --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
b G12
RIFT_B200_WGRAD_GROUP_SIZE=6 b G6
RIFT_B200_WGRAD_GROUP_SIZE=10 b G10
RIFT_B200_WGRAD_GROUP_SIZE=3 b G3
b G12_again
