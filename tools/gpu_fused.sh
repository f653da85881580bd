set -x
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -4
b() { timeout 300 python bench.py --steps 40 --warmup 5 --no-kernels --no-weak --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
b LATE
RIFT_B200_PDL_LATE=0 b NOLATE
b LATE_again
RIFT_B200_PDL_LATE=0 b NOLATE_again
