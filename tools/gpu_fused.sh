set -x
b() { timeout 300 python bench.py --steps 40 --warmup 5 --no-kernels --no-weak --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
b BASE
RIFT_B200_DEBUG_SKIP_WGRAD=1 b SKIPW
b BASE_again
