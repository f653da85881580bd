set -x
b() { timeout 600 python bench.py --steps 40 --warmup 5 --no-kernels --no-weak 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
b TALLCAP
b TALLCAP_again
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
