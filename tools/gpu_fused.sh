set -x
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -8
b() { timeout 600 python bench.py --steps 40 --warmup 5 --no-kernels --no-weak 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
b ACT_ON
b ACT_ON_again
timeout 600 python tools/step_timeline.py > gpurun_out/timeline_r2c.json 2> gpurun_out/timeline_err.txt; tail -3 gpurun_out/timeline_err.txt
