set -x
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -8
b() { timeout 600 python bench.py --steps 40 --warmup 5 --no-kernels --no-weak 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
b NAT_ON
b NAT_ON_again
RIFT_B200_PACK_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-kernels --no-weak 2> gpurun_out/pack_trace_raw.txt | tail -1 | cut -c1-100
grep '^PACK' gpurun_out/pack_trace_raw.txt | sort | uniq -c | sort -rn > gpurun_out/pack_trace2.txt
rm gpurun_out/pack_trace_raw.txt
