set -x
TRACE_DBGS=0 TRACE_FEW=1 timeout 300 python tools/gemm_trace.py > gpurun_out/gemm_trace_r2.json 2>gpurun_out/gemm_trace_err.txt; tail -2 gpurun_out/gemm_trace_err.txt
b() { timeout 600 python bench.py --steps 40 --warmup 5 --no-kernels --no-weak --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
b V2ON
RIFT_B200_GEMM_V2=0 b V2OFF
b V2ON_again
