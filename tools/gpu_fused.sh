mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rl_gpu.py -q -m gpu -k "advantage" 2>&1 | tail -3
ADV_G=72,12,24,48,96,128,9,100 ADV_LG=16,20 timeout 300 python tools/advantage_bench.py 2>&1 | cut -c1-200 | tee gpurun_out/advantage_bench_r2.jsonl
