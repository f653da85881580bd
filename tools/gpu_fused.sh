set -x
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -8
python tools/attn_probe.py
RIFT_B200_ATTN_FWD_TILED=0 python tools/attn_probe.py
b() { timeout 600 python bench.py --steps 40 --warmup 5 --no-kernels --no-weak 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
b TILED2
b TILED2_again
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention -c 2 -f -o gpurun_out/attn_r2b python tools/attn_probe.py 64 52 8 32 > /dev/null 2>&1
