set -x
b() { timeout 600 python bench.py --steps 40 --warmup 5 --no-kernels --no-weak 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
b CAP24
RIFT_B200_WGRAD_SPLIT_CAP=16 b CAP16
RIFT_B200_WGRAD_SPLIT_CAP=12 b CAP12
RIFT_B200_WGRAD_SPLIT_CAP=8 b CAP8
RIFT_B200_WGRAD_SPLIT_CAP=4 b CAP4
b CAP24_again
RIFT_B200_WGRAD_CTAS=64 b WAVE64
RIFT_B200_ATTN_BWD_TILED=0 b BWD_UNTILED
