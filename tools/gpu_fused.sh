mkdir -p gpurun_out
RIFT_B200_FUSED=0 timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -k "baseline_shape" 2>&1 | grep -E "AssertionError|passed|failed|gmax" | cut -c1-1500
timeout 600 python -m pytest tests/test_plugin_gpu.py -q -m gpu -k "get_action" 2>&1 | tail -25
