mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -k "baseline_shape" -s 2>&1 | grep -oE "\[cfg[^]]*\].*" | cut -c1-700
timeout 600 python -m pytest tests/test_plugin_gpu.py tests/test_model_gpu.py -q -m gpu -k "get_action or micro_batched" 2>&1 | grep -E "^E  |passed|failed" | cut -c1-400 | head -20
timeout 900 python -m pytest tests/test_buffer.py -q -m gpu 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-400 | head -20
