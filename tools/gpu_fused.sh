set -x
timeout 600 python -m pytest tests/test_evaluator.py tests/test_plugin_gpu.py -m gpu -q -x 2>&1 | tail -25
