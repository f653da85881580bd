mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "tcgen05 or wgrad" 2>&1 | tail -3
for cfg in "lite:X=1" "lite_w96:RIFT_B200_WGRAD_CTAS=96" "lite_w296:RIFT_B200_WGRAD_CTAS=296"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --no-kernels > gpurun_out/q_$name.json 2> gpurun_out/q_$name.err
  tail -1 gpurun_out/q_$name.err | cut -c1-300
  python - "$name" <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/q_{f}.json")); print(f, "ms", round(d["ms_per_step"], 3), "mean", round(d["config"]["ms_per_step_mean"], 3), "launches", d["gpu_launches"])
except Exception as e:
    print(f, "failed", e)
PY
done
