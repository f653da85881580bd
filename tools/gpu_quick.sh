#!/bin/bash
# tests + benches + (optional) launch list (profiling numbers are never bench values)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/q_full.json 2> gpurun_out/q_full.err; tail -2 gpurun_out/q_full.err
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-kernels --trainable pi_head > gpurun_out/q_pi.json 2> gpurun_out/q_pi.err
RIFT_B200_WGRAD_ATOMIC=0 RIFT_B200_LN_ATOMIC=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-kernels > gpurun_out/q_full_noatomic.json 2> gpurun_out/q_full_noatomic.err
python - <<'PY'
import json
for f in ("q_full", "q_pi", "q_full_noatomic"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "launches", d["gpu_launches"], "step frac", d["roofline_step"]["frac"])
        for k, v in d.get("kernels", {}).items(): print("   ", k, round(v["kernel_ms"] * 1e3, 1), "us", round(v["frac_hbm"], 3))
    except Exception as e:
        print(f, "failed", e)
PY
if [ "$1" = "list" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_full.csv python bench.py --ncu --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu list rc=$?"
fi
