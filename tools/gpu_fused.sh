#!/bin/bash
# scratch script for one gpurun call (rewritten per experiment): the round's final evidence run
set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "FAILED|passed|failed" | cut -c1-160 | tail -8
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; tail -3 gpurun_out/r2_bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null
timeout 600 python bench.py --trainable pi_head --no-cpu --no-kernels > gpurun_out/r2_bench_pi_head.json 2>/dev/null
timeout 600 python bench.py --workload cfg4 --no-cpu --no-kernels --steps 20 > gpurun_out/r2_bench_cfg4_1gpu.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_full_step.csv python bench.py --ncu --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu list rc=$?"
timeout 600 python tools/step_timeline.py > gpurun_out/timeline_final.json 2> gpurun_out/timeline_err.txt
python - <<'PY'
import json
for f in ("r2_bench_full", "r2_bench_reference_arm", "r2_bench_pi_head", "r2_bench_cfg4_1gpu"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, d.get("value"), d.get("ms_per_step"), d.get("e2e", {}).get("value"), d.get("gpu_launches"), d.get("cpu_baseline", {}).get("value"), d.get("e2e_device_buffer", {}).get("ms_per_step"))
    except Exception as e:
        print(f, "failed", e)
PY
