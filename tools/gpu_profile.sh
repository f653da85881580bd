#!/bin/bash
# round evidence: bench lines (both arms), ncu launch list of one policy update, ncu --set full of the dominant GEMM
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full rc=$?"
timeout 600 python bench.py --trainable pi_head > gpurun_out/bench_pi.json 2> gpurun_out/bench_pi.err; echo "bench pi rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_full.csv python bench.py --ncu --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 2 -f -o gpurun_out/prof_gemm_big python tools/gemm_probe.py 46080 256 256 5 > gpurun_out/ncu_gemm_big.log 2>&1; echo "ncu full gemm rc=$?"
timeout 300 python tools/gemm_trace.py > gpurun_out/gemm_trace_final.json 2>/dev/null; echo "trace rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
