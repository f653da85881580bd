#!/bin/bash
# one gpurun call: GPU parity tests, benches, ncu launch lists (profiling numbers are never bench values)
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full rc=$?"
cat gpurun_out/bench_full.json
timeout 600 python bench.py --steps 20 --warmup 5 --trainable pi_head > gpurun_out/bench_pi.json 2> gpurun_out/bench_pi.err; echo "bench pi rc=$?"
cat gpurun_out/bench_pi.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_full.csv python bench.py --ncu --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu list full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_pi.csv python bench.py --ncu --no-cpu --trainable pi_head > gpurun_out/ncu_pi.log 2>&1; echo "ncu list pi rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc -c 40 -f -o gpurun_out/prof_gemm_tc python bench.py --ncu --no-cpu > gpurun_out/ncu_gemm.log 2>&1; echo "ncu full gemm rc=$?"
ls -la gpurun_out
