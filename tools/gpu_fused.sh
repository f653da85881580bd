mkdir -p gpurun_out
timeout 300 python tools/step_timeline.py > gpurun_out/timeline.json 2> gpurun_out/timeline.err; tail -3 gpurun_out/timeline.err; head -60 gpurun_out/timeline.json
