set -x
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 2>gpurun_out/bench_n$N.err | tail -1 | tee gpurun_out/bench_n$N.json
tail -5 gpurun_out/bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 --scaling weak --no-kernels 2>/dev/null | tail -1 | tee gpurun_out/bench_n${N}_weak.json
