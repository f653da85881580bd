set -x
N=${1:-2}
date
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 2>gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
echo "rc=$?"; date
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print('N=$N', d['value'], d['ms_per_step'], d['scaling'], d['e2e']['value'], d.get('weak_scaling'))"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
date
