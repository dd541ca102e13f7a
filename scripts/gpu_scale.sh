#!/bin/bash
# bench.py on N GPUs of one box the way the driver launches it (torchrun, 127.0.0.1); usage: gpu_scale.sh N
N=${1:-2}; O=gpurun_out/scale; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err
grep '^{' $O/bench_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'strong_512', d.get('strong_512',{}).get('value'))"
