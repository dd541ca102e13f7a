#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; tail -2 $O/bench_n2.err
python -c "
import json; d=json.load(open('$O/bench_n2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('strong_512'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; tail -c 600 $O/bench_ref_n2.json
