#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
for rep in 1 2 3; do
for k in NONE S3D_NO_CLS_CHAIN; do env $k=1 timeout 400 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $O/bench_$k.json 2> $O/bench_$k.err; python -c "
import json; d=json.load(open('$O/bench_$k.json')); print('$k', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['roofline']['ms_per_launch'],3), d['roofline']['other_aggregation_ms'], d['roofline_hbm'] and round(d['roofline_hbm']['ms_per_launch'],3))"; done; done
nvidia-smi --query-gpu=power.draw,power.limit,clocks.sm,temperature.gpu --format=csv
