#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/pytest_all.log; tail -5 $O/pytest_all.log
for rep in 1 2; do for k in NONE S3D_NO_CONV_FIRST_TC; do env $k=1 timeout 400 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline > $O/bench_$k.json 2> $O/bench_$k.err; python -c "
import json; d=json.load(open('$O/bench_$k.json')); print('$k', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"; done; done
