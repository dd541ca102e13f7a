#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_full.json 2> $O/bench_full.err; tail -3 $O/bench_full.err; python -c "
import json; d=json.load(open('$O/bench_full.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])
print(json.dumps(d['roofline'], indent=0)[:1500])
print(d['roofline_hbm'])
print(d.get('fp32_mode')); print(d.get('latency')); print(d.get('gpu_stock_baseline'))"
