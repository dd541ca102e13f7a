#!/bin/bash
O=gpurun_out/r2d; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_split.py -x -q -k "z_split" 2>&1 | tail -15 > $O/zsplit.log; tail -6 $O/zsplit.log
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/pytest.log; tail -5 $O/pytest.log
timeout 200 python scripts/graph_latency.py 1 2 4 8 > $O/latency.txt 2>&1; cat $O/latency.txt
timeout 100 python scripts/layer_times.py 1 bf16 > $O/layers_b1.txt 2>&1; cat $O/layers_b1.txt
timeout 100 python scripts/layer_times.py 64 bf16 u8 > $O/layers_b64_u8.txt 2>&1; grep "conv_first\|forward" $O/layers_b64_u8.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline_hbm']['frac'])"
