#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/pytest_all.log; tail -5 $O/pytest_all.log
timeout 300 python scripts/layer_times.py 64 bf16x3 > $O/layers_x3b.txt 2>&1; grep "concat\|dres0a\|enc5\|forward" $O/layers_x3b.txt
