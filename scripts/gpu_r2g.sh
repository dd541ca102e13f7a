#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > $O/pytest_all.log; tail -4 $O/pytest_all.log
timeout 200 python scripts/layer_times.py 64 bf16 > $O/layers_final.txt 2>&1; cat $O/layers_final.txt
