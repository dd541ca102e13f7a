#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/pytest_all.log; tail -5 $O/pytest_all.log
timeout 200 python scripts/layer_times.py 64 bf16 > $O/layers_two.txt 2>&1; grep "mrg\|forward" $O/layers_two.txt
S3D_SCATTER_ONE_CTA=1 timeout 200 python scripts/layer_times.py 64 bf16 > $O/layers_one.txt 2>&1; grep "mrg\|forward" $O/layers_one.txt
