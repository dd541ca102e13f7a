#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 200 python scripts/layer_times.py 64 bf16 > $O/layers_ig2.txt 2>&1; grep "enc2\|rec\|dec\|forward" $O/layers_ig2.txt
