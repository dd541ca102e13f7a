#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_igemm.py -m gpu -q -x -k "ref_once" 2>&1 | tail -15 > $O/pytest_ro.log; tail -15 $O/pytest_ro.log
timeout 200 python scripts/layer_times.py 64 bf16 > $O/layers_ro.txt 2>&1; grep -i "concat\|forward\|dres\|cls" $O/layers_ro.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:concat_ro -s 2 -c 1 -o $O/concat_ro2 -f python scripts/prof_kernels.py concat_ro > $O/ncu_concat_ro.log 2>&1; tail -2 $O/ncu_concat_ro.log
