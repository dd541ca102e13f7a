#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_scatter_cls -s 2 -c 1 -o $O/cls_chain -f python scripts/prof_kernels.py cls_chain > $O/ncu_chain.log 2>&1; tail -1 $O/ncu_chain.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_scatter_cls|partials_soft" --csv python scripts/prof_kernels.py cls_chain 2>&1 | grep -i "gpu__time" | tail -6
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_scatter_kernel" --csv python scripts/prof_kernels.py agg_bf16 2>&1 | grep -i "gpu__time" | tail -3
