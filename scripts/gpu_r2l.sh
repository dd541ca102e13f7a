#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
./scripts/f32x2_rate > $O/f32x2_rate.txt 2>&1; cat $O/f32x2_rate.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chamfer_sym_kernel -s 1 -c 1 -o $O/chamfer_sym -f python scripts/prof_kernels.py chamfer > $O/ncu_chamfer.log 2>&1; tail -1 $O/ncu_chamfer.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_tc -s 1 -c 1 -o $O/corr_tc -f python scripts/prof_kernels.py corr_tc > $O/ncu_corr.log 2>&1; tail -1 $O/ncu_corr.log
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "chamfer" 2>&1 | tail -3
