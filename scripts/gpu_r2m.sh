#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_baseline_configs.py -m gpu -q -x -k "corr" 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chamfer|corr_tc" --csv --log-file $O/times.csv python scripts/prof_kernels.py chamfer > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chamfer|corr_tc" --csv --log-file $O/times2.csv python scripts/prof_kernels.py corr_tc > /dev/null 2>&1
grep -h '^"' $O/times.csv $O/times2.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | head -20
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chamfer_sym_kernel -s 1 -c 1 -o $O/chamfer_sym -f python scripts/prof_kernels.py chamfer > $O/ncu_chamfer.log 2>&1; tail -1 $O/ncu_chamfer.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_tc -s 1 -c 1 -o $O/corr_tc -f python scripts/prof_kernels.py corr_tc > $O/ncu_corr.log 2>&1; tail -1 $O/ncu_corr.log
timeout 300 python - <<'PY'
import json, torch, bench
s = bench.extra_costvolume_sweep(torch.device('cuda', 0), 6551.0)
for p in s['points']:
    print({k: round(v, 4) for k, v in p.items() if k.startswith('corr') or k in 'CD'})
PY
