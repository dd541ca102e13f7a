#!/bin/bash
# symmetric Chamfer + lean corr_tc epilogue: tests, timings, e2e diagnosis
O=gpurun_out/r2k; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_golden.py tests/test_gpu_baseline_configs.py -m gpu -q -x -k "chamfer or corr" 2>&1 | tail -6 > $O/pytest_ch.log; cat $O/pytest_ch.log
timeout 600 python - > $O/extras.txt 2>&1 <<'PY'
import json, torch, bench
from config import cfg
dev = torch.device('cuda', 0)
pm = bench.extra_peaks(dev)
print(json.dumps(pm))
r = bench.extra_stereo2point(cfg, dev, pm.get('fp32_fma_per_s'))
print(json.dumps(r))
s = bench.extra_costvolume_sweep(dev, 6551.0)
for p in s['points']:
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in p.items()}))
PY
cat $O/extras.txt
