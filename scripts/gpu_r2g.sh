#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
python - <<'PY'
import json, torch, bench
pk = bench.peaks()
r = bench.extra_costvolume_sweep(torch.device('cuda:0'), pk['hbm_gbs'])
for p in r['points']:
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in p.items()})
PY
