#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/pytest.log; tail -4 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2o/bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('roofline', {k: d['roofline'][k] for k in ('frac', 'ms_per_launch', 'other_aggregation_ms', 'aggregation_share_of_step')})
print('roofline_hbm', d['roofline_hbm'])
print('chamfer', {k: v for k, v in d['stereo2point_chamfer'].items() if k in ('value', 'ms_per_step', 'chamfer_ms', 'chamfer_two_pass_ms', 'chamfer_frac_of_measured_fp32_issue_peak')})
print('fp32_mode', d['fp32_mode'])
print('latency', d['latency'])
PY
S3D_NO_SHEARED=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $O/bench_nosheared.json 2>> $O/bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2o/bench_nosheared.json')); print('no_sheared: value', d['value'], 'ms', d['ms_per_step'], d['roofline']['other_aggregation_ms'])"
