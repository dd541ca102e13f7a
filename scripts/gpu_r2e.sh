#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
cd stereo_3d_reconstruction_b200/csrc && make -s -j8 2>&1 | tail -3; cd ../..
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest.log; tail -4 $O/pytest.log
timeout 200 python scripts/graph_latency.py 1 2 8 > $O/latency.txt 2>&1; cat $O/latency.txt
# launch list of the bench command (per-launch times are cold-cache and serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
for k in agg_bf16x3 agg_res_bf16x3 corr_tc soft_argmin; do
  pat=conv_scatter; [ $k = corr_tc ] && pat=corr_tc; [ $k = soft_argmin ] && pat=soft_argmin
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -o $O/$k -f python scripts/prof_kernels.py $k > $O/ncu_$k.log 2>&1; tail -2 $O/ncu_$k.log
done
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print(d['value'], d['e2e']['value'], d['latency'], [ (p['C'],p['D'],round(p.get('softargmin_frac_hbm',0),2)) for p in d['costvolume_sweep']['points']])"
