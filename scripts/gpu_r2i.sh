#!/bin/bash
# final round-2 evidence: GPU tests, launch list of the bench command, ncu captures of the new kernels, bench lines, parity report
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/pytest.log; tail -2 $O/pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
python scripts/launch_summary.py $O/launches_bench.csv 2 > $O/launches_summary.txt 2>&1; head -24 $O/launches_summary.txt
for k in agg_bf16:conv_scatter_kernel agg_res_bf16:conv_scatter_rm concat_ro:concat_ro cls_chain:conv_scatter_cls; do
  what=${k%%:*}; pat=${k##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -o $O/$what -f python scripts/prof_kernels.py $what > $O/ncu_$what.log 2>&1; tail -1 $O/ncu_$what.log
done
timeout 400 ncu --set full --clock-control none -k regex:conv_first_tc -s 8 -c 2 -o $O/first_tc -f python scripts/layer_times.py 64 bf16 > $O/ncu_first_tc.log 2>&1; tail -1 $O/ncu_first_tc.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python scripts/parity_report.py > $O/parity_report.txt 2>&1; tail -3 $O/parity_report.txt
timeout 300 python scripts/sass_summary.py > $O/sass_summary.txt 2>&1; tail -3 $O/sass_summary.txt
