#!/bin/bash
# final round-2 evidence: GPU tests, launch list of the bench command, ncu captures of the new kernels, bench lines, SASS summary
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/pytest.log; tail -2 $O/pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
python scripts/launch_summary.py $O/launches_bench.csv 2 > $O/launches_summary.txt 2>&1; head -30 $O/launches_summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_scatter_rm -s 2 -c 1 -o $O/agg_rm -f python scripts/prof_kernels.py agg_res_bf16 > $O/ncu_rm.log 2>&1; tail -1 $O/ncu_rm.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:concat_ro -s 2 -c 1 -o $O/concat_ro -f python scripts/prof_kernels.py concat_ro > $O/ncu_ro.log 2>&1; tail -1 $O/ncu_ro.log
timeout 400 ncu --set full --clock-control none -k regex:conv_igemm -s 16 -c 8 -o $O/igemm -f python scripts/layer_times.py 64 bf16 > $O/ncu_igemm.log 2>&1; tail -1 $O/ncu_igemm.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python scripts/parity_report.py > $O/parity_report.txt 2>&1; tail -5 $O/parity_report.txt
