#!/bin/bash
# final evidence of the round: GPU tests, launch list, bench lines, sanitizer on the last new kernel
O=gpurun_out/r2r; mkdir -p $O
for t in racecheck synccheck memcheck; do
  timeout 600 compute-sanitizer --tool $t python scripts/sanitize_small.py halo2d > $O/sanitizer_${t}_halo2d.log 2>&1
  echo "halo2d $t: $(grep -c 'hazard detected\|Barrier error\|Invalid' $O/sanitizer_${t}_halo2d.log) reports; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $O/sanitizer_${t}_halo2d.log | tail -1)"
done
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/pytest.log; tail -2 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
python scripts/launch_summary.py $O/launches_bench.csv 2 > $O/launches_summary.txt 2>&1; head -24 $O/launches_summary.txt
timeout 300 ncu --set full --clock-control none -k regex:conv2d_halo -s 2 -c 1 -o $O/conv2d_halo -f python scripts/layer_times.py 64 bf16 > $O/ncu_halo.log 2>&1; tail -1 $O/ncu_halo.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python scripts/parity_report.py > $O/parity_report.txt 2>&1; tail -2 $O/parity_report.txt
timeout 300 python scripts/sass_summary.py > $O/sass_summary.txt 2>&1; tail -1 $O/sass_summary.txt
timeout 200 python scripts/layer_times.py 64 bf16 > $O/layers.txt 2>&1; tail -1 $O/layers.txt
