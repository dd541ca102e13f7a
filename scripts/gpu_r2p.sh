#!/bin/bash
# end-of-round evidence: sanitizer on the new kernels, launch list, ncu captures, bench lines, parity report, SASS summary
O=gpurun_out/r2p; mkdir -p $O
for c in sheared chamfer_sym corr_tc; do
  for t in racecheck synccheck memcheck; do
    timeout 600 compute-sanitizer --tool $t python scripts/sanitize_small.py $c > $O/sanitizer_${t}_$c.log 2>&1
    echo "$c $t: $(grep -c 'hazard detected\|Barrier error\|Invalid' $O/sanitizer_${t}_$c.log) reports; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $O/sanitizer_${t}_$c.log | tail -1)"
  done
done
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/pytest.log; tail -2 $O/pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
python scripts/launch_summary.py $O/launches_bench.csv 2 > $O/launches_summary.txt 2>&1; head -26 $O/launches_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:map_conv -s 8 -c 1 -o $O/map_conv5 -f python scripts/sheared_time.py > $O/ncu_mc.log 2>&1; tail -1 $O/ncu_mc.log
timeout 300 ncu --set full --clock-control none -k regex:gonce_assemble -s 2 -c 1 -o $O/assemble -f python scripts/sheared_time.py > $O/ncu_as.log 2>&1; tail -1 $O/ncu_as.log
timeout 300 ncu --set full --clock-control none -k regex:chamfer_sym_kernel -s 1 -c 1 -o $O/chamfer_sym -f python scripts/prof_kernels.py chamfer > $O/ncu_ch.log 2>&1; tail -1 $O/ncu_ch.log
timeout 300 ncu --set full --clock-control none -k regex:corr_tc -s 1 -c 1 -o $O/corr_tc -f python scripts/prof_kernels.py corr_tc > $O/ncu_corr.log 2>&1; tail -1 $O/ncu_corr.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python scripts/parity_report.py > $O/parity_report.txt 2>&1; tail -3 $O/parity_report.txt
timeout 300 python scripts/sass_summary.py > $O/sass_summary.txt 2>&1; tail -2 $O/sass_summary.txt
