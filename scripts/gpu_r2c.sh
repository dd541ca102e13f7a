#!/bin/bash
# GPU session r2c: bench with all extras, sanitizer passes, ncu launch list
O=gpurun_out/r2c; mkdir -p $O
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2>> $O/bench.err
for tool in racecheck synccheck; do
  for c in scatter_pair scatter_single scatter_split concat cls_fused corr_tc igemm; do
    echo "=== $tool $c" >> $O/sanitizer_$tool.log
    timeout 240 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py $c 2>&1 | grep -v "^=========     \|^ran" | tail -12 >> $O/sanitizer_$tool.log
  done
done
tail -5 $O/sanitizer_racecheck.log $O/sanitizer_synccheck.log
cat $O/bench.json | head -c 6000
