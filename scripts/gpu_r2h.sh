#!/bin/bash
# compute-sanitizer on the kernels added in the second half of round 2 (+ the generic engine, now two CTAs per SM)
O=gpurun_out/r2h; mkdir -p $O
for c in scatter_rm concat_ro igemm; do
  for t in racecheck synccheck memcheck; do
    timeout 600 compute-sanitizer --tool $t python scripts/sanitize_small.py $c > $O/sanitizer_${t}_$c.log 2>&1
    echo "$c $t: $(grep -c 'hazard detected\|Barrier error\|Invalid' $O/sanitizer_${t}_$c.log) reports; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $O/sanitizer_${t}_$c.log | tail -1)"
  done
done
