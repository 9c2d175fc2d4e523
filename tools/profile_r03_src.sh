#!/bin/bash
# Source-level ncu capture (round 2, third session): per-SASS-instruction counters of the k-NN fast path,
# the normals kernel and the region-growing classify kernel of one cfg3 step; the source pages come back as CSV
# (the .ncu-rep stays on the box) and are attributed to source lines here with tools/ncu_by_line.py.
set -x
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg5 --no-batch --no-parity"
$B > $O/r03_plain.json 2> $O/r03_plain.err || exit 1
for K in knn_cell_kernel normals_kernel rg_classify_kernel rg_hook_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 4 --launch-count 1 -f -o /tmp/src_$K $B > /tmp/ncu_$K.log 2>&1
  ncu -i /tmp/src_$K.ncu-rep --page source --csv > $O/r03_src_$K.csv 2> /dev/null
  ncu -i /tmp/src_$K.ncu-rep --page raw --csv > $O/r03_raw_$K.csv 2> /dev/null
  tail -2 /tmp/ncu_$K.log
done
du -sh $O
