#!/bin/bash
# ncu captures of the end of round 2 (third session; one B200, under gpurun): --set full of the top kernels, the
# launch list of one bench step, the IFT level launches.  The .ncu-rep files stay in /tmp on the box; what comes
# back are the text summaries (tools/ncu_raw_summary.py, ncu_launch_summary.py, ncu_ift_levels.py) for profiles/.
set -x
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg5 --no-batch --no-parity"
$B > $O/r03_plain.json 2> $O/r03_plain.err || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'knn_cell_kernel|normals_kernel|rg_classify|rg_hook_kernel|knn_general_kernel|graph_fill_rev|rg_sizes|gather_sorted' --launch-skip 36 --launch-count 9 -f -o /tmp/prof_r03_top $B > /tmp/ncu_b.log 2>&1
python tools/ncu_raw_summary.py /tmp/prof_r03_top.ncu-rep "ncu --set full --clock-control none, bench.py cfg3 (10M facade, k=32), last step, end of round 2: k-NN fast path (both launches), general path, normals, region-growing classify (ballot form) + hook, segment sizes, reverse-arc fill, Morton gather" > $O/r03_top_kernels_ncu_full.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file $O/r03_launches.csv $B > /tmp/ncu_a.log 2>&1
python tools/ncu_launch_summary.py $O/r03_launches.csv $O/r03_launches.txt
[ -n "$SKIP_IFT" ] && exit 0
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex_op_atom_dot_alu.sum,lts__t_requests_srcunit_tex_op_atom_dot_cas.sum,smsp__inst_executed.sum --clock-control none -k regex:ift_level_kernel --launch-skip 908 --launch-count 227 --csv --log-file /tmp/ift_levels_r03.csv $B > /tmp/ncu_c.log 2>&1
python tools/ncu_ift_levels.py /tmp/ift_levels_r03.csv 227 $O/r03_ift_levels_ncu.txt
tail -3 /tmp/ncu_a.log /tmp/ncu_b.log /tmp/ncu_c.log
grep -c "==PROF==" /tmp/ncu_b.log
