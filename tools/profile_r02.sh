#!/bin/bash
# ncu captures of round 2 (one B200, under gpurun): launch list of the default bench step, --set full of the
# top kernels, the IFT level launches, the shard.cu kernels.  Reports land in gpurun_out/ (scratch); the
# summaries made from them go to profiles/ (tools/ncu_launch_summary.py, ncu_raw_summary.py, ncu_ift_levels.py).
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg5 --no-batch"
$B > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_r02_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'knn_cell_kernel|normals_kernel|rg_classify_kernel|rg_hook_kernel|knn_general_kernel|graph_fill_rev' --launch-skip 32 --launch-count 8 -f -o gpurun_out/prof_r02_top $B > gpurun_out/ncu_r02_b.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex_op_atom_dot_alu.sum,lts__t_requests_srcunit_tex_op_atom_dot_cas.sum,smsp__inst_executed.sum --clock-control none -k regex:ift_level_kernel --launch-skip 908 --launch-count 227 --csv --log-file gpurun_out/ift_levels_r02.csv $B > gpurun_out/ncu_r02_c.log 2>&1
python tools/profile_shards.py > gpurun_out/r02_shards_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'shard_' --launch-skip 40 --launch-count 80 -f -o gpurun_out/prof_r02_shards python tools/profile_shards.py > gpurun_out/ncu_r02_d.log 2>&1
