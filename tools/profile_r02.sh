#!/bin/bash
# ncu captures of round 2 (one B200, under gpurun): launch list of the default bench step, --set full of the
# top kernels, the IFT level launches, the shard.cu kernels.  The .ncu-rep files stay in /tmp on the box (gpurun
# brings back at most 64 MiB); what comes back are the text summaries made from them on the spot
# (tools/ncu_launch_summary.py, ncu_raw_summary.py, ncu_ift_levels.py), which are then copied to profiles/.
set -x
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg5 --no-batch"
$B > $O/r02_plain.json 2> $O/r02_plain.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file $O/r02_launches.csv $B > /tmp/ncu_a.log 2>&1
python tools/ncu_launch_summary.py $O/r02_launches.csv $O/r02_launches.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'knn_cell_kernel|normals_kernel|rg_classify_kernel|rg_hook_kernel|knn_general_kernel|graph_fill_rev' --launch-skip 32 --launch-count 8 -f -o /tmp/prof_r02_top $B > /tmp/ncu_b.log 2>&1
python tools/ncu_raw_summary.py /tmp/prof_r02_top.ncu-rep "ncu --set full --clock-control none, bench.py cfg3 (10M facade, k=32), last step: k-NN fast path (both launches), general path, normals, region-growing classify + hook, reverse-arc fill" > $O/r02_top_kernels_ncu_full.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex_op_atom_dot_alu.sum,lts__t_requests_srcunit_tex_op_atom_dot_cas.sum,smsp__inst_executed.sum --clock-control none -k regex:ift_level_kernel --launch-skip 908 --launch-count 227 --csv --log-file /tmp/ift_levels_r02.csv $B > /tmp/ncu_c.log 2>&1
python tools/ncu_ift_levels.py /tmp/ift_levels_r02.csv 227 $O/r02_ift_levels_ncu.txt
python tools/profile_shards.py --points 4000000 > $O/r02_shards_plain.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none -k regex:'shard_' --launch-skip 64 --launch-count 44 -f -o /tmp/prof_r02_shards python tools/profile_shards.py --points 4000000 > /tmp/ncu_d.log 2>&1
python tools/ncu_raw_summary.py /tmp/prof_r02_shards.ncu-rep "ncu --set full --clock-control none, tools/profile_shards.py: 4 shards time-sharing one B200, 4M-point facade, k=16 — the kernels of csrc/shard.cu (second call: no radius estimate)" > $O/r02_shard_kernels_ncu_full.txt
tail -3 /tmp/ncu_a.log /tmp/ncu_b.log /tmp/ncu_c.log /tmp/ncu_d.log
du -sh $O
