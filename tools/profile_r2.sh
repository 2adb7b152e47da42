#!/bin/bash
# Round-2 ncu evidence (one B200). Usage on the GPU box: bash tools/profile_r2.sh   -> gpurun_out/r2_*
O=gpurun_out
NCU="ncu --clock-control none"
PT='regex:trace_kernel|shade_|raygen|ray_cost|bin_count|bin_offsets|bin_scatter|set_words|next_bounce'
# launch lists (device time per launch; serialised, cold cache: compare shares)
timeout 400 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --profile > $O/r2_launches_bench.out 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum -k "$PT" -c 400 --csv --log-file $O/r2_launches_c5.csv python tools/prof_targets.py c5 > $O/r2_launches_c5.out 2>&1
ATLAS_BENCH_BINNING=1 timeout 300 $NCU --metrics gpu__time_duration.sum -k "$PT" -c 400 --csv --log-file $O/r2_launches_c5_binning.csv python tools/prof_targets.py c5 > $O/r2_launches_c5_binning.out 2>&1
# full captures of the dominant kernel: C2 (tree in L2) and the 8M terrain (tree 7x L2); third launch of each
timeout 300 $NCU --set full --import-source on -k regex:trace_kernel -s 2 -c 1 -o $O/r2_trace_c2 -f python tools/prof_targets.py c2 > $O/r2_trace_c2.out 2>&1
timeout 300 $NCU --set full --import-source on -k regex:trace_kernel -s 2 -c 1 -o $O/r2_trace_terrain -f python tools/prof_targets.py terrain > $O/r2_trace_terrain.out 2>&1
# the hit shader and the opacity-aware closest-hit trace of the second pass's first bounce
timeout 300 $NCU --set full --import-source on -k regex:shade_finish -s 5 -c 1 -o $O/r2_shade_finish -f python tools/prof_targets.py c5 > $O/r2_shade_finish.out 2>&1
timeout 300 $NCU --set full --import-source on -k regex:trace_kernel -s 10 -c 1 -o $O/r2_trace_c5_bounce0 -f python tools/prof_targets.py c5 > $O/r2_trace_c5_bounce0.out 2>&1
ls -la $O/r2_*
