#!/bin/bash
# Round-2 ncu evidence (one B200). Usage on the GPU box: bash tools/profile_r2.sh   -> gpurun_out/r2_*
set -x
O=gpurun_out
NCU="ncu --clock-control none"
# launch lists (device time per launch; serialised, cold cache: compare shares)
$NCU --metrics gpu__time_duration.sum -c 800 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --profile > $O/r2_launches_bench.out 2>&1
$NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $O/r2_launches_c5.csv python tools/prof_targets.py c5 > $O/r2_launches_c5.out 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r2_launches_e2e.csv python tools/prof_targets.py e2e > $O/r2_launches_e2e.out 2>&1
# full captures of the dominant kernel: C2 (tree in L2) and the 8M terrain (tree 7x L2); third launch of each
$NCU --set full --import-source on -k regex:trace_kernel -s 2 -c 1 -o $O/r2_trace_c2 -f python tools/prof_targets.py c2 > $O/r2_trace_c2.out 2>&1
$NCU --set full --import-source on -k regex:trace_kernel -s 2 -c 1 -o $O/r2_trace_terrain -f python tools/prof_targets.py terrain > $O/r2_trace_terrain.out 2>&1
# the hit shader of the first bounce of the second pass
$NCU --set full --import-source on -k regex:shade_finish -s 5 -c 1 -o $O/r2_shade_finish -f python tools/prof_targets.py c5 > $O/r2_shade_finish.out 2>&1
ls -la $O/r2_*
