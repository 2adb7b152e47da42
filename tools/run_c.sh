#!/bin/bash
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -q --tb=short -x > $O/pytest9.log 2>&1; tail -4 $O/pytest9.log
NCU="ncu --clock-control none"
PT='regex:trace_kernel|shade_|raygen|ray_cost|bin_count|bin_offsets|bin_scatter|set_words|next_bounce|add_accum'
timeout 400 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --profile > $O/r2_launches_bench.out 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum -k "$PT" -c 600 --csv --log-file $O/r2_launches_c5.csv python tools/prof_targets.py c5 > $O/r2_launches_c5.out 2>&1
python tools/ncu_summary.py launches $O/r2_launches_bench.csv | head -30
python tools/ncu_summary.py launches $O/r2_launches_c5.csv | head -20
