#!/bin/bash
O=gpurun_out
NCU="ncu --clock-control none"
timeout 300 $NCU --set full --import-source on -k regex:trace_kernel -s 2 -c 1 -o $O/r2_trace_c2 -f python tools/prof_targets.py c2 > $O/r2_trace_c2.out 2>&1
timeout 300 $NCU --set full --import-source on -k regex:trace_kernel -s 2 -c 1 -o $O/r2_trace_terrain -f python tools/prof_targets.py terrain > $O/r2_trace_terrain.out 2>&1
ls -la $O/r2_trace_c2.ncu-rep $O/r2_trace_terrain.ncu-rep
