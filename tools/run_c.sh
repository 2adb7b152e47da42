#!/bin/bash
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline"
show() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1', 'e2e_ms', round(d['e2e']['ms_per_step'],4), 'equal', d['e2e'].get('host_records_equal_device_path'))
except Exception as e: print('$1', 'FAILED', e)"; }
$B 2>/dev/null | show default
for l in "32,64,128,192,192,192,96,32" "48,96,192,192,192,192,96,48" "32,48,96,128,128,128,64,32" "96,96,128,128,128,128,96,48" "96,96,96,96,96,96,64,32" "96,96,96,96,96,96,48,32"; do
ATLAS_RT_TRACE_MIN_BLOCKS_PER_SM=1 ATLAS_RT_PIPE_RPW=$l $B 2>/dev/null | show "minblocks1_rpw_$l"
done
ATLAS_RT_PIPE_RPW="96,96,96,96,96,96,64,32" $B 2>/dev/null | show "minblocks2_rpw_96,96,96,96,96,96,64,32"
