#!/bin/bash
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline"
show() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1', 'e2e_ms', round(d['e2e']['ms_per_step'],4), 'latency', round(d['e2e']['latency_ms_one_synchronous_call'],4), 'equal', d['e2e'].get('host_records_equal_device_path'))
except Exception as e: print('$1', 'FAILED', e)"; }
for ch in 4 6 8 12; do ATLAS_RT_PIPE_CHUNKS=$ch $B 2>/dev/null | show chunks$ch; done
ATLAS_RT_PIPE_CHUNKS=6 ATLAS_RT_TRACE_RAYS_PER_WARP=192 ATLAS_RT_TRACE_MIN_BLOCKS_PER_SM=1 $B 2>/dev/null | show chunks6_rpw192
ATLAS_RT_PIPE_CHUNKS=8 ATLAS_RT_TRACE_RAYS_PER_WARP=192 ATLAS_RT_TRACE_MIN_BLOCKS_PER_SM=1 $B 2>/dev/null | show chunks8_rpw192
