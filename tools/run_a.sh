#!/bin/bash
# GPU-box batch: full GPU tests, host-buffer (streaming) sweeps, L2 window sweep, divcheck. Every step under its own timeout.
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short > $O/pytest6.log 2>&1; tail -4 $O/pytest6.log
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline"
show() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4), 'equal', d['e2e'].get('host_records_equal_device_path'))
except Exception as e: print('$1', 'FAILED', e)"; }
$B 2>$O/run_a_default.err | tee $O/bench_r2c.json | show default
ATLAS_RT_TRACE_STREAMING=0 $B 2>/dev/null | show chunked_pipeline
for c in 8 24 48; do ATLAS_RT_STREAM_CHUNKS=$c $B 2>/dev/null | show stream_chunks_$c; done
for mb in 32 64 96; do ATLAS_RT_L2_PERSIST_MB=$mb $B 2>/dev/null | show l2_persist_$mb; done
ATLAS_RT_L2_PERSIST_MB=64 ATLAS_RT_L2_HIT_RATIO=0.6 $B 2>/dev/null | show l2_persist_64_ratio0.6
timeout 120 ./tools/divcheck > $O/r2_divcheck.txt 2>&1; cat $O/r2_divcheck.txt
