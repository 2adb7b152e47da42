#!/bin/bash
for bps in 7 6 4; do
ATLAS_RT_TRACE_STREAMING=1 ATLAS_RT_STREAM_BLOCKS_PER_SM=$bps ATLAS_RT_STREAM_CHUNKS=8 ATLAS_RT_PIPE_TIMELINE=1 timeout 200 python tools/prof_targets.py e2e 2>&1 | grep timeline | tail -2
done
