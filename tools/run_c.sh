#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_pathtrace.py tests/test_gpu_callers.py -m gpu -q --tb=short -x 2>&1 | tail -3
for parts in 1 2 4 8; do
for lanes in 1 4; do ATLAS_RT_PT_LANES=$lanes timeout 200 python tools/c5_shard_time.py $parts 2>&1 | tail -1; done
done
ATLAS_RT_PT_LANES=3 timeout 200 python tools/c5_shard_time.py 1 2>&1 | tail -1
