#!/bin/bash
for parts in 1 4 8; do
for lanes in 4 6 8; do ATLAS_RT_PT_LANES=$lanes timeout 200 python tools/c5_shard_time.py $parts 2>&1 | tail -1; done
done
