#!/bin/bash
# usage: bash tools/run_c.sh N   -- C2 weak scaling with the NCCL gather and with the peer-memory window
O=gpurun_out
N=$1
for mode in nccl peer; do
ATLAS_BENCH_GATHER=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $O/r2_peer_${mode}_$N.json 2> $O/r2_peer_${mode}_$N.err
python - <<PY
import json
try:
    d = json.loads(open("$O/r2_peer_${mode}_$N.json").read().strip().splitlines()[-1])
    print("$mode N=$N value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "parity", d.get("parity"), "launches", d.get("gpu_launches"))
except Exception as e:
    print("$mode N=$N FAILED", e)
PY
tail -n 3 $O/r2_peer_${mode}_$N.err | cut -c1-300
done
