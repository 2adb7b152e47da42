#!/bin/bash
# usage on the GPU box: bash tools/scale_r2.sh N     (N ranks of one node) -> gpurun_out/r2_scale_c2_N.json, r2_scale_c5_N.json
N=$1
O=gpurun_out
if [ "$N" = "1" ]; then
  timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2_scale_c2_1.json 2> $O/r2_scale_c2_1.err
  timeout 400 python bench.py --workload c5 --steps 16 --warmup 3 > $O/r2_scale_c5_1.json 2> $O/r2_scale_c5_1.err
else
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_scale_c2_$N.json 2> $O/r2_scale_c2_$N.err
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --workload c5 --steps 16 --warmup 3 > $O/r2_scale_c5_$N.json 2> $O/r2_scale_c5_$N.err
fi
python - <<PY
import json
for w in ("c2", "c5"):
    try:
        d = json.load(open("$O/r2_scale_%s_$N.json" % w))
        print(w, "N=$N", "value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "parity", d.get("parity"))
    except Exception as e:
        print(w, "N=$N FAILED", e)
PY
for f in $O/r2_scale_c2_$N.err $O/r2_scale_c5_$N.err; do tail -n 2 $f; done; true
