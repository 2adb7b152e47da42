#!/bin/bash
ATLAS_RT_BUILD_WIDE=1 timeout 500 python -m pytest tests/test_gpu_build.py -m gpu -q --tb=short -x 2>&1 | tail -3
for w in 1; do ATLAS_RT_BUILD_WIDE=$w timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline 2>&1 | grep "build samples"; done
ATLAS_RT_BUILD_WIDE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wide_level -c 200 --csv --log-file gpurun_out/wide_launches.csv python tools/prof_targets.py c2 > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/wide_launches.csv | head -6
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/wide_launches.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[h]; ki,vi=H.index("Kernel Name"),H.index("Metric Value")
out=[]
for r in rows[h+1:]:
    if len(r)>vi:
        n=r[ki]
        cls="c" if "<16" in n else ("g" if "<12" in n else "t")
        out.append((cls,float(r[vi].replace(",",""))/1000))
print(" ".join(f"{c}{t:.0f}" for c,t in out[:75]))
PY
