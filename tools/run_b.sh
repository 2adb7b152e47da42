#!/bin/bash
O=gpurun_out
timeout 500 python -m pytest tests -m gpu -q --tb=short -x > $O/pytest7.log 2>&1; tail -4 $O/pytest7.log
timeout 200 python tools/l2_sweep.py 2>&1 | grep -v "^oracle" | tee $O/r2_l2_sweep.txt
timeout 120 ./tools/divcheck > $O/r2_divcheck.txt 2>&1; cat $O/r2_divcheck.txt
