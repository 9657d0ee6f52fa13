#!/bin/bash
O=gpurun_out/r01f; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
for v in "" _w3r224; do
  UVS_B200_LIB=$PWD/uncalibrated-visual-servoing_b200/libuvs_b200$v.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-blockdiag-extra > $O/bench$v.json 2>$O/bench$v.err
  python - <<PY
import json;d=json.load(open("$O/bench$v.json"));print("variant '$v'",d["value"],d["roofline"]["frac"])
PY
done
