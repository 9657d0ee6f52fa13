#!/bin/bash
O=gpurun_out/r01m; mkdir -p $O
for rep in 1 2; do
for v in "" _cs; do
  UVS_B200_LIB=$PWD/uncalibrated-visual-servoing_b200/libuvs_b200$v.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench$v.json 2>$O/bench$v.err
  python - <<PY
import json;d=json.load(open("$O/bench$v.json"));print("variant '$v'",d["value"],d["roofline"]["frac"],"bd",d["blockdiag"]["value"])
PY
done
done
