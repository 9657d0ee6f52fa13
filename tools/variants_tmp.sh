#!/bin/bash
O=gpurun_out/r01j; mkdir -p $O
for v in "" _bd3 _bd4; do
  UVS_B200_LIB=$PWD/uncalibrated-visual-servoing_b200/libuvs_b200$v.so python bench.py --steps 3 --warmup 3 --mode blockdiag --no-cpu-baseline --no-blockdiag-extra > $O/bench$v.json 2>$O/bench$v.err
  python - <<PY
import json;d=json.load(open("$O/bench$v.json"));print("variant '$v'",d["value"],"e2e",d["e2e"]["value"])
PY
done
