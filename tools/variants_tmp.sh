#!/bin/bash
# usage: tools/variants_tmp.sh <tag> <variant suffixes...>   ("" = the default library)
TAG=$1; shift
O=gpurun_out/$TAG; mkdir -p $O
for rep in 1 2; do
for v in "$@"; do
  [ "$v" = "default" ] && sfx="" || sfx="_$v"
  UVS_B200_LIB=$PWD/uncalibrated-visual-servoing_b200/libuvs_b200$sfx.so python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-blockdiag-extra --no-extras > $O/bench$sfx.json 2>$O/bench$sfx.err
  python - <<PY
import json;d=json.load(open("$O/bench$sfx.json"));print("variant '$v'",d["value"],d["roofline"]["frac"])
PY
done
done
