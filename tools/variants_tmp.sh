#!/bin/bash
O=gpurun_out/r01d; mkdir -p $O
for c in 1 2; do
  UVS_DENSE_CTAS_PER_SM=$c python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-blockdiag-extra > $O/bench_cta$c.json 2>$O/bench_cta$c.err
  python - <<PY
import json;d=json.load(open("$O/bench_cta$c.json"));print("ctas/sm $c",d["value"],d["roofline"]["frac"])
PY
done
UVS_B200_LIB=$PWD/uncalibrated-visual-servoing_b200/libuvs_b200_minb3.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_closed_loop -c 1 -o $O/dense_minb3 -f \
    python bench.py --steps 1 --warmup 1 --trials 14800 --no-cpu-baseline --no-blockdiag-extra > $O/ncu_minb3.log 2>&1
