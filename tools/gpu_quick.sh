#!/bin/bash
# quick iteration: GPU parity suite + bench line (+ optional ncu of the dense kernel with NCU=1)
TAG=${1:-q}
O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python - <<PY
import json;d=json.load(open("$O/bench.json"));print("dense",d["value"],"frac",d["roofline"]["frac"],"bd",d.get("blockdiag",{}).get("value"),"e2e",d["e2e"]["value"])
PY
if [ -n "$NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_closed_loop -c 1 -o $O/dense_full -f \
    python bench.py --steps 1 --warmup 1 --trials 14800 --no-cpu-baseline --no-blockdiag-extra --no-extras > $O/ncu_dense.log 2>&1
fi
