#!/bin/bash
# Multi-GPU leg of a round (gpurun --gpus N): the G-invariance test on hardware + bench.py at N ranks.
# usage: tools/gpu_scale.sh <tag> <N>
TAG=${1:-r02s}; N=${2:-8}
O=gpurun_out/$TAG; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $O/nvidia-smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.log 2>&1; tail -3 $O/pytest_multi.log
for n in $(seq 1 $N); do
  case $n in 1|2|4|8)
    [ $n -gt $N ] && continue
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-blockdiag-extra > $O/bench_n1.json 2> $O/bench_n1.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
          bench.py --gpus $n --steps 5 --warmup 3 --no-extras --no-blockdiag-extra > $O/bench_n$n.json 2> $O/bench_n$n.err
    fi
    python - <<PY
import json
try:
    d=json.load(open("$O/bench_n$n.json")); print("N=$n", "%.4g"%d["value"], d["scaling"], "ms/step %.2f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], d["config"]["trials_per_gpu"])
except Exception as e: print("N=$n FAILED", e)
PY
  ;; esac
done
