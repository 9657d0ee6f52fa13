#!/bin/bash
# One gpurun call: GPU parity suite, bench line, ncu launch list, ncu --set full of the dense closed-loop kernel.
# usage: tools/gpu_round.sh <tag>
TAG=${1:-r01x}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi > $O/nvidia-smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
cat $O/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dense_closed_loop|dense_quad" -c 1 -o $O/dense_full -f \
    python bench.py --steps 1 --warmup 1 --trials 14800 --no-cpu-baseline --no-blockdiag-extra --no-extras > $O/ncu_dense.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"dense_closed_loop|dense_quad" -c 1 -o $O/dense_full_100k -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-blockdiag-extra --no-extras > $O/ncu_dense_100k.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bd_closed_loop -c 1 -o $O/bd_full -f \
    python bench.py --steps 1 --warmup 1 --mode blockdiag --no-cpu-baseline --no-blockdiag-extra --no-extras > $O/ncu_bd.log 2>&1
ls -la $O
timeout 900 ncu --set full --clock-control none -k regex:twist_closed_loop -c 1 -o $O/twist_full -f \
    python bench.py --steps 1 --warmup 1 --trials 14800 --no-cpu-baseline --no-blockdiag-extra > $O/ncu_twist.log 2>&1
