#!/bin/bash
# One gpurun call of round 2.  usage: tools/gpu_r2.sh <tag> [steps...]   steps: test bench probe ncu variants
TAG=${1:-r02x}; shift
STEPS=${@:-test bench}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/nvidia-smi.txt 2>&1
for s in $STEPS; do
case $s in
test)
    timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
    tail -15 $O/pytest_gpu.log ;;
testall)
    timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
    tail -40 $O/pytest_gpu.log ;;
bench)
    timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
    cat $O/bench.json ;;
probe)
    timeout 120 tools/dmma_probe > $O/dmma_probe.txt 2>&1; cat $O/dmma_probe.txt ;;
variants)
    # every libuvs_b200*.so built in-tree: the headline number only
    for lib in uncalibrated-visual-servoing_b200/libuvs_b200*.so; do
        UVS_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-blockdiag-extra \
            > $O/bench_$(basename $lib .so).json 2> $O/bench_$(basename $lib .so).err
        python - <<PY
import json
try:
    d=json.load(open("$O/bench_$(basename $lib .so).json")); print("$(basename $lib)", "%.4g"%d["value"], "kernel_ms %.2f"%d["roofline"]["kernel_ms"])
except Exception as e: print("$(basename $lib)", "FAILED", e)
PY
    done ;;
bdvariants)
    for lib in uncalibrated-visual-servoing_b200/libuvs_b200*.so; do
        UVS_B200_LIB=$PWD/$lib timeout 300 python bench.py --mode blockdiag --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-blockdiag-extra \
            > $O/bd_$(basename $lib .so).json 2> $O/bd_$(basename $lib .so).err
        python - <<PY
import json
try:
    d=json.load(open("$O/bd_$(basename $lib .so).json")); print("blockdiag $(basename $lib)", "%.4g"%d["value"], "kernel_ms %.2f"%d["roofline"]["kernel_ms"])
except Exception as e: print("$(basename $lib)", "FAILED", e)
PY
    done ;;
ncu)
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dense_closed_loop|dense_quad" -c 1 -o $O/dense_full -f \
        python bench.py --steps 1 --warmup 1 --trials 14800 --no-cpu-baseline --no-blockdiag-extra --no-extras > $O/ncu_dense.log 2>&1
    timeout 900 ncu --set full --clock-control none -k regex:"dense_closed_loop|dense_quad" -c 1 -o $O/dense_full_100k -f \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-blockdiag-extra --no-extras > $O/ncu_dense_100k.log 2>&1
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:bd_closed_loop -c 1 -o $O/bd_full -f \
        python bench.py --steps 1 --warmup 1 --mode blockdiag --no-cpu-baseline --no-blockdiag-extra --no-extras > $O/ncu_bd.log 2>&1
    # the reports are too large to travel back (64 MiB cap): summarise here, keep the summaries only
    python tools/profile_summary.py $TAG ${TAG} $O/profiles > $O/profile_summary.log 2>&1; tail -2 $O/profile_summary.log
    ls -la $O | head -30
    rm -f $O/*.ncu-rep ;;
esac
done
