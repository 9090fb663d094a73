#!/bin/bash
# ncu evidence for profiles/: (1) launch list of a short bench run, (2) one --set full capture of the
# fused forward and backward kernels.  One GPU, never multi-rank.  Numbers printed under ncu are NOT bench values.
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --rays 8192 --ray-batch 8192 --no-cpu-baseline ${BENCH_EXTRA}"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/ncu_launch_run.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_render -s 2 -c 2 -f -o gpurun_out/prof \
    python bench.py $ARGS > gpurun_out/ncu_full_run.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep
