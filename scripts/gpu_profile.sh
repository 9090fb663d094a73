#!/bin/bash
# launch list + one full capture of every kernel of the split/tensor-core pipeline (default bench config, L=16)
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --rays 8192 --ray-batch 8192 --no-cpu-baseline --eager-rays 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/ncu_launch_run.log 2>&1
echo "launch list rc=$?"
timeout 2400 ncu --set full --clock-control none --import-source on -k regex:"k_fwd_|k_bwd_|k_fold_" -s 13 -c 13 -f -o gpurun_out/prof \
    python bench.py $ARGS > gpurun_out/ncu_full_run.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep; grep -E "passes|Profiling" gpurun_out/ncu_full_run.log | head -12
