#!/bin/bash
mkdir -p gpurun_out
Q="--no-cpu-baseline --eager-rays 0 --steps 3 --warmup 3 --streams 1"
run() { timeout 600 python bench.py $Q > /tmp/ab.log 2> /tmp/ab.err; echo "$1 $(grep -E 'timed:' /tmp/ab.err | tail -1) $(grep -oE 'e2e [0-9.]+ ms' /tmp/ab.err | tail -1)"; }
run new
BENCH_OLD_LOSS=1 run oldloss
L4D_NO_MEMO=1 run nomemo
BENCH_OLD_LOSS=1 L4D_NO_MEMO=1 run oldloss+nomemo
(cd _ab_old && timeout 600 python bench.py $Q > /tmp/ab.log 2> /tmp/ab.err; echo "old $(grep -E 'timed:' /tmp/ab.err | tail -1)")
