#!/bin/bash
Q="--no-cpu-baseline --eager-rays 0 --steps 3 --warmup 3 --streams 1"
run() { timeout 600 python bench.py $Q > /tmp/ab.log 2> /tmp/ab.err; echo "$1 $(grep -E 'timed:' /tmp/ab.err | tail -1 | cut -c18-) | $(grep -oE 'kernels .*' /tmp/ab.err | tail -1 | cut -c1-330)"; }
run new
BENCH_OLD_LOSS=1 run oldloss
