#!/bin/bash
Q="--no-cpu-baseline --eager-rays 0 --steps 3 --warmup 3"
run() { timeout 600 python bench.py $Q "$@" > /tmp/ab.log 2> /tmp/ab.err; echo "$* $(grep -E 'timed:' /tmp/ab.err | tail -1 | cut -c18-) | $(grep -oE 'k_adam[^,]*' /tmp/ab.err | tail -1)"; }
run --loss main
run --loss unit
(cd _ab_old && sed -i 's/BENCH_NOTHING//' bench.py && timeout 600 python bench.py $Q --streams 1 > /tmp/ab.log 2> /tmp/ab.err; echo "old(unit loss, old adam) $(grep -E 'timed:' /tmp/ab.err | tail -1 | cut -c18-)")
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q --no-header -k "adam" 2>&1 | tail -2
