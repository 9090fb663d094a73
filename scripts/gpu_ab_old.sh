#!/bin/bash
mkdir -p gpurun_out
Q="--no-cpu-baseline --eager-rays 0 --steps 3 --warmup 3 --streams 1"
for v in new old new old; do
  if [ $v = old ]; then d=_ab_old; else d=.; fi
  (cd $d && timeout 600 python bench.py $Q > /tmp/ab_$v.log 2> /tmp/ab_$v.err)
  echo "$v $(grep -E 'timed:' /tmp/ab_$v.err | tail -1) $(grep -oE 'e2e [0-9.]+ ms' /tmp/ab_$v.err | tail -1)"
done
