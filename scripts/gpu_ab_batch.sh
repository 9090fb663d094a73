#!/bin/bash
Q="--no-cpu-baseline --eager-rays 0 --steps 3 --warmup 3"
for rb in 16384 32768 65536; do
  timeout 600 python bench.py $Q --ray-batch $rb > /tmp/ab.log 2> /tmp/ab.err; echo "ray_batch=$rb $(grep -E 'timed:' /tmp/ab.err | tail -1 | cut -c18-) $(grep -oE 'e2e [0-9.]+ ms' /tmp/ab.err | tail -1) peak_mem $(python -c "import json;print(round(json.loads(open('/tmp/ab.log').read().strip().splitlines()[-1])['peak_mem_gb'],1))")"
done
