#!/bin/bash
# microbenchmark ceilings -> gpurun_out/micro_*.txt (copied into profiles/ by hand with the round tag)
mkdir -p gpurun_out
./scripts/micro/gather_bench > gpurun_out/micro_gather.json 2>&1; cat gpurun_out/micro_gather.json
./scripts/micro/red_bench > gpurun_out/micro_red.txt 2>&1; cat gpurun_out/micro_red.txt
