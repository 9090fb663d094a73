#!/bin/bash
# microbenchmark ceilings -> gpurun_out/micro_*.txt (copied into profiles/ by hand with the round tag)
mkdir -p gpurun_out
./scripts/micro/gather_bench > gpurun_out/micro_gather.json 2>&1; cat gpurun_out/micro_gather.json
./scripts/micro/red_bench > gpurun_out/micro_red.txt 2>&1; cat gpurun_out/micro_red.txt
./scripts/micro/plane_stage_bench > gpurun_out/micro_plane_stage.json 2>&1; cat gpurun_out/micro_plane_stage.json
timeout 600 python tests/probes/chamfer_probe.py > gpurun_out/chamfer_probe.log 2>&1; grep "chamfer\]" gpurun_out/chamfer_probe.log; tail -1 gpurun_out/chamfer_probe.log > gpurun_out/chamfer_vs_reference.json
