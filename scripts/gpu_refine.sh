timeout 600 python tests/probes/refine_probe.py > gpurun_out/refine_probe.log 2>&1; grep "refine\]" gpurun_out/refine_probe.log; tail -1 gpurun_out/refine_probe.log > gpurun_out/refine_probe.json
