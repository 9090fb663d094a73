#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chamfer.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python tests/probes/chamfer_probe.py 2>&1 | tee gpurun_out/chamfer_probe.log | grep "chamfer\]"
