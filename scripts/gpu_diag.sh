#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python scripts/${DIAG:-diag_fullsize.py} "$@" > gpurun_out/diag.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/diag.log
