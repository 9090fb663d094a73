#!/bin/bash
# evidence for DESIGN.md 2(5): the same kernels / rays against the oracle with white-noise and with band-limited tables
mkdir -p gpurun_out
timeout 600 python scripts/diag_tc.py ref_full_L16_interior > /dev/null 2>&1; cp gpurun_out/diag_tc.txt gpurun_out/parity_diag_white.txt
SMOOTH=1 timeout 600 python scripts/diag_tc.py ref_full_L16_interior > /dev/null 2>&1; cp gpurun_out/diag_tc.txt gpurun_out/parity_diag_smooth.txt
grep -E "####|hash_static|grid_enc" gpurun_out/parity_diag_white.txt gpurun_out/parity_diag_smooth.txt | head -40
