#!/bin/bash
mkdir -p gpurun_out
Q="--no-cpu-baseline --eager-rays 0 --steps 3 --warmup 3"
for st in 1 2 3; do
  timeout 600 python bench.py $Q --streams $st > gpurun_out/bench_streams$st.log 2> gpurun_out/bench_streams$st.err
  echo "streams=$st rc=$? $(grep -E 'timed:' gpurun_out/bench_streams$st.err | tail -1)"
done
timeout 600 python bench.py $Q --ray-batch 8192 --streams 2 > gpurun_out/bench_rb8k_s2.log 2> gpurun_out/bench_rb8k_s2.err; echo "rb8192 s2 $(grep -E 'timed:' gpurun_out/bench_rb8k_s2.err | tail -1)"
timeout 600 python bench.py $Q --ray-batch 8192 --streams 4 > gpurun_out/bench_rb8k_s4.log 2> gpurun_out/bench_rb8k_s4.err; echo "rb8192 s4 $(grep -E 'timed:' gpurun_out/bench_rb8k_s4.err | tail -1)"
timeout 600 python bench.py $Q --rays 1024 --steps 20 > gpurun_out/bench_1024.log 2> gpurun_out/bench_1024.err; echo "1024 rays $(grep -E 'timed:' gpurun_out/bench_1024.err | tail -1)"
timeout 600 python bench.py $Q --rays 4096 --steps 10 > gpurun_out/bench_4096.log 2> gpurun_out/bench_4096.err; echo "4096 rays $(grep -E 'timed:' gpurun_out/bench_4096.err | tail -1)"
timeout 600 python bench.py $Q --mode infer > gpurun_out/bench_infer.log 2> gpurun_out/bench_infer.err; echo "infer rc=$? $(grep -E 'timed:' gpurun_out/bench_infer.err | tail -1)"; tail -3 gpurun_out/bench_infer.err | cut -c1-300
