#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python scripts/perf_probe.py 8,16 4096 2>&1 | tee gpurun_out/probe.log | grep -E "probe.*split-tc"
ARGS="--steps 1 --warmup 1 --rays 8192 --ray-batch 8192 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/ncu_launch_run.log 2>&1
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>10]
hdr=rows[0]; i_name=hdr.index('Kernel Name'); i_val=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    a=agg.setdefault(re.sub(r'\(.*','',r[i_name]),[0,0.0]); a[0]+=1; a[1]+=float(r[i_val].replace(',',''))/1e6
tot=sum(a[1] for a in agg.values())
for n,(c,ms) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:10]: print(f"{ms:10.2f} ms {100*ms/tot:6.2f}% x{c:4d} avg {ms/c:8.3f}  {n[:70]}")
PY
