#!/bin/bash
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
for try in 1 2 3 4; do
timeout 560 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread \
  --clock-control none -k regex:qkv_bwd --csv --log-file gpurun_out/r2c_qkv_bwd_metrics.csv python tests/probe_qkvbwd.py > gpurun_out/r2c_qkv_bwd_ncu.log 2>&1 && break
echo "ncu try $try rc=$?"
done
tail -3 gpurun_out/r2c_qkv_bwd_ncu.log
wc -l gpurun_out/r2c_qkv_bwd_metrics.csv
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2c_qkv_bwd_metrics.csv")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
idx = {h: i for i, h in enumerate(rows[hi])}
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(rows[hi]): continue
    per.setdefault((r[idx["ID"]], r[idx["Kernel Name"]][:40]), {})[r[idx["Metric Name"]]] = (r[idx["Metric Value"]], r[idx["Metric Unit"]])
for k, m in list(per.items())[-4:]:
    print(k, m)
PY
