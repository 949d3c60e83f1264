#!/bin/bash
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
for try in 1 2 3; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_sample_launches.csv \
      python tests/probe_sample_once.py 4 > gpurun_out/r2b_ncu_sample.log 2>&1 && break
done
python tools/summarize_launches.py gpurun_out/r2b_sample_launches.csv 3 2>/dev/null | head -45
tail -3 gpurun_out/r2b_ncu_sample.log
