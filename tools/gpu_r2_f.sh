#!/bin/bash
# round 2, GPU call F: ncu launch list of an eager training step (per-launch durations of every kernel)
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
for try in 1 2 3; do
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_train_launches.csv \
      python tests/probe_train_once.py 8 2 > gpurun_out/r2b_ncu_launches.log 2>&1 && break
done
python tools/summarize_launches.py gpurun_out/r2b_train_launches.csv 2 | head -50
timeout 300 python -m pytest tests/test_gpu_dataset.py -q -k trainer 2>&1 | tail -3
