#!/bin/bash
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name regex:qkv_bwd --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2b_full_qkv python tests/probe_qkvbwd.py > gpurun_out/r2b_full_qkv.log 2>&1
echo "ncu rc=$?"
tail -15 gpurun_out/r2b_full_qkv.log
ls -la gpurun_out/*.ncu-rep 2>/dev/null
if [ -f gpurun_out/r2b_full_qkv.ncu-rep ]; then
  ncu -i gpurun_out/r2b_full_qkv.ncu-rep --page raw --csv > gpurun_out/r2b_full_qkv_ln_bwd_L0.raw.csv 2>/dev/null
  python tools/ncu_stalls.py gpurun_out/r2b_full_qkv_ln_bwd_L0.raw.csv | head -24
  rm -f gpurun_out/r2b_full_qkv.ncu-rep
fi
