#!/bin/bash
# full ncu capture (with source) of the 64->64 3x3 cgemm at level 0 and of the fused qkv backward
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
cap() {  # name regex skip cmd...
  local name=$1 regex=$2 skip=$3; shift 3
  for try in 1 2 3; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip $skip -c 1 -f -o gpurun_out/r2b_full_${name} "$@" > gpurun_out/r2b_full_${name}.log 2>&1 && break
  done
  ncu -i gpurun_out/r2b_full_${name}.ncu-rep --page raw --csv > gpurun_out/r2b_full_${name}.raw.csv 2>/dev/null
  ncu -i gpurun_out/r2b_full_${name}.ncu-rep --page source --csv > gpurun_out/r2b_full_${name}.src.csv 2>/dev/null
  python tools/ncu_stalls.py gpurun_out/r2b_full_${name}.raw.csv 2>&1 | head -30
  rm -f gpurun_out/r2b_full_${name}.ncu-rep
}
cap cgemm_conv3x3_64x64_L0 cgemm 5 python tests/probe_cgemm.py perf1
cap qkv_ln_bwd_L0 qkv_bwd 2 python tests/probe_qkvbwd.py
ls -la gpurun_out | grep r2b_full
