#!/bin/bash
# usage: tools/ncu_kernels.sh tag kernel_regex[:skip] ...   -> gpurun_out/ncu_<tag>_<kernel>.ncu-rep (+ raw csv)
tag=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  k=${spec%%:*}; skip=0; [[ "$spec" == *:* ]] && skip=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip -c 1 -f -o gpurun_out/ncu_${tag}_${k} \
     python tests/probe_train_once.py 8 1 > gpurun_out/ncu_${tag}_${k}.log 2>&1
  ncu -i gpurun_out/ncu_${tag}_${k}.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_${k}.raw.csv 2>/dev/null
done
