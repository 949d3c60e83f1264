#!/bin/bash
# Round profile on the GPU box: launch list of two eager training steps + ncu --set full captures of the top kernels.
# usage: tools/profile_round.sh TAG      (outputs under gpurun_out/, summaries are turned into profiles/*.md by tools/make_profiles.py)
tag=${1:-r1}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_train_launches.csv \
    python tests/probe_train_once.py 8 2 > gpurun_out/${tag}_ncu_launches.log 2>&1
full() {  # name, regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip $skip -c 1 -f -o gpurun_out/${tag}_full_${name} "$@" \
      > gpurun_out/${tag}_full_${name}.log 2>&1
  ncu -i gpurun_out/${tag}_full_${name}.ncu-rep --page raw --csv > gpurun_out/${tag}_full_${name}.raw.csv 2>/dev/null
}
full cgemm_conv3x3_64x64_L0 cgemm 5 python tests/probe_cgemm.py perf1
full cgemm_conv3x3_256x256_L2 cgemm 5 python tests/probe_cgemm.py perf256
full wgrad_first_L0 wgrad_kernel 2 python tests/probe_train_once.py 8 1
full tattn_bwd_L0 tattn_bwd 0 python tests/probe_train_once.py 8 1
full tattn_fwd_L0 tattn_fwd 1 python tests/probe_train_once.py 8 1
full lattn_bwd_L0 lattn_bwd_mma 0 python tests/probe_train_once.py 8 1
full gn_silu_bwd_apply_L0 gn_silu_bwd_apply 0 python tests/probe_train_once.py 8 1
full gn_silu_bwd_reduce_L0 gn_silu_bwd_reduce 0 python tests/probe_train_once.py 8 1
full gn_silu_fwd_L0 gn_silu_fwd 0 python tests/probe_train_once.py 8 1
ls gpurun_out | grep ${tag}_ | head -50
# DRAM traffic of every GEMM launch of one eager step (two metrics only: cheap), for roofline.traffic
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'cgemm|wgrad_kernel' --csv \
    --log-file gpurun_out/${tag}_gemm_dram.csv python tests/probe_train_once.py 8 1 > gpurun_out/${tag}_gemm_dram.log 2>&1
