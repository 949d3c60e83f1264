#!/bin/bash
# Round-2 profile on the GPU box (one GPU): launch list of eager training steps + ncu --set full captures of the top kernels.
# Outputs under gpurun_out/; tools/make_profiles.py r2 turns them into profiles/r2_*.md here.
tag=r2
export PYTHONPATH=.
mkdir -p gpurun_out
# (ncu occasionally ends with "application returned an error code (11)" right after start-up on this pool: every call retries)
for try in 1 2 3; do
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_train_launches.csv \
      python tests/probe_train_once.py 8 2 > gpurun_out/${tag}_ncu_launches.log 2>&1 && break
done
full() {  # name, regex, skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  for try in 1 2 3; do
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip $skip -c 1 -f -o gpurun_out/${tag}_full_${name} "$@" \
        > gpurun_out/${tag}_full_${name}.log 2>&1 && break
  done
  ncu -i gpurun_out/${tag}_full_${name}.ncu-rep --page raw --csv > gpurun_out/${tag}_full_${name}.raw.csv 2>/dev/null
}
full ftattn_fwd_sampling_L0 ftattn_fwd_kernel 1 python tests/probe_ftattn.py 8
full ftattn_fwd_training_L0 ftattn_fwd_kernel 0 python tests/probe_train_once.py 8 1
full cgemm_conv3x3_64x64_L0 cgemm 5 python tests/probe_cgemm.py perf1
full wgrad_conv3x3_64x64_L0 wgrad_kernel 2 python tests/probe_train_once.py 8 1
full tattn_bwd_L0 tattn_bwd 0 python tests/probe_train_once.py 8 1
full lattn_bwd_L0 lattn_bwd_mma 0 python tests/probe_train_once.py 8 1
full gn_silu_bwd_apply_L0 gn_silu_bwd_apply 0 python tests/probe_train_once.py 8 1
full gn_silu_bwd_reduce_L0 gn_silu_bwd_reduce 0 python tests/probe_train_once.py 8 1
full lattn_out_L0 lattn_out_mma 0 python tests/probe_train_once.py 8 1
full lattn_ctx_L0 lattn_ctx_mma 0 python tests/probe_train_once.py 8 1
ls -la gpurun_out | grep ${tag}_ | head -40
for try in 1 2 3; do
  timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'cgemm|wgrad_kernel|ftattn' --csv \
      --log-file gpurun_out/${tag}_gemm_dram.csv python tests/probe_train_once.py 8 1 > gpurun_out/${tag}_gemm_dram.log 2>&1 && break
done
# the .ncu-rep files are large: keep only the two fused-kernel reports for the source page, drop the others after the raw csv is made
for f in gpurun_out/${tag}_full_*.ncu-rep; do case $f in *ftattn_fwd_sampling*) ;; *) rm -f $f;; esac; done
du -sh gpurun_out
