#!/bin/bash
# round 2, GPU call E: device dataset tests (after the Pillow-compat change), config-flag network tests, dataset bench row
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dataset.py -q > gpurun_out/e_dataset_tests.log 2>&1; echo "dataset tests rc=$?" | tee gpurun_out/e_summary.txt
timeout 900 python -m pytest tests/test_gpu_unet.py -q -s -k "config_flags" > gpurun_out/e_flag_tests.log 2>&1; echo "flag tests rc=$?" | tee -a gpurun_out/e_summary.txt
timeout 600 python tools/bench_device_dataset.py 1024 > gpurun_out/e_dataset_bench.log 2>&1; echo "dataset bench rc=$?" | tee -a gpurun_out/e_summary.txt
tail -c 1500 gpurun_out/e_dataset_tests.log
grep -a "config flags\|passed\|failed\|Error" gpurun_out/e_flag_tests.log | tail -12
tail -c 1200 gpurun_out/e_dataset_bench.log
