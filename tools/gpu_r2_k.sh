#!/bin/bash
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "fused_qkv_backward" 2>&1 | tail -15
timeout 120 python tests/probe_qkvbwd.py 2>&1 | tail -3
