#!/bin/bash
set -u
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_dataset.py -q 2>&1 | tail -4
timeout 600 python tools/bench_device_dataset.py 1024 2>&1 | tail -1
VMM_GIF_NO_SMEM_OUT=1 timeout 600 python tools/bench_device_dataset.py 1024 2>&1 | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('global-memory form: decode_kernels_ms', d['decode_kernels_ms'])"
