#!/bin/bash
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
for v in 0 8 16 32 64; do
  echo "VMM_GN_REDUCE_MIN_VEC=$v"
  VMM_GN_REDUCE_MIN_VEC=$v timeout 300 python tests/probe_norm.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  hw', d['hw'], 'C', d['C'], 'gn_fwd', d['gn_fwd_us'], 'gn_bwd', d['gn_bwd_us'], 'ln_fwd', d['ln_fwd_us'], 'ln_bwd', d['ln_bwd_us'])
    elif 'Error' in l or 'error' in l: print(l.rstrip())
"
done
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "gn_ or norm" 2>&1 | tail -3
