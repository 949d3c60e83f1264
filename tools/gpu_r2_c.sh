#!/bin/bash
set -u
export PYTHONPATH=.
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "wgrad or dgrad or grads" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_unet.py -q -k "wrap_mode or training or graph_replayed or differentiable" 2>&1 | tail -5
for v in 1 0; do
  if [ $v = 1 ]; then export VMM_WGRAD_NO_SLAB=1; else unset VMM_WGRAD_NO_SLAB; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-ddim --no-strong --no-torch-gpu --no-cpu-baseline > gpurun_out/c_bench_$v.json 2> gpurun_out/c_bench_$v.err
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/c_bench_$v.json") if l.startswith("{")][-1])
print("NO_SLAB=$v ms/step", round(d["ms_per_step"], 2), "wgrad", d["roofline"]["wgrad"])
PY
done
