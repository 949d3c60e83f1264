#!/bin/bash
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "flattn or fused_linear or linear_attention" 2>&1 | tail -5
timeout 300 python tests/probe_flattn.py 2>&1 | tail -12
timeout 900 python bench.py --steps 5 --warmup 3 --no-strong --no-torch-gpu --no-cpu-baseline --no-ddim > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/j_bench.json") if l.startswith("{")][-1])
    print("ms/step", round(d["ms_per_step"], 2), "clips/s", round(d["value"], 1), "p_sample", d.get("p_sample", {}).get("value"))
except Exception as e:
    print("bench unreadable:", e)
PY
