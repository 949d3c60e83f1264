#!/bin/bash
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "quantile or sampler or cfg" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_full_parity.py -q -x -k "sampl or ddim or ancestral" 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 --no-strong --no-torch-gpu --no-cpu-baseline > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/i_bench.json") if l.startswith("{")][-1])
    print("ms/step", round(d["ms_per_step"], 2), "clips/s", round(d["value"], 1), "p_sample", d.get("p_sample"))
except Exception as e:
    print("bench unreadable:", e)
PY
