#!/bin/bash
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_dataset.py -q -x 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-strong --no-torch-gpu --no-cpu-baseline --no-ddim > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/u_bench.json") if l.startswith("{")][-1])
    print("ms/step", round(d["ms_per_step"], 2), "clips/s", round(d["value"], 1), "e2e", d["e2e"])
except Exception as e:
    print("bench unreadable:", e); print(open("gpurun_out/u_bench.err").read()[-1500:])
PY
