#!/bin/bash
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_full_parity.py tests/test_gpu_trainer.py -q -x 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --no-strong --no-torch-gpu --no-cpu-baseline > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/v_bench.json") if l.startswith("{")][-1])
    print("ms/step", round(d["ms_per_step"], 2), "p_sample", d["p_sample"]["value"], "ddim", d["p_sample"]["ddim250_sample"]["seconds"])
except Exception as e:
    print("bench unreadable:", e); print(open("gpurun_out/v_bench.err").read()[-1500:])
PY
VMM_SIDE_COND=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-strong --no-torch-gpu --no-cpu-baseline > gpurun_out/v_bench0.json 2> gpurun_out/v_bench0.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/v_bench0.json") if l.startswith("{")][-1])
    print("VMM_SIDE_COND=0: p_sample", d["p_sample"]["value"], "ddim", d["p_sample"]["ddim250_sample"]["seconds"])
except Exception as e:
    print("bench0 unreadable:", e)
PY
