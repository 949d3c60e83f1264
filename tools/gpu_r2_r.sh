#!/bin/bash
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_full_parity.py tests/test_gpu_trainer.py -q -x -k "sampl or ddim or ancestral or guid or train_evaluate or p_sample" 2>&1 | tail -4
tail -3 gpurun_out/full_parity_report.jsonl 2>/dev/null | cut -c1-400
timeout 900 python bench.py --steps 5 --warmup 3 --no-strong --no-torch-gpu --no-cpu-baseline > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r_bench.json") if l.startswith("{")][-1])
    print("ms/step", round(d["ms_per_step"], 2), "p_sample", d.get("p_sample"))
except Exception as e:
    print("bench unreadable:", e)
PY
