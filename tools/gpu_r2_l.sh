#!/bin/bash
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/l_gpu_tests.log 2>&1; echo "gpu suite rc=$?"
tail -c 600 gpurun_out/l_gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-strong --no-torch-gpu --no-cpu-baseline --no-ddim > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/l_bench.json") if l.startswith("{")][-1])
    print("ms/step", round(d["ms_per_step"], 2), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "p_sample", d.get("p_sample", {}).get("value"))
    print("qkv_bwd_fused", d["roofline"]["other_kernels"].get("qkv_bwd_fused"))
    print("wgrad", d["roofline"]["wgrad"], "cgemm frac", d["roofline"]["frac"], "other", d["roofline"]["other_gemm"])
except Exception as e:
    print("bench unreadable:", e)
PY
