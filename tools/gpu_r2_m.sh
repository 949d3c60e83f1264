#!/bin/bash
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "fused_qkv_backward" 2>&1 | tail -15
timeout 120 python tests/probe_qkvbwd.py 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_kernels.py::test_fused_qkv_backward > gpurun_out/m_gpu_tests.log 2>&1; echo "gpu suite rc=$?"
tail -c 300 gpurun_out/m_gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-strong --no-torch-gpu --no-cpu-baseline --no-ddim > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/m_bench.json") if l.startswith("{")][-1])
    print("ms/step", round(d["ms_per_step"], 2), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "p_sample", d.get("p_sample", {}).get("value"))
    print("qkv_bwd_fused", d["roofline"]["other_kernels"].get("qkv_bwd_fused"))
except Exception as e:
    print("bench unreadable:", e)
PY
