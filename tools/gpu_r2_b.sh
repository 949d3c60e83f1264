#!/bin/bash
# round 2, GPU call B: whole -m gpu suite (fused temporal attention on), smoke(), bench with all arms.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/full_parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/b_gpu_tests.log 2>&1; echo "gpu suite rc=$?" | tee gpurun_out/b_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/b_summary.txt
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?" | tee -a gpurun_out/b_summary.txt
grep -E "passed|failed" gpurun_out/b_gpu_tests.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/b_gpu_tests.log | head -20
tail -3 gpurun_out/b_smoke.log
python - <<'PY'
import json
try:
    rows = [l for l in open("gpurun_out/b_bench.json") if l.startswith("{")]
    d = json.loads(rows[-1])
    print("ms/step", round(d["ms_per_step"], 2), "clips/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "p_sample", (d.get("p_sample") or {}),
          "\nstrong", (d.get("config4_strong_scaling") or {}), "\nspeedup_vs_torch_gpu", d.get("speedup_vs_torch_gpu"))
    print("other kernels", json.dumps(d["roofline"].get("other_kernels"), indent=0)[:1500])
except Exception as e:
    print("bench unreadable:", e)
PY
