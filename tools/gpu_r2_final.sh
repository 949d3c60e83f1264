#!/bin/bash
# round 2, final validation: whole -m gpu suite, smoke(), the driver's bench lines (own arm with every leg, reference arm), launch list
set -u
export PYTHONPATH=.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/final_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_gpu_tests.log 2>&1; echo "gpu suite rc=$?" | tee gpurun_out/final_summary.txt
tail -c 300 gpurun_out/final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/final_summary.txt
tail -2 gpurun_out/final_smoke.log
timeout 1500 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?" | tee -a gpurun_out/final_summary.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference arm rc=$?" | tee -a gpurun_out/final_summary.txt
python - <<'PY'
import json
for f in ("final_bench.json", "final_bench_reference.json"):
    try:
        d = json.loads([l for l in open("gpurun_out/" + f) if l.startswith("{")][-1])
        print(f, "ms/step", round(d.get("ms_per_step", 0), 2), "value", round(d["value"], 2), d["unit"], "e2e", d.get("e2e", {}).get("value"),
              "p_sample", (d.get("p_sample") or {}).get("value"), "ddim", ((d.get("p_sample") or {}).get("ddim250_sample") or {}).get("seconds"),
              "strong", (d.get("config4_strong_scaling") or {}).get("ms_per_step"),
              "torch_gpu", {k: (round(v.get("value"), 1) if isinstance(v, dict) and v.get("value") else v) for k, v in (d.get("torch_gpu_baseline") or {}).items() if k in ("fp32", "bf16_autocast", "fp16_autocast", "p_sample_fp32", "error")},
              "clocks", d.get("clocks"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
if [ "${1:-}" = "launches" ]; then
  for try in 1 2; do
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_train_launches.csv \
        python tests/probe_train_once.py 8 2 > gpurun_out/r2c_ncu_launches.log 2>&1 && break
  done
  python tools/summarize_launches.py gpurun_out/r2c_train_launches.csv 2 2>/dev/null | head -30
fi
