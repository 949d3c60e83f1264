#!/bin/bash
# round 2, GPU call A: fused temporal-attention kernel tests first (own process: a trap must not take the suite down), then the
# whole -m gpu suite, then bench unfused vs fused with the new arms.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -k "fused" -x -q -s > gpurun_out/a_fused_tests.log 2>&1; echo "fused tests rc=$?" | tee -a gpurun_out/a_summary.txt
timeout 1500 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_kernels.py::test_fused_temporal_attention_block > gpurun_out/a_gpu_tests.log 2>&1; echo "gpu suite rc=$?" | tee -a gpurun_out/a_summary.txt
VMM_NO_FUSED_TATTN=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-ddim --no-strong --no-torch-gpu --no-cpu-baseline > gpurun_out/a_bench_unfused.json 2> gpurun_out/a_bench_unfused.err; echo "bench unfused rc=$?" | tee -a gpurun_out/a_summary.txt
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench_fused.json 2> gpurun_out/a_bench_fused.err; echo "bench fused rc=$?" | tee -a gpurun_out/a_summary.txt
tail -c 1500 gpurun_out/a_fused_tests.log
tail -c 1200 gpurun_out/a_gpu_tests.log
python - <<'PY'
import json
for f in ("a_bench_unfused.json", "a_bench_fused.json"):
    try:
        rows = [l for l in open("gpurun_out/" + f) if l.startswith("{")]
        d = json.loads(rows[-1])
        print(f, "ms/step", round(d["ms_per_step"], 2), "clips/s", round(d["value"], 1), "p_sample", (d.get("p_sample") or {}).get("value"),
              "strong", (d.get("config4_strong_scaling") or {}).get("ms_per_step"), "torch_gpu", {k: (v.get("value") if isinstance(v, dict) else v) for k, v in (d.get("torch_gpu_baseline") or {}).items() if k in ("fp32", "bf16_autocast", "fp16_autocast", "p_sample_fp32", "error")})
    except Exception as e:
        print(f, "unreadable:", e)
PY
