#!/bin/bash
# round 2, GPU call D: device dataset (GIF decode + item kernels) tests in their own process, the dataset bench row, then the
# whole -m gpu suite (state check after the container re-creation) and a short own-arm bench.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/d_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dataset.py -x -q > gpurun_out/d_dataset_tests.log 2>&1; echo "dataset tests rc=$?" | tee gpurun_out/d_summary.txt
timeout 600 python tools/bench_device_dataset.py 256 > gpurun_out/d_dataset_bench.log 2>&1; echo "dataset bench rc=$?" | tee -a gpurun_out/d_summary.txt
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_dataset.py > gpurun_out/d_gpu_tests.log 2>&1; echo "gpu suite rc=$?" | tee -a gpurun_out/d_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-ddim --no-strong --no-torch-gpu --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?" | tee -a gpurun_out/d_summary.txt
tail -c 2500 gpurun_out/d_dataset_tests.log
tail -c 1500 gpurun_out/d_dataset_bench.log
tail -c 800 gpurun_out/d_gpu_tests.log
python - <<'PY'
import json
try:
    rows = [l for l in open("gpurun_out/d_bench.json") if l.startswith("{")]
    d = json.loads(rows[-1])
    print("ms/step", round(d["ms_per_step"], 2), "clips/s", round(d["value"], 1), "e2e", d["e2e"]["value"], "p_sample", (d.get("p_sample") or {}).get("value"))
except Exception as e:
    print("bench unreadable:", e)
PY
