#!/usr/bin/env python
"""Benchmark of the VideoMetamaterials hot path (driver contract: see the repository task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W]         own arm: B200 kernels
  python bench.py --impl reference [...]                       reference arm: the UNMODIFIED reference (baseline/_ref) on the host CPU
  python bench.py --impl torch-gpu                             the UNMODIFIED reference's own torch path on the B200 (fp32 and bf16
                                                               autocast, b=8): the number the kernels have to beat; the own arm runs
                                                               it in a subprocess and reports it as `torch_gpu_baseline`
  python bench.py --global-batch 32 [--gpus N]                 BASELINE configs[3]: strong scaling, 32/N clips per GPU

Metric (BASELINE.json): UNet3D fwd+bwd video-clips/s at 96x96x11 (+ p_sample steps/s as an extra key).
Workload at every N: configs[1] = "Unet3D fwd+bwd bf16, batch=8, 96x96x11" per GPU (weak scaling, global batch 8N);
a step is one full optimisation step of the Trainer (forward, backward, NCCL gradient all-reduce when N > 1, fused
Adam + EMA + weight repack) on synthetic clips of the Dataset's value range.
  value : device-timed (CUDA events), inputs resident in HBM, max over ranks.
  e2e   : the same step through the public API (`Trainer.train_step`) with pinned HOST buffers copied to the device
          and the loss read back to the host inside the timed region.
  roofline : the implicit-GEMM tcgen05 kernel (vmm_cgemm), tensor bound: algorithmic FLOPs of every launch in two
          profiled steps / their summed CUDA-event durations, against MEASURED_PEAKS.json bf16_tflops_sustained.
  cpu_baseline : the oracle port (fp32 torch on the host cores), one clip forward+backward.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "UNet3D fwd+bwd video-clips/sec (96x96x11)"
UNIT = "clips/s"
FWD_BWD_GFLOP_PER_CLIP = 1159.87      # SURVEY.md section 8d (torch flop counter on the reference)
FWD_GFLOP_PER_CLIP = 387.26


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), tflops=d.get("bf16_tflops", 1590.0), tflops_sustained=d.get("bf16_tflops_sustained", 1400.0),
                    source="measured")
    return dict(hbm=6650.0, tflops=1590.0, tflops_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def oracle_step_fn():
    import torch
    from oracle import vdm_oracle as O
    cfg = O.UnetCfg()
    sd = O.synthetic_state_dict(cfg, seed=0)
    P = {k: v.clone().requires_grad_(v.is_floating_point() and "freqs" not in k) for k, v in sd.items()}
    S = O.schedule(256)
    g = torch.Generator().manual_seed(1)
    x01 = torch.rand(1, 3, 11, 96, 96, generator=g)
    cond = torch.rand(1, 11, generator=g) * 2 - 1
    noise = torch.randn(1, 3, 11, 96, 96, generator=g)
    t = torch.randint(0, 256, (1,), generator=g)
    mask = torch.zeros(1, dtype=torch.bool)

    def step():
        for p in P.values():
            p.grad = None
        loss = O.p_losses(P, cfg, S, x01, t, cond, noise, mask)
        loss.backward()
        return float(loss)

    return step


def reference_step_fn(batch: int = 1, device: str = "cpu", autocast_dtype=None):
    """One forward+backward of the UNMODIFIED reference (baseline/_ref, staged by baseline/stage_ref.py) through its own public
    API: GaussianDiffusion.forward(x, cond=..., null_cond_prob=0.1) -> loss; loss.backward().  Returns (step, kind)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import stage_ref
    cwd = os.getcwd()
    vddp = stage_ref.import_reference()
    os.chdir(cwd)
    torch.manual_seed(0)
    model = vddp.Unet3D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, attn_heads=8, attn_dim_head=32, init_dim=None, init_kernel_size=7,
                        use_sparse_linear_attn=True, resnet_groups=8, cond_bias=True, cond_attention='self-stacked',
                        cond_attention_tokens=16, cond_att_GRU=False, use_temporal_attention_cond=True, cond_to_time='add',
                        per_frame_cond=True, padding_mode='zeros')
    gd = vddp.GaussianDiffusion(model, image_size=96, channels=3, num_frames=11, timesteps=256, loss_type='l1', use_dynamic_thres=True,
                                sampling_timesteps=256).to(device)
    torch.manual_seed(1)
    x01 = torch.rand(batch, 3, 11, 96, 96, device=device)
    cond = torch.rand(batch, 11, device=device) * 2 - 1

    def step():
        for p in gd.parameters():
            p.grad = None
        if autocast_dtype is not None:
            with torch.autocast(device_type="cuda", dtype=autocast_dtype):
                loss = gd(x01, cond=cond, null_cond_prob=0.1)
        else:
            loss = gd(x01, cond=cond, null_cond_prob=0.1)
        loss.backward()
        return loss

    return step, gd


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind = "reference"
    try:
        step, _ = reference_step_fn(1)
        what = "the UNMODIFIED reference (baseline/_ref), GaussianDiffusion.forward + backward"
    except Exception as e:  # noqa: BLE001     baseline/_ref not staged: the oracle port is the same algorithm
        step, kind = oracle_step_fn(), "port"
        what = f"oracle port (baseline/_ref unavailable: {str(e)[:80]})"
    t0 = time.perf_counter()
    step()                                   # first warm-up also gives the per-step cost
    est = time.perf_counter() - t0
    warm = max(args.warmup - 1, 0)
    budget = 170.0
    steps = args.steps
    if est * (warm + steps) > budget:        # keep the whole run within a few minutes
        warm = 1 if est * 3 <= budget else 0
        steps = max(1, int(budget / est) - warm)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    val = 1.0 / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm + 1,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Unet3D fwd+bwd, 96x96x11 (BASELINE configs[1] shape), " + what + " on the host CPU",
                   "sample": "1 clip per step (the own arm runs 8 per step per GPU; a CPU step of 8 clips takes minutes)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": "1 clip forward+backward per step, fp32, all host threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_torch_gpu(args):
    """The reference's own PyTorch path (cuDNN / cuBLAS / ATen eager) on this GPU: SURVEY.md section 8d "the number to beat"."""
    import torch
    assert torch.cuda.is_available()
    out = {"impl": "torch-gpu", "what": "UNMODIFIED reference (baseline/_ref) in torch eager on cuda:0, GaussianDiffusion.forward + backward, "
                                        "b=8, 96x96x11, random-init weights, synthetic data", "unit": UNIT, "torch": torch.__version__}
    B = 8

    def timed(step, n_warm=2, n=3):
        for _ in range(n_warm):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for name, dt in (("fp32", None), ("bf16_autocast", torch.bfloat16), ("fp16_autocast", torch.float16)):
        try:
            step, gd = reference_step_fn(B, "cuda", dt)
            ms = timed(step)
            out[name] = {"ms_per_step": ms, "value": B / (ms * 1e-3), "tflops": B * FWD_BWD_GFLOP_PER_CLIP / ms,
                         "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                         "tf32": {"matmul": torch.backends.cuda.matmul.allow_tf32, "cudnn": torch.backends.cudnn.allow_tf32}}
            if name == "fp32":
                # the reference's sampling step (fp32, guidance w = 5: two sequential forwards + torch.quantile), b = 4
                cond = torch.rand(4, 11, device="cuda") * 2 - 1
                img = torch.randn(4, 3, 11, 96, 96, device="cuda")
                t = torch.full((4,), 128, device="cuda", dtype=torch.long)
                ms_s = timed(lambda: gd.p_sample(img, t, cond=cond, guidance_scale=5.0), 2, 4)
                out["p_sample_fp32"] = {"ms_per_step": ms_s, "value": 1e3 / ms_s, "unit": "p_sample steps/s", "batch": 4, "guidance_scale": 5.0}
            del step, gd
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": str(e)[:300]}
    print(json.dumps(out), flush=True)


def cpu_baseline_bounded():
    """One clip forward+backward of the unmodified reference (else the oracle port) on all host cores: one warm-up run, one
    timed run (~20-30 s of CPU work)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cwd, path = os.getcwd(), list(sys.path)
    try:
        step, _ = reference_step_fn(1)
        kind = "reference"
    except Exception:  # noqa: BLE001
        step, kind = oracle_step_fn(), "port"
    step()
    t0 = time.perf_counter()
    step()
    dt = time.perf_counter() - t0
    os.chdir(cwd)
    sys.path[:] = path
    return {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "1 clip (b=1) forward+backward, fp32, all host threads, second of two runs"}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons every 0.2 s during the timed region.  NVML in-process (nvidia_ml_py); forking
    `nvidia-smi` from a process that holds a CUDA context stalls the launching thread for tens of milliseconds."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _sample_nvml(self):
        n = self.nvml
        sm = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        bits = [getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8), getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)]
        self.rows.append([str(sm), str(self.sm_max)] + ["Active" if (r & b) else "Not Active" for b in bits])

    def _sample_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            self.rows.append([c.strip() for c in out.split(",")])

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(r[0])) for r in self.rows if r[0].replace('.', '').isdigit())
        mx = max(int(float(r[1])) for r in self.rows if r[1].replace('.', '').isdigit())
        reasons = sorted({self.NAMES[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------
def run_own(args):
    import torch
    import torch.distributed as dist
    from videometamaterials_b200 import Accelerator, GaussianDiffusion, Trainer, Unet3D, _lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    acc = Accelerator(mixed_precision="bf16")
    dev = acc.device
    strong = args.global_batch is not None
    if strong and args.global_batch % world != 0:
        raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
    B = args.global_batch // world if strong else 8
    torch.manual_seed(0)
    model = Unet3D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, attn_heads=8, attn_dim_head=32, init_dim=None, init_kernel_size=7,
                   use_sparse_linear_attn=True, resnet_groups=8, cond_bias=True, cond_attention='self-stacked', cond_attention_tokens=16,
                   cond_att_GRU=False, use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True, padding_mode='zeros')
    gd = GaussianDiffusion(model, image_size=96, channels=3, num_frames=11, timesteps=256, loss_type='l1', use_dynamic_thres=True,
                           sampling_timesteps=256)
    trainer = Trainer(gd, folder=None, validation_folder=None, selected_channels=[0, 1, 3], train_batch_size=B, test_batch_size=4,
                      train_lr=1e-4, train_num_steps=10 ** 9, results_folder=os.path.join(ROOT, "gpurun_out", "bench_run"), log=False,
                      null_cond_prob=0.1, per_frame_cond=True, reference_frame='lagrangian', accelerator=acc)
    torch.manual_seed(1 + rank)
    x_host = torch.rand(B, 3, 11, 96, 96).pin_memory()
    c_host = (torch.rand(B, 11) * 2 - 1).pin_memory()
    x_dev, c_dev = x_host.to(dev), c_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        trainer.step += 1
        return trainer.train_step(x_dev, c_dev)

    # e2e: the inputs of step i + 1 leave pinned host memory on a copy stream while step i runs (what the Trainer's prepared loader does,
    # accel._DeviceLoader); every timed step issues one such copy and reads its loss back
    copy_stream = torch.cuda.Stream()
    pending = []

    def issue_copy():
        with torch.cuda.stream(copy_stream):
            xd = x_host.to(dev, non_blocking=True)
            cd = c_host.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        pending.append((xd, cd, ev))

    def step_e2e():
        trainer.step += 1
        if not pending:
            issue_copy()
        xd, cd, ev = pending.pop(0)
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        xd.record_stream(cur)
        cd.record_stream(cur)
        issue_copy()                          # the next step's inputs, under this step's kernels
        return float(trainer.train_step(xd, cd).item())

    for _ in range(3):                       # priming: two eager steps, then the CUDA-graph capture of forward + backward
        step_resident()
    for _ in range(8):                       # the first replays of a fresh graph are slow to enqueue; settle them one by one
        step_resident()
        torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):     # the W untimed warm-up steps proper (graph replays)
        step_resident()
    barrier()
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    ms = e0.elapsed_time(e1) / args.steps
    # e2e leg: host buffers in, loss out
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    # roofline of the dominant kernel (vmm_cgemm): profile two more steps with per-launch events.  Every rank runs
    # them (the steps contain the gradient all-reduce); only rank 0 reports.
    ops.PROFILE, ops.PROFILE_TAGS = [], True
    for _ in range(2):
        step_resident()                      # Trainer falls back to eager launches while ops.PROFILE is set
    barrier()
    prof, ops.PROFILE, ops.PROFILE_TAGS = ops.PROFILE, None, False

    # BASELINE configs[3] next to the weak-scaling line: global batch 32 split over the ranks (strong scaling), same step
    strong_leg = None
    if not strong and not args.no_strong and 32 % world == 0:
        try:
            import gc
            trainer.__dict__.pop("_graph_states", None)      # release the CUDA graph (and its private activation pool) of the b=8 shape
            trainer.__dict__.pop("_graph_state", None)
            gc.collect()
            torch.cuda.empty_cache()
            B2 = 32 // world
            torch.manual_seed(100 + rank)
            x2 = torch.rand(B2, 3, 11, 96, 96, device=dev)
            c2 = torch.rand(B2, 11, device=dev) * 2 - 1

            def step2():
                trainer.step += 1
                return trainer.train_step(x2, c2)
            for _ in range(3):
                step2()                       # two eager steps, then the capture
                torch.cuda.synchronize()
            for _ in range(4):
                step2()
                torch.cuda.synchronize()
            barrier()
            n2 = max(3, min(args.steps, 10))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n2):
                step2()
            e1.record()
            barrier()
            ms2 = e0.elapsed_time(e1) / n2
            if world > 1:
                t = torch.tensor([ms2], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms2 = float(t[0])
            strong_leg = {"workload": "BASELINE configs[3]: full Trainer step, global batch 32 (strong scaling)", "global_batch": 32,
                          "per_gpu_batch": B2, "n_gpus": world, "steps": n2, "ms_per_step": ms2, "value": 32 / (ms2 * 1e-3), "unit": UNIT,
                          "scaling": "strong", "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
            del x2, c2
            trainer.__dict__.pop("_graph_states", None)
            trainer.__dict__.pop("_graph_state", None)
            gc.collect()
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            strong_leg = {"error": str(e)[:300]}
            if world > 1:
                raise                        # a rank that fails here would leave the others waiting in a collective
    if rank != 0:
        return
    pk = peaks()
    by = {}
    for name, flops, a, b, nbytes in prof:
        kern, _, shape = name.partition("|")
        cls = kern
        if kern == "cgemm":
            cls = "cgemm/conv3x3" if re.search(r"taps=(9|18)x1", shape) else "cgemm/other"
        for key in {kern, cls, name}:
            d = by.setdefault(key, [0.0, 0.0, 0, 0.0])
            d[0] += flops
            d[1] += a.elapsed_time(b) * 1e-3
            d[2] += 1
            d[3] += nbytes
    cg = by.get("cgemm", [0.0, 1.0, 1, 0.0])
    achieved = cg[0] / cg[1] / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_gemm_dram.json")       # ncu dram__bytes_{read,write}.sum of every cgemm launch of one step
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("cgemm", {}).get("dram_bytes_per_launch")

    def sub(key):
        d = by.get(key)
        if not d:
            return None
        t = d[0] / d[1] / 1e12
        return {"achieved": t, "frac": t / pk["tflops_sustained"], "launches_per_step": d[2] // 2, "share_of_step": (d[1] / 2) / (ms * 1e-3)}
    top = sorted(((k, v) for k, v in by.items() if "|" in k and k.startswith("cgemm")), key=lambda kv: -kv[1][1])[:6]
    def mem_kernel(keys):
        """HBM-bound kernels: algorithmic bytes / summed CUDA-event durations against the measured copy bandwidth."""
        tot = [0.0, 0.0, 0, 0.0]
        for k in keys:
            d = by.get(k)
            if d:
                tot = [tot[0] + d[0], tot[1] + d[1], tot[2] + d[2], tot[3] + d[3]]
        if tot[1] <= 0:
            return None
        gbs = tot[3] / tot[1] / 1e9
        return {"bound": "hbm", "achieved": gbs, "unit": "GB/s", "peak": pk["hbm"], "frac": gbs / pk["hbm"], "tflops": tot[0] / tot[1] / 1e12,
                "launches_per_step": tot[2] // 2, "share_of_step": (tot[1] / 2) / (ms * 1e-3)}
    others = {"tattn_fwd": mem_kernel(["tattn_fwd"]), "tattn_bwd": mem_kernel(["tattn_bwd"]), "tattn_fused_fwd": mem_kernel(["ftattn_fwd"]),
              "lattn_fwd": mem_kernel(["lattn_fwd"]), "lattn_bwd": mem_kernel(["lattn_bwd"]),
              "sattn": mem_kernel(["sattn_fwd", "sattn_bwd"]), "groupnorm": mem_kernel(["gn_silu_fwd", "gn_silu_bwd"]),
              "layernorm": mem_kernel(["ln_fwd", "ln_bwd"]),
              # fused to_qkv data + weight gradient of the 64-channel levels (csrc/qkvbwd.cu): tcgen05, bound by the one read of d(qkv)
              "qkv_bwd_fused": mem_kernel(["qkv_bwd"])}
    roof = {"bound": "tensor", "kernel": "vmm::cgemm_kernel (all launches of a step)", "achieved": achieved, "peak": pk["tflops_sustained"],
            "unit": "TFLOP/s", "frac": achieved / pk["tflops_sustained"], "traffic": traffic,
            "traffic_note": "average DRAM bytes per launch, ncu (profiles/r1_gemm_dram.json); algorithmic_bytes_per_launch = operands read once + output written once",
            "algorithmic_bytes_per_launch": cg[3] / max(1, cg[2]), "algorithmic_flops_per_launch": cg[0] / max(1, cg[2]),
            "avg_launch_us": cg[1] / max(1, cg[2]) * 1e6,
            "peak_source": pk["source"] + " (bf16_tflops_sustained)",
            "launches_per_step": cg[2] // 2, "share_of_step": (cg[1] / 2) / (ms * 1e-3),
            "conv3x3": sub("cgemm/conv3x3"), "other_gemm": sub("cgemm/other"),
            "top_shapes": [{"shape": k.partition("|")[2], "tflops": v[0] / v[1] / 1e12, "ms_per_step": v[1] / 2 * 1e3} for k, v in top],
            "wgrad": {"achieved": by["wgrad"][0] / by["wgrad"][1] / 1e12 if "wgrad" in by else None,
                      "frac": by["wgrad"][0] / by["wgrad"][1] / 1e12 / pk["tflops_sustained"] if "wgrad" in by else None,
                      "share_of_step": (by["wgrad"][1] / 2) / (ms * 1e-3) if "wgrad" in by else None},
            "wgrad_top_shapes": [{"shape": k.partition("|")[2], "tflops": v[0] / v[1] / 1e12, "gbs": v[3] / v[1] / 1e9, "ms_per_step": v[1] / 2 * 1e3,
                                  "launches_per_step": v[2] // 2}
                                 for k, v in sorted(((k, v) for k, v in by.items() if "|" in k and k.startswith("wgrad")), key=lambda kv: -kv[1][1])[:10]],
            "other_kernels": {k: v for k, v in others.items() if v is not None},
            "note": "shares are eager per-launch CUDA-event times over the graph-replayed step time; they need not sum to 1"}
    # sampling metric (second half of BASELINE.json's metric): ancestral p_sample with guidance, b=4, replayed from a CUDA graph,
    # timed on the device; plus BASELINE configs[2], one full 250-step DDIM sample() of the 4 conditionings
    ps = None
    try:
        ema = trainer.ema_model
        ema.denoise_fn.set_compute_dtype(torch.float16)
        ema.use_cuda_graph = True
        condp = torch.rand(4, 11, device=dev) * 2 - 1

        def timed(fn):
            fn()                                               # warm-up (captures the step graph)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e-3

        T = ema.num_timesteps
        ema.num_timesteps = 8
        dt = timed(lambda: ema.sample(cond=condp, guidance_scale=5.0)) / 8
        ema.num_timesteps = T
        ps = {"value": 1.0 / dt, "unit": "p_sample steps/s", "batch": 4, "guidance_scale": 5.0, "dtype": "f16", "steps_timed": 8,
              "tflops": 4 * 2 * FWD_GFLOP_PER_CLIP / dt / 1e3}
        if not args.no_ddim:
            st, ddim = ema.sampling_timesteps, ema.is_ddim_sampling
            ema.sampling_timesteps, ema.is_ddim_sampling = 250, True
            try:
                total = timed(lambda: ema.sample(cond=condp, guidance_scale=5.0))
                ps["ddim250_sample"] = {"seconds": total, "steps_per_s": 250 / total, "samples_per_s": 4 / total, "batch": 4,
                                        "workload": "GaussianDiffusion.sample 250-step DDIM, 4 conditionings, w=5 (BASELINE configs[2])"}
            finally:
                ema.sampling_timesteps, ema.is_ddim_sampling = st, ddim
    except Exception as e:  # noqa: BLE001
        ps = dict(ps or {}, error=str(e)[:200])
    cpu = cpu_baseline_bounded() if world == 1 and not args.no_cpu_baseline else None
    # the reference's own torch path on this GPU (subprocess: its package name collides with the drop-in re-export of this repo)
    tgb = None
    if world == 1 and not args.no_torch_gpu and os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "denoising_diffusion_pytorch")):
        try:
            import gc
            ema._graphs.clear()               # the sampler's step graphs (and their pools) are no longer needed
            gc.collect()
            torch.cuda.empty_cache()          # cached blocks go back to the driver: the child process needs ~60 GB for b=8 in fp32
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "torch-gpu"], capture_output=True, text=True, timeout=900)
            rows = [l for l in r.stdout.splitlines() if l.startswith("{")]
            tgb = json.loads(rows[-1]) if rows else {"error": (r.stderr or "no output")[-300:]}
        except Exception as e:  # noqa: BLE001
            tgb = {"error": str(e)[:300]}
    value = world * B / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": (f"Unet3D fwd+bwd bf16, global batch {world * B} = {B} per GPU, 96x96x11 (BASELINE configs[3]); " if strong else
                                "Unet3D fwd+bwd bf16, batch=8 per GPU, 96x96x11 (BASELINE configs[1]); ") +
                               "step = forward + backward + gradient all-reduce + fused Adam/EMA + weight repack", "global_batch": world * B,
                   "l2": "activation working set per step (>20 GB) is far larger than the 126 MB L2; no explicit flush",
                   "e2e_input_pipeline": "pinned host buffers -> device on a copy stream, one step ahead (the copy of step i+1 runs under step i); loss read back every step",
                   "launch": "forward + backward replayed from one CUDA graph per step; all-reduce, Adam/EMA and repack launched eagerly",
                   "parallelism": f"dp{world}", "model_tflops_per_gpu": B * FWD_BWD_GFLOP_PER_CLIP / ms},
        "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4 + c_host.numel() * 4,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
        "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof, "p_sample": ps,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if strong_leg is not None:
        line["config4_strong_scaling"] = strong_leg
    if tgb is not None:
        line["torch_gpu_baseline"] = tgb
        try:
            line["speedup_vs_torch_gpu"] = {k: value / tgb[k]["value"] for k in ("fp32", "bf16_autocast", "fp16_autocast") if "value" in tgb.get(k, {})}
            if ps and "value" in ps and "p_sample_fp32" in tgb and "value" in tgb["p_sample_fp32"]:
                line["speedup_vs_torch_gpu"]["p_sample_fp32"] = ps["value"] / tgb["p_sample_fp32"]["value"]
        except Exception:  # noqa: BLE001
            pass
    print(json.dumps(line), flush=True)


def run_sweep(args):
    """BASELINE configs[4]: resolution sweep 64 / 96 / 128 px (x 11 frames; 22-frame rows when --sweep-frames lists them: no
    reference value exists there, the reference raises at 22 frames under per_frame_cond).  Per row: the
    full training step (graph replayed), model TFLOP/s against the sustained tensor peak, and the per-kernel roofline fractions
    of two eager steps timed with CUDA events per launch."""
    import gc
    import torch
    from videometamaterials_b200 import Accelerator, GaussianDiffusion, Trainer, Unet3D, ops
    pk = peaks()
    fwd_gflop = {64: 172.10, 96: 387.26, 128: 688.73}            # SURVEY.md section 8d (11 frames); scales linearly with frames
    rows = []
    B = args.sweep_batch
    for frames in [int(f) for f in args.sweep_frames.split(",")]:
        for px in (64, 96, 128):
            row = {"px": px, "frames": frames, "batch": B, "reference_parity": frames == 11}
            try:
                torch.manual_seed(0)
                kw = {} if frames == 11 else {"num_frames": frames}
                model = Unet3D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True,
                               resnet_groups=8, cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True,
                               cond_to_time='add', per_frame_cond=True, **kw)
                gd = GaussianDiffusion(model, image_size=px, channels=3, num_frames=frames, timesteps=256, use_dynamic_thres=True,
                                       sampling_timesteps=256)
                tr = Trainer(gd, None, None, [0, 1, 3], train_batch_size=B, results_folder=os.path.join(ROOT, "gpurun_out", "bench_run"),
                             log=False, null_cond_prob=0.1, per_frame_cond=True, reference_frame='lagrangian', accelerator=Accelerator("bf16"))
                x = torch.rand(B, 3, frames, px, px, device="cuda")
                c = torch.rand(B, frames, device="cuda") * 2 - 1
                for _ in range(12):
                    tr.step += 1
                    loss = tr.train_step(x, c)
                    torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    tr.step += 1
                    tr.train_step(x, c)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.steps
                gflop = fwd_gflop[px] * frames / 11.0 * (FWD_BWD_GFLOP_PER_CLIP / FWD_GFLOP_PER_CLIP)
                tf = B * gflop / ms
                row.update(ms_per_step=ms, clips_per_s=B / ms * 1e3, model_tflops=tf, frac_of_sustained_tensor_peak=tf / pk["tflops_sustained"],
                           loss_finite=bool(torch.isfinite(loss)))
                ops.PROFILE, ops.PROFILE_TAGS = [], False
                for _ in range(2):
                    tr.step += 1
                    tr.train_step(x, c)
                torch.cuda.synchronize()
                prof, ops.PROFILE = ops.PROFILE, None
                by = {}
                for name, flops, a, b, nbytes in prof:
                    d = by.setdefault(name, [0.0, 0.0, 0.0])
                    d[0] += flops
                    d[1] += a.elapsed_time(b) * 1e-3
                    d[2] += nbytes
                kern = {}
                for k, (fl, sec, nb) in by.items():
                    if k in ("cgemm", "wgrad"):
                        kern[k] = {"tflops": fl / sec / 1e12, "frac_tensor": fl / sec / 1e12 / pk["tflops_sustained"], "share": sec / 2 / (ms * 1e-3)}
                    else:
                        kern[k] = {"gbs": nb / sec / 1e9, "frac_hbm": nb / sec / 1e9 / pk["hbm"], "share": sec / 2 / (ms * 1e-3)}
                row["kernels"] = kern
                del tr, gd, model, x, c
            except Exception as e:  # noqa: BLE001
                row["error"] = str(e)[:200]
            gc.collect()
            torch.cuda.empty_cache()
            rows.append(row)
    print(json.dumps({"metric": "resolution sweep (BASELINE configs[4]): Trainer step, bf16", "unit": UNIT, "peak_tflops_sustained": pk["tflops_sustained"],
                      "peak_hbm_gbs": pk["hbm"], "rows": rows}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference", "torch-gpu"])
    ap.add_argument("--global-batch", type=int, default=None, help="strong scaling: this many clips per step over all ranks (BASELINE configs[3]: 32)")
    ap.add_argument("--no-strong", action="store_true", help="skip the extra global-batch-32 leg of the default run")
    ap.add_argument("--no-torch-gpu", action="store_true", help="skip the reference-torch-on-this-GPU baseline")
    ap.add_argument("--sweep", action="store_true", help="BASELINE configs[4]: resolution sweep 64 / 96 / 128 px with per-row roofline fractions")
    ap.add_argument("--sweep-frames", default="11", help="comma list of frame counts for --sweep (22: no reference parity)")
    ap.add_argument("--sweep-batch", type=int, default=8)
    ap.add_argument("--no-ddim", dest="no_ddim", action="store_true", help="skip the 250-step DDIM sample() timing")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.sweep:
        run_sweep(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch-gpu":
        run_torch_gpu(args)
    else:
        run_own(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
