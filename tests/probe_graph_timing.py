"""Why do back-to-back graph replays run slower than synchronised ones?  Variants of the bench loop."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videometamaterials_b200 import Accelerator, GaussianDiffusion, Trainer, Unet3D
B = 8
torch.manual_seed(0)
model = Unet3D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
               cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True)
gd = GaussianDiffusion(model, image_size=96, channels=3, num_frames=11, timesteps=256, use_dynamic_thres=True, sampling_timesteps=256)
tr = Trainer(gd, None, None, [0, 1, 3], train_batch_size=B, results_folder="gpurun_out/probe_run", log=False, null_cond_prob=0.1,
             per_frame_cond=True, reference_frame='lagrangian', accelerator=Accelerator("bf16"))
x = torch.rand(B, 3, 11, 96, 96, device="cuda"); c = torch.rand(B, 11, device="cuda") * 2 - 1
def step():
    tr.step += 1
    return tr.train_step(x, c)
for _ in range(5):
    step()
torch.cuda.synchronize()
def timeit(name, fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); th = time.perf_counter() - t0
    torch.cuda.synchronize(); tw = time.perf_counter() - t0
    print(f"{name:40s} event {e0.elapsed_time(e1)/n:8.2f} ms/step   host-enqueue {th/n*1e3:8.2f}  wall {tw/n*1e3:8.2f}", flush=True)
timeit("full step, no sync", step)
timeit("full step, sync each", lambda: (step(), torch.cuda.synchronize()))
timeit("full step, no sync (again)", step)
st = tr._graph_state
timeit("graph replay only, no sync", lambda: st["graph"].replay())
timeit("graph replay only, sync each", lambda: (st["graph"].replay(), torch.cuda.synchronize()))
def opt_only():
    tr.opt.step(ema_flat=None, ema_mode=0, ema_beta=0.995)
timeit("optimizer+repack only", opt_only)
timeit("replay + opt, no sync", lambda: (st["graph"].replay(), opt_only()))
timeit("full step, no sync, 30 steps", step, 30)
