"""Per-shape timing table of the cgemm / wgrad launches of one training step (b=8, bf16), CUDA events around each launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videometamaterials_b200 import Accelerator, GaussianDiffusion, Trainer, Unet3D, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
model = Unet3D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
               cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True)
gd = GaussianDiffusion(model, image_size=96, channels=3, num_frames=11, timesteps=256, use_dynamic_thres=True, sampling_timesteps=256)
tr = Trainer(gd, None, None, [0, 1, 3], train_batch_size=B, results_folder="gpurun_out/probe_run", log=False, null_cond_prob=0.1,
             per_frame_cond=True, reference_frame='lagrangian', accelerator=Accelerator("bf16"))
x = torch.rand(B, 3, 11, 96, 96, device="cuda"); c = torch.rand(B, 11, device="cuda") * 2 - 1
for _ in range(3):
    tr.step += 1
    tr.train_step(x, c)
torch.cuda.synchronize()
ops.PROFILE, ops.PROFILE_TAGS = [], True
for _ in range(2):
    tr.step += 1
    tr.train_step(x, c)
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
by = {}
for name, fl, e0, e1 in prof:
    a = by.setdefault(name, [0.0, 0.0, 0])
    a[0] += fl; a[1] += e0.elapsed_time(e1) * 1e-3; a[2] += 1
tot = sum(a[1] for a in by.values()) / 2
print(f"total GEMM time per step {tot * 1e3:.2f} ms")
for name, a in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1] / 2 * 1e6:9.0f} us/step  {a[2] // 2:3d}x  {a[1] / a[2] * 1e6:7.0f} us  {a[0] / a[1] / 1e12:7.1f} TF/s  {name}")
