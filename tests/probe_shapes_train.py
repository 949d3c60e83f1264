"""Per-shape table of the GEMM launches (cgemm / wgrad) of one b=8 training step: events around every launch."""
import os, sys, collections
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
    tr.step += 1; tr.train_step(x, c)
torch.cuda.synchronize()
ops.PROFILE = []; ops.PROFILE_TAGS = True
N = 3
for _ in range(N):
    tr.step += 1; tr.train_step(x, c)
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
agg = collections.OrderedDict()
for name, flops, a, b, _ in prof:
    d = agg.setdefault(name, [0, 0.0, 0.0])
    d[0] += 1; d[1] += flops; d[2] += a.elapsed_time(b) * 1e-3
tot = sum(d[2] for d in agg.values()) / N
print(f"total GEMM time per step {tot*1e3:.2f} ms")
print("| launches/step | ms/step | TFLOP/s | shape |")
for name, d in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"| {d[0]//N} | {d[2]/N*1e3:.3f} | {d[1]/d[2]/1e12:.0f} | {name} |")
