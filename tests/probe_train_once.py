"""Two training steps (b=8, bf16) for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videometamaterials_b200 import Accelerator, GaussianDiffusion, Trainer, Unet3D

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
model = Unet3D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
               cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True)
gd = GaussianDiffusion(model, image_size=96, channels=3, num_frames=11, timesteps=256, use_dynamic_thres=True, sampling_timesteps=256)
tr = Trainer(gd, None, None, [0, 1, 3], train_batch_size=B, results_folder="gpurun_out/probe_run", log=False, null_cond_prob=0.1,
             per_frame_cond=True, reference_frame='lagrangian', accelerator=Accelerator("bf16"))
tr.use_cuda_graph = False          # eager launches: one ncu row per kernel
x = torch.rand(B, 3, 11, 96, 96, device="cuda"); c = torch.rand(B, 11, device="cuda") * 2 - 1
for _ in range(steps):
    tr.step += 1
    tr.train_step(x, c)
torch.cuda.synchronize()
