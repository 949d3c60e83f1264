"""Two guided ancestral sampling steps (b = 4 conditionings, w = 5: a forward of 8, fp16, eager launches) for
`ncu --metrics gpu__time_duration.sum` launch lists of the `p_sample` metric (BASELINE configs[2])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videometamaterials_b200 import GaussianDiffusion, Unet3D

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
model = Unet3D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
               cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True)
model.set_compute_dtype(torch.float16)
gd = GaussianDiffusion(model, image_size=96, channels=3, num_frames=11, timesteps=256, loss_type='l1', use_dynamic_thres=True,
                       sampling_timesteps=256).cuda()
x = torch.randn(B, 3, 11, 96, 96, device="cuda")
cond = torch.rand(B, 11, device="cuda") * 2 - 1
for step in (200, 199, 198):
    t = torch.full((B,), step, device="cuda", dtype=torch.long)
    x = gd._p_sample_core(x, t, cond, 5.0, torch.randn_like(x))
torch.cuda.synchronize()
