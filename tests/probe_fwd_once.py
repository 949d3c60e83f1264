"""One warm-up + one measured forward (b=8, bf16) for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videometamaterials_b200 import Unet3D, blocks

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
model = Unet3D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
               cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True).cuda()
x = torch.randn(B, 3, 11, 96, 96, device="cuda"); t = torch.randint(0, 256, (B,), device="cuda"); cond = torch.rand(B, 11, device="cuda") * 2 - 1
mask = torch.zeros(B, dtype=torch.bool, device="cuda")
for _ in range(2):
    blocks.unet_forward(model, x, None, None, t, cond, mask)
torch.cuda.synchronize()
