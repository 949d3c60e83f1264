"""Quick timing probe (GPU box): forward at b=8 and one p_sample step at b=4, eager and CUDA-graph."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videometamaterials_b200 import GaussianDiffusion, Unet3D, _lib, blocks

dtype = torch.bfloat16 if len(sys.argv) < 2 else getattr(torch, sys.argv[1])
torch.manual_seed(0)
model = Unet3D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
               cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True).cuda()
model.set_compute_dtype(dtype)
gd = GaussianDiffusion(model, image_size=96, channels=3, num_frames=11, timesteps=256, use_dynamic_thres=True, sampling_timesteps=256).cuda()
for B in (8,):
    x = torch.randn(B, 3, 11, 96, 96, device="cuda"); t = torch.randint(0, 256, (B,), device="cuda"); cond = torch.rand(B, 11, device="cuda") * 2 - 1
    mask = torch.zeros(B, dtype=torch.bool, device="cuda")
    for _ in range(2):
        blocks.unet_forward(model, x, None, None, t, cond, mask)
    torch.cuda.synchronize()
    n0 = _lib.launch_count(); t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        blocks.unet_forward(model, x, None, None, t, cond, mask)
    e1.record(); torch.cuda.synchronize()
    print(f"forward eager b={B} {dtype}: {e0.elapsed_time(e1)/5:.2f} ms/iter (host {1e3*(time.time()-t0)/5:.2f} ms), launches/iter {(_lib.launch_count()-n0)/5:.0f}")
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        blocks.unet_forward(model, x, None, None, t, cond, mask)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        y = blocks.unet_forward(model, x, None, None, t, cond, mask)
    g.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"forward graph b={B}: {ms:.2f} ms -> {B/ms*1e3:.1f} clip-forwards/s, {B*387.26/ms:.1f} TFLOP/s")
# p_sample b=4 with guidance (2b forward)
gd.use_cuda_graph = True
cond = torch.rand(4, 11, device="cuda") * 2 - 1
gd.num_timesteps_backup = gd.num_timesteps
gd.num_timesteps = 8
out = gd.sample(cond=cond, guidance_scale=5.0)
torch.cuda.synchronize(); t0 = time.time()
out = gd.sample(cond=cond, guidance_scale=5.0)
torch.cuda.synchronize(); dt_ = time.time() - t0
print(f"p_sample graph b=4 w=5: {dt_/8*1e3:.2f} ms/step -> {8/dt_:.1f} steps/s ; finite={bool(torch.isfinite(out).all())}")
