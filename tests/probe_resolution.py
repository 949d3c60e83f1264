"""Resolution sweep (BASELINE configs[4]): one eager + graph-replayed training step at 64 / 96 / 128 px, 11 frames, b=2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videometamaterials_b200 import Accelerator, GaussianDiffusion, Trainer, Unet3D
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for px, fwd_gflop in ((64, 172.10), (96, 387.26), (128, 688.73)):
    torch.manual_seed(0)
    model = Unet3D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
                   cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True)
    gd = GaussianDiffusion(model, image_size=px, channels=3, num_frames=11, timesteps=256, use_dynamic_thres=True, sampling_timesteps=256)
    tr = Trainer(gd, None, None, [0, 1, 3], train_batch_size=B, results_folder="gpurun_out/probe_run", log=False, null_cond_prob=0.1,
                 per_frame_cond=True, reference_frame='lagrangian', accelerator=Accelerator("bf16"))
    x = torch.rand(B, 3, 11, px, px, device="cuda"); c = torch.rand(B, 11, device="cuda") * 2 - 1
    losses = []
    for _ in range(25):
        tr.step += 1
        losses.append(float(tr.train_step(x, c)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tr.step += 1
        tr.train_step(x, c)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{px}x{px}x11 b={B}: {ms:.2f} ms/step, {B / ms * 1e3:.1f} clips/s, model {B * 3 * fwd_gflop / ms:.0f} TFLOP/s (3 x forward FLOPs), "
          f"losses {losses[0]:.4f} -> {losses[-1]:.4f} finite={all(l == l for l in losses)}", flush=True)
    del tr, gd, model
    torch.cuda.empty_cache()
