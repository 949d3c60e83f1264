"""Whole-network gradient accuracy (small model, fp16 + loss scale / bf16) against the fp32 CPU oracle, for both rotary
placements (VMM_NO_ROT_EPILOGUE toggles at import time, so each mode runs in its own process)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import GaussianDiffusion, Unet3D
    torch.set_num_threads(8)
    for seed in (1, 2, 3):
        for dt, scale in ((torch.float16, 4096.0), (torch.bfloat16, 1.0)):
            cfg = O.UnetCfg(dim=16, dim_mults=(1, 2))
            sd = O.synthetic_state_dict(cfg, seed=seed)
            g = torch.Generator().manual_seed(100 + seed)
            x01 = torch.rand(2, 3, 11, 16, 16, generator=g); cond = torch.rand(2, 11, generator=g) * 2 - 1
            noise = torch.randn(2, 3, 11, 16, 16, generator=g); t = torch.tensor([3, 6])
            P = {k: v.clone().requires_grad_(v.is_floating_point() and "freqs" not in k) for k, v in sd.items()}
            loss_ref = O.p_losses(P, cfg, O.schedule(8), x01, t, cond, noise, torch.zeros(2, dtype=torch.bool))
            loss_ref.backward()
            model = Unet3D(dim=16, dim_mults=(1, 2), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
                           cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True)
            model.load_state_dict(sd); model.set_compute_dtype(dt)
            gd = GaussianDiffusion(model, image_size=16, channels=3, num_frames=11, timesteps=8, loss_type='l1', use_dynamic_thres=True,
                                   sampling_timesteps=8).cuda()
            loss = gd.p_losses((x01 * 2 - 1).cuda(), t.cuda(), cond=cond.cuda(), noise=noise.cuda(), null_cond_prob=0.0)
            (loss * scale).backward()
            errs = []
            for k, p in model.named_parameters():
                r = P[k].grad
                if r is None or float(r.norm()) == 0 or p.grad is None:
                    continue
                errs.append(float((p.grad.cpu().float() / scale - r).norm() / r.norm()))
            errs.sort()
            print(f"seed {seed} {str(dt):15s} loss err {abs(float(loss) - float(loss_ref)) / float(loss_ref):.2e}  grad rel-L2: median {errs[len(errs)//2]:.4f} "
                  f"mean {sum(errs)/len(errs):.4f} max {errs[-1]:.4f}", flush=True)
else:
    for mode in ("epilogue", "in-kernel"):
        env = dict(os.environ)
        if mode == "in-kernel":
            env["VMM_NO_ROT_EPILOGUE"] = "1"
        print("== rotary", mode, flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__), "run"], env=env)
