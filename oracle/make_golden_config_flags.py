"""Generate tests/golden/config_flags.pt from the UNMODIFIED reference (build container only):

    python oracle/make_golden_config_flags.py

Two entries of the config surface beside the shipped values (model.yaml:21-22): `use_temporal_attention_cond=False` (temporal attention
without the stacked conditioning keys, VDDP:792-795) and `cond_to_time='concat'` (time embedding | label embedding as the input of every
ResnetBlock MLP, VDDP:666, 789-790), alone and together, on a small configuration: forward with and without the label, the training
loss and every parameter gradient on seeded inputs.  The oracle restates both flags and is pinned to these fixtures
(tests/test_cpu_oracle.py); the product is held to the oracle (tests/test_cpu_forward_glue.py, tests/test_gpu_unet.py).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
os.chdir(REF)

from denoising_diffusion_pytorch.video_denoising_diffusion_pytorch import GaussianDiffusion, Unet3D  # noqa: E402
from oracle import vdm_oracle as O  # noqa: E402

CASES = {"no_temporal_cond": dict(temporal_cond=False, cond_to_time="add"),
         "concat": dict(temporal_cond=True, cond_to_time="concat"),
         "both": dict(temporal_cond=False, cond_to_time="concat")}


def main():
    out = {}
    for name, kw in CASES.items():
        cfg = O.UnetCfg(dim=16, dim_mults=(1, 2), **kw)
        model = Unet3D(dim=cfg.dim, dim_mults=cfg.dim_mults, channels=cfg.channels, attn_heads=cfg.heads, attn_dim_head=cfg.dim_head,
                       init_dim=None, init_kernel_size=cfg.init_kernel, use_sparse_linear_attn=True, resnet_groups=cfg.groups,
                       cond_bias=True, cond_attention="self-stacked", cond_attention_tokens=16, cond_att_GRU=False,
                       use_temporal_attention_cond=cfg.temporal_cond, cond_to_time=cfg.cond_to_time, per_frame_cond=True, padding_mode="zeros")
        sd = O.synthetic_state_dict(cfg, seed=31)
        assert list(model.state_dict().keys()) == list(sd.keys()), "state_dict keys / order differ from the reference"
        model.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(32)
        x = torch.randn(2, 3, 11, 16, 16, generator=g)
        cond = torch.rand(2, 11, generator=g) * 2 - 1
        t = torch.tensor([1, 6])
        with torch.no_grad():
            y_cond = model(x, t, cond=cond, null_cond_prob=0.0)
            y_null = model(x, t, cond=cond, null_cond_prob=1.0)
            o_cond = O.unet_forward(sd, cfg, x, t, cond, torch.zeros(2, dtype=torch.bool))
            o_null = O.unet_forward(sd, cfg, x, t, cond, torch.ones(2, dtype=torch.bool))
        # training loss + gradients through the reference's own p_losses (VDDP:1044-1060), l1, no label drop
        gd = GaussianDiffusion(model, image_size=16, num_frames=11, channels=3, timesteps=8, loss_type="l1", sampling_timesteps=8)
        x01 = torch.rand(2, 3, 11, 16, 16, generator=g)
        noise = torch.randn(2, 3, 11, 16, 16, generator=g)
        loss = gd.p_losses(x01 * 2 - 1, t, cond=cond, noise=noise)
        loss.backward()
        grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()}
        rel = lambda a, b: float((a - b).norm() / b.norm())
        S = O.schedule(8)
        P = {k: v.clone().requires_grad_(v.is_floating_point() and "freqs" not in k) for k, v in sd.items()}
        lo = O.p_losses(P, cfg, S, x01, t, cond, noise, torch.zeros(2, dtype=torch.bool))
        lo.backward()
        worst = max(rel(P[k].grad, gv) for k, gv in grads.items() if gv is not None and float(gv.norm()) > 0)
        unused = sorted(k for k, gv in grads.items() if gv is None)
        print(f"{name}: oracle vs reference rel-L2 cond {rel(o_cond, y_cond):.2e} null {rel(o_null, y_null):.2e} loss {abs(float(lo) - float(loss)):.2e} "
              f"worst gradient {worst:.2e}; {len(unused)} parameters without a gradient")
        out[name] = dict(cfg=kw, seed=31, data_seed=32, t=t, y_cond=y_cond, y_null=y_null, loss=loss.detach(),
                         grad_norms={k: (float(v.norm()) if v is not None else None) for k, v in grads.items()},
                         grad_slices={k: v.reshape(-1)[:: max(v.numel() // 64, 1)][:64].clone() for k, v in grads.items() if v is not None})
    path = os.path.join(ROOT, "tests", "golden", "config_flags.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
