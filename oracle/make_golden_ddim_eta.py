"""Generate tests/golden/ddim_eta.pt from the UNMODIFIED reference (build container only):

    python oracle/make_golden_ddim_eta.py

`GaussianDiffusion(..., sampling_timesteps=4 < timesteps=8, ddim_sampling_eta=0.5)`: DDIM with a stochastic part (VDDP:986-1018),
guided (w = 5), on a small network with seeded weights and a recorded noise sequence.  Pins the oracle's `eta` branch."""
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


def main():
    cfg = O.UnetCfg(dim=16, dim_mults=(1, 2))
    model = Unet3D(dim=cfg.dim, dim_mults=cfg.dim_mults, channels=cfg.channels, attn_heads=cfg.heads, attn_dim_head=cfg.dim_head,
                   init_dim=None, init_kernel_size=cfg.init_kernel, use_sparse_linear_attn=True, resnet_groups=cfg.groups, cond_bias=True,
                   cond_attention="self-stacked", cond_attention_tokens=16, cond_att_GRU=False, use_temporal_attention_cond=True,
                   cond_to_time="add", per_frame_cond=True, padding_mode="zeros")
    sd = O.synthetic_state_dict(cfg, seed=31)
    model.load_state_dict(sd, strict=True)
    out = {}
    for eta in (0.5, 1.0):
        gd = GaussianDiffusion(model, image_size=12, channels=3, num_frames=11, timesteps=8, loss_type="l1", use_dynamic_thres=True,
                               sampling_timesteps=4, ddim_sampling_eta=eta)
        g = torch.Generator().manual_seed(32)
        cond = torch.rand(2, 11, generator=g) * 2 - 1
        noises = [torch.randn(2, 3, 11, 12, 12, generator=g) for _ in range(4)]          # x_T, then one draw per step with time_next >= 0
        it = iter(noises)
        randn, randn_like = torch.randn, torch.randn_like
        torch.randn = lambda *a, **k: next(it).clone()
        torch.randn_like = lambda *a, **k: next(it).clone()
        try:
            y = gd.sample(cond=cond, guidance_scale=5.0)
        finally:
            torch.randn, torch.randn_like = randn, randn_like
        o = O.ddim_sample(sd, cfg, O.schedule(8), noises[0], cond, 5.0, 4, eta=eta, noises=noises[1:])
        print(f"eta={eta}: oracle vs reference rel-L2 {float((o - y).norm() / y.norm()):.2e}")
        out[eta] = dict(seed=31, data_seed=32, sample=y)
    path = os.path.join(ROOT, "tests", "golden", "ddim_eta.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
