"""Generate tests/golden/padding_modes.pt from the UNMODIFIED reference (build container only):

    python oracle/make_golden_padding.py

`Unet3D(padding_mode='circular' | 'circular_1d')` (model.yaml:13, VDDP:153-243, 270-273, 624-628) on a small configuration:
state_dict key order (the circular variants wrap some convolutions in a module, which renames their keys) and the forward
outputs with and without conditioning on seeded inputs.  The oracle restates both modes; the
product's wrap-mode path is held to these fixtures through the oracle (tests/test_cpu_forward_glue.py, tests/test_gpu_unet.py).
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

from denoising_diffusion_pytorch.video_denoising_diffusion_pytorch import Unet3D  # noqa: E402
from oracle import vdm_oracle as O  # noqa: E402


def main():
    out = {}
    for mode in ("circular", "circular_1d"):
        cfg = O.UnetCfg(dim=16, dim_mults=(1, 2), padding_mode=mode)
        model = Unet3D(dim=cfg.dim, dim_mults=cfg.dim_mults, channels=cfg.channels, attn_heads=cfg.heads, attn_dim_head=cfg.dim_head,
                       init_dim=None, init_kernel_size=cfg.init_kernel, use_sparse_linear_attn=True, resnet_groups=cfg.groups,
                       cond_bias=True, cond_attention="self-stacked", cond_attention_tokens=16, cond_att_GRU=False,
                       use_temporal_attention_cond=True, cond_to_time="add", per_frame_cond=True, padding_mode=mode)
        sd = O.synthetic_state_dict(cfg, seed=21)
        assert list(model.state_dict().keys()) == list(sd.keys()), "state_dict keys / order differ from the reference"
        model.load_state_dict(sd, strict=True)
        model.eval()
        g = torch.Generator().manual_seed(22)
        x = torch.randn(2, 3, 11, 16, 16, generator=g)
        cond = torch.rand(2, 11, generator=g) * 2 - 1
        t = torch.tensor([1, 6])
        with torch.no_grad():
            y_cond = model(x, t, cond=cond, null_cond_prob=0.0)
            y_null = model(x, t, cond=cond, null_cond_prob=1.0)
            o_cond = O.unet_forward(sd, cfg, x, t, cond, torch.zeros(2, dtype=torch.bool))
            o_null = O.unet_forward(sd, cfg, x, t, cond, torch.ones(2, dtype=torch.bool))
        rel = lambda a, b: float((a - b).norm() / b.norm())
        print(f"{mode}: oracle vs reference rel-L2 cond {rel(o_cond, y_cond):.2e} null {rel(o_null, y_null):.2e}; "
              f"differs from zero padding by {rel(O.unet_forward(_zeros_sd(sd, mode), O.UnetCfg(dim=16, dim_mults=(1, 2)), x, t, cond, torch.zeros(2, dtype=torch.bool)), y_cond):.2e}")
        out[mode] = dict(keys=list(sd.keys()), seed=21, data_seed=22, t=t, y_cond=y_cond, y_null=y_null)
    path = os.path.join(ROOT, "tests", "golden", "padding_modes.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path), "bytes")


def _zeros_sd(sd, mode):
    """The same weights under the zero-padding key names (to show that the padding mode changes the output)."""
    return {k.replace(".conv_transpose.", ".").replace(".conv.", "."): v for k, v in sd.items()}


if __name__ == "__main__":
    main()
