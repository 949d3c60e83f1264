"""Full-configuration goldens from the UNMODIFIED reference (build container only; ~1.5 h of CPU for the two samplers).

    python oracle/make_golden_full.py grads      # tests/golden/full_grads.pt   (~1 min)
    python oracle/make_golden_full.py ancestral  # tests/golden/full_sample_ancestral.pt (256 steps, ~45 min on 8 cores)
    python oracle/make_golden_full.py ddim       # tests/golden/full_sample_ddim.pt      (250 steps, ~45 min)

Shipped configuration (dim 64, 4 levels, 96x96x11, 256 timesteps, guidance w = 5, dynamic thresholding), batch 1,
synthetic weights `O.synthetic_state_dict(cfg, seed=0)`.  All inputs are regenerated from seeds by the tests:
the k-th noise draw of a sampler is `torch.randn(shape, generator=torch.Generator().manual_seed(NOISE_SEED + k))`
on the CPU (`noise_draw` below), so no noise tensor is stored.  Stored: the final sample (full fp32 tensor),
strided slices of intermediate states, and for the gradients every parameter's gradient norm plus strided slices.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
os.chdir(REF)

from denoising_diffusion_pytorch.video_denoising_diffusion_pytorch import GaussianDiffusion, Unet3D  # noqa: E402
from oracle import vdm_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
NOISE_SEED = 1000
SHAPE = (1, 3, 11, 96, 96)
TARGET_ROW0 = [2.81, 30.4, 49.7, 60.1, 67.4, 73.1, 77.4, 81.3, 84.3, 86.1, 85.1]    # data/target_responses.csv row 0
CHECKPOINTS = (255, 252, 248, 240, 224, 192, 128, 64, 16, 0)                          # states kept after the step at these t


_RANDN = torch.randn          # the unpatched generator (SeededNoise replaces torch.randn while a sampler runs)


def noise_draw(k: int, shape=SHAPE):
    return _RANDN(shape, generator=torch.Generator().manual_seed(NOISE_SEED + k))


def sample_cond():
    """Row 0 of the reference's target responses under the bench's synthetic global-min-max-2 scaling (min 0, max 100)."""
    return (torch.tensor([TARGET_ROW0]) / 100.0) * 2 - 1


def sl(x):
    return x[:, :, :, ::4, ::4].clone()


class SeededNoise:
    """torch.randn / torch.randn_like return noise_draw(0), noise_draw(1), ... in call order."""

    def __enter__(self):
        self.k = 0
        self._a, self._b = torch.randn, torch.randn_like

        def nxt(*a, **kw):
            t = noise_draw(self.k)
            self.k += 1
            return t

        torch.randn, torch.randn_like = nxt, nxt
        return self

    def __exit__(self, *e):
        torch.randn, torch.randn_like = self._a, self._b


def build(sampling_timesteps, loss_type="l1"):
    cfg = O.UnetCfg()
    model = Unet3D(dim=cfg.dim, dim_mults=cfg.dim_mults, channels=cfg.channels, attn_heads=cfg.heads, attn_dim_head=cfg.dim_head,
                   init_dim=None, init_kernel_size=cfg.init_kernel, use_sparse_linear_attn=True, resnet_groups=cfg.groups,
                   cond_bias=True, cond_attention="self-stacked", cond_attention_tokens=16, cond_att_GRU=False,
                   use_temporal_attention_cond=True, cond_to_time="add", per_frame_cond=True, padding_mode="zeros")
    gd = GaussianDiffusion(model, image_size=96, channels=3, num_frames=11, timesteps=256, loss_type=loss_type,
                           use_dynamic_thres=True, sampling_timesteps=sampling_timesteps)
    model.load_state_dict(O.synthetic_state_dict(cfg, seed=0), strict=True)
    model.eval()
    return model, gd


def grads():
    out = dict(seed=0, data_seed=1)
    g = torch.Generator().manual_seed(1)
    x01 = torch.rand(*SHAPE, generator=g)
    cond = torch.rand(1, 11, generator=g) * 2 - 1
    noise = torch.randn(*SHAPE, generator=g)
    out["t"] = 100
    t = torch.tensor([100])
    for lt in ("l1", "l2"):
        model, gd = build(256, lt)
        model.zero_grad()
        loss = gd.p_losses(x01 * 2 - 1, t, cond=cond, noise=noise, null_cond_prob=0.0)
        loss.backward()
        out[lt] = dict(
            loss=float(loss),
            grad_norms={k: float(p.grad.norm()) for k, p in model.named_parameters() if p.grad is not None},
            grad_none=[k for k, p in model.named_parameters() if p.grad is None and p.requires_grad],
            # every parameter: up to 256 evenly strided gradient entries
            grad_slices={k: p.grad.flatten()[::max(1, p.grad.numel() // 256)][:256].clone()
                         for k, p in model.named_parameters() if p.grad is not None})
        print(lt, "loss", float(loss), "params with grad", len(out[lt]["grad_norms"]), flush=True)
    torch.save(out, os.path.join(OUT, "full_grads.pt"))


def ancestral():
    model, gd = build(256)
    cond = sample_cond()
    states = {}
    t0 = time.time()
    with torch.no_grad(), SeededNoise():
        img = torch.randn(SHAPE)
        for i in reversed(range(256)):
            img = gd.p_sample(img, torch.full((1,), i, dtype=torch.long), cond=cond, guidance_scale=5.0)
            if i in CHECKPOINTS:
                states[i] = dict(slice=sl(img), norm=float(img.norm()))
            if i % 16 == 0:
                print("ancestral t", i, "elapsed", round(time.time() - t0), "s |x|", float(img.norm()), flush=True)
    torch.save(dict(kind="ancestral", noise_seed=NOISE_SEED, cond=cond, guidance_scale=5.0, states=states,
                    final=((img + 1) * 0.5).clone()), os.path.join(OUT, "full_sample_ancestral.pt"))


def ddim():
    model, gd = build(250)
    cond = sample_cond()
    assert gd.is_ddim_sampling
    # the reference's own loop (VDDP:986-1018), instrumented by wrapping the network call to record intermediate states
    states = {}
    calls = {"n": 0}
    fwd = model.forward_with_guidance_scale
    t0 = time.time()

    def wrapped(x, t, **kw):
        n = calls["n"]
        if n in (1, 4, 8, 16, 32, 64, 128, 192, 249):
            states[n] = dict(slice=sl(x), norm=float(x.norm()))     # state BEFORE model call n = after n updates
        if n % 16 == 0:
            print("ddim call", n, "elapsed", round(time.time() - t0), "s |x|", float(x.norm()), flush=True)
        calls["n"] = n + 1
        return fwd(x, t, **kw)

    model.forward_with_guidance_scale = wrapped
    with torch.no_grad(), SeededNoise():
        out = gd.sample(cond=cond, guidance_scale=5.0)
    torch.save(dict(kind="ddim250", noise_seed=NOISE_SEED, cond=cond, guidance_scale=5.0, states=states, final=out.clone()),
               os.path.join(OUT, "full_sample_ddim.pt"))


if __name__ == "__main__":
    torch.set_num_threads(int(os.environ.get("GOLD_THREADS", "6")))
    for what in sys.argv[1:]:
        dict(grads=grads, ancestral=ancestral, ddim=ddim)[what]()
