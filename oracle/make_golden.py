"""Generate tests/golden/* from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # writes tests/golden/*.pt, *.json

/root/reference is imported through the stand-in packages in oracle/shims (einops_exts,
rotary_embedding_torch, accelerate, imageio, matplotlib are not installed and there is no
network).  Nothing here is used at run time on the GPU box: the committed fixtures are.
Every fixture stores the reference's outputs for inputs that can be regenerated from seeds, so
the files stay small.  The script also prints oracle-vs-reference differences so that a drift in
oracle/vdm_oracle.py is visible at generation time.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)                                   # the reference's `denoising_diffusion_pytorch` must shadow this repo's re-export
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
os.chdir(REF)   # the reference imports `src.*` relative to its repo root

from denoising_diffusion_pytorch.video_denoising_diffusion_pytorch import (  # noqa: E402
    GaussianDiffusion, RelativePositionBias, SinusoidalPosEmb, Unet3D, cosine_beta_schedule)
from oracle import vdm_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def build_reference(cfg: O.UnetCfg, timesteps: int, sampling_timesteps: int, image_size: int):
    model = Unet3D(dim=cfg.dim, dim_mults=cfg.dim_mults, channels=cfg.channels, attn_heads=cfg.heads,
                   attn_dim_head=cfg.dim_head, init_dim=None, init_kernel_size=cfg.init_kernel,
                   use_sparse_linear_attn=True, resnet_groups=cfg.groups, cond_bias=True,
                   cond_attention="self-stacked", cond_attention_tokens=16, cond_att_GRU=False,
                   use_temporal_attention_cond=True, cond_to_time="add", per_frame_cond=True, padding_mode="zeros")
    gd = GaussianDiffusion(model, image_size=image_size, channels=cfg.channels, num_frames=cfg.frames,
                           timesteps=timesteps, loss_type="l1", use_dynamic_thres=True,
                           sampling_timesteps=sampling_timesteps)
    return model, gd


class ReplayNoise:
    """Make torch.randn / torch.randn_like return a recorded list, in call order."""

    def __init__(self, tensors):
        self.tensors = list(tensors)
        self.i = 0

    def __enter__(self):
        self._randn, self._randn_like = torch.randn, torch.randn_like

        def nxt(*a, **k):
            t = self.tensors[self.i]
            self.i += 1
            return t.clone()

        torch.randn, torch.randn_like = nxt, nxt
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self._randn, self._randn_like


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def known_answers():
    kat = {}
    b = cosine_beta_schedule(256)
    kat["betas64_first3"] = b[:3].tolist()
    kat["betas64_last3"] = b[-3:].tolist()
    _, gd = build_reference(O.UnetCfg(dim=16, dim_mults=(1, 2)), 256, 256, 16)
    kat["buffers_0_128_255"] = {k: [float(getattr(gd, k)[i]) for i in (0, 128, 255)] for k in O.SCHEDULE_KEYS}
    bk = RelativePositionBias._relative_position_bucket(
        torch.arange(11)[None, :] - torch.arange(11)[:, None], num_buckets=32, max_distance=32)
    kat["rel_pos_buckets_11"] = bk.tolist()
    kat["sinusoidal64_t128"] = SinusoidalPosEmb(64)(torch.tensor([128.0]))[0].tolist()
    times = list(reversed(torch.linspace(-1, 255, steps=251).int().tolist()))
    kat["ddim_pairs_250_of_256"] = list(zip(times[:-1], times[1:]))
    full = O.UnetCfg()
    m, _ = build_reference(full, 256, 256, 96)
    sd = m.state_dict()
    kat["full_state_dict_keys"] = list(sd.keys())
    kat["full_state_dict_shapes"] = [list(v.shape) for v in sd.values()]
    kat["full_param_count"] = sum(p.numel() for p in m.parameters())
    kat["quantile_index_304128_q0.9"] = 0.9 * (304128 - 1)
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(kat, f)
    # oracle cross-checks
    S = O.schedule(256)
    for k in O.SCHEDULE_KEYS:
        assert torch.equal(S[k], getattr(gd, k)), k
    assert torch.equal(O.rel_pos_buckets(11), bk)
    assert list(O.unet_param_shapes(full).keys()) == list(sd.keys())
    assert [tuple(v) for v in O.unet_param_shapes(full).values()] == [tuple(v.shape) for v in sd.values()]
    print("kat.json written; schedule / buckets / state-dict layout agree with the oracle")


def small_case():
    cfg = O.UnetCfg(dim=16, dim_mults=(1, 2))
    T, size, B = 8, 16, 2
    model, gd = build_reference(cfg, T, T, size)
    sd = O.synthetic_state_dict(cfg, seed=3)
    model.load_state_dict(sd, strict=True)
    model.eval()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 3, cfg.frames, size, size, generator=g)
    cond = torch.rand(B, cfg.frames, generator=g) * 2 - 1
    t = torch.tensor([1, 6])
    x01 = torch.rand(B, 3, cfg.frames, size, size, generator=g)
    noise = torch.randn(B, 3, cfg.frames, size, size, generator=g)
    step_noise = [torch.randn(B, 3, cfg.frames, size, size, generator=g) for _ in range(T + 1)]
    gold = dict(cfg=dict(dim=16, dim_mults=(1, 2)), T=T, size=size, seed=3, x=x, cond=cond, t=t, x01=x01,
                noise=noise, step_noise=torch.stack(step_noise))
    with torch.no_grad():
        gold["y_cond"] = model(x, t, cond=cond, null_cond_prob=0.0)
        gold["y_null"] = model(x, t, cond=cond, null_cond_prob=1.0)
        gold["y_guided"] = model.forward_with_guidance_scale(x, t, cond=cond, guidance_scale=5.0)
    # training loss + gradients; null_cond_prob=0 keeps the label mask deterministic
    model.zero_grad()
    loss = gd.p_losses(x01 * 2 - 1, t, cond=cond, noise=noise, null_cond_prob=0.0)
    loss.backward()
    gold["loss"] = loss.detach()
    gold["grad_norms"] = {k: float(p.grad.norm()) for k, p in model.named_parameters() if p.grad is not None}
    gold["grad_none"] = [k for k, p in model.named_parameters() if p.grad is None and p.requires_grad]
    gold["grad_samples"] = {k: model.get_parameter(k).grad.flatten()[:64].clone() for k in (
        "downs.0.0.block1.proj.weight", "downs.0.3.fn.fn.fn.to_qkv.weight", "downs.1.2.fn.fn.to_out.weight",
        "mid_spatial_attn.fn.fn.fn.to_qkv.weight", "ups.0.4.weight", "downs.0.4.weight", "init_conv.weight",
        "null_text_token", "time_mlp.1.weight", "sign_emb.weight", "final_conv.1.weight",
        "time_rel_pos_bias.relative_attention_bias.weight", "downs.0.2.fn.norm.gamma")}
    # one ancestral step (teacher forced) at t=[1,6] and at t=0, then the whole T-step loop, then DDIM 4-of-8
    with torch.no_grad():
        with ReplayNoise([step_noise[0]]):
            gold["p_sample_t"] = gd.p_sample(x, t, cond=cond, guidance_scale=5.0)
        with ReplayNoise([step_noise[0]]):
            gold["p_sample_t0"] = gd.p_sample(x, torch.zeros(B, dtype=torch.long), cond=cond, guidance_scale=5.0)
        with ReplayNoise(step_noise):
            gold["loop"] = gd.sample(cond=cond, guidance_scale=5.0)
        _, gd4 = build_reference(cfg, T, 4, size)
        gd4.denoise_fn.load_state_dict(sd)
        with ReplayNoise(step_noise):
            gold["ddim4"] = gd4.sample(cond=cond, guidance_scale=5.0)
    torch.save(gold, os.path.join(OUT, "small_unet.pt"))

    # oracle vs reference
    S = O.schedule(T)
    keep = torch.zeros(B, dtype=torch.bool)
    with torch.no_grad():
        print("small y_cond   rel", rel(O.unet_forward(sd, cfg, x, t, cond, keep), gold["y_cond"]))
        print("small y_null   rel", rel(O.unet_forward(sd, cfg, x, t, cond, ~keep), gold["y_null"]))
        print("small guided   rel", rel(O.unet_forward_guided(sd, cfg, x, t, cond, 5.0), gold["y_guided"]))
        print("small p_sample rel", rel(O.p_sample(sd, cfg, S, x, t, cond, 5.0, step_noise[0]), gold["p_sample_t"]))
        print("small loop     rel", rel(O.p_sample_loop(sd, cfg, S, step_noise[0], cond, 5.0, step_noise[1:]), gold["loop"]))
        print("small ddim4    rel", rel(O.ddim_sample(sd, cfg, S, step_noise[0], cond, 5.0, 4), gold["ddim4"]))
    P = {k: v.clone().requires_grad_(v.is_floating_point() and "freqs" not in k) for k, v in sd.items()}
    l2 = O.p_losses(P, cfg, S, x01, t, cond, noise, keep)
    l2.backward()
    print("small loss     abs", float((l2 - gold["loss"]).abs()))
    worst = max(abs(float(P[k].grad.norm()) - v) / max(v, 1e-12) for k, v in gold["grad_norms"].items())
    print("small grads    worst rel norm diff", worst, " never-used:", len(gold["grad_none"]))


def full_case():
    cfg = O.UnetCfg()
    model, gd = build_reference(cfg, 256, 256, 96)
    sd = O.synthetic_state_dict(cfg, seed=0)
    model.load_state_dict(sd, strict=True)
    model.eval()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 3, 11, 96, 96, generator=g)
    cond = torch.rand(1, 11, generator=g) * 2 - 1
    t = torch.tensor([128])
    noise = torch.randn(1, 3, 11, 96, 96, generator=g)
    with torch.no_grad():
        yc = model(x, t, cond=cond, null_cond_prob=0.0)
        yn = model(x, t, cond=cond, null_cond_prob=1.0)
        with ReplayNoise([noise]):
            ps = gd.p_sample(x, t, cond=cond, guidance_scale=5.0)
    sl = (slice(None), slice(None), slice(None, None, 2), slice(None, None, 8), slice(None, None, 8))
    gold = dict(seed=0, data_seed=1, t=128,
                y_cond_slice=yc[sl].clone(), y_null_slice=yn[sl].clone(), p_sample_slice=ps[sl].clone(),
                y_cond_norm=float(yc.norm()), y_null_norm=float(yn.norm()), p_sample_norm=float(ps.norm()),
                y_cond_mean=float(yc.mean()), y_cond_absmax=float(yc.abs().max()))
    torch.save(gold, os.path.join(OUT, "full_unet_slices.pt"))
    keep = torch.zeros(1, dtype=torch.bool)
    with torch.no_grad():
        oc = O.unet_forward(sd, cfg, x, t, cond, keep)
    print("full y_cond rel", rel(oc, yc), " |y|", float(yc.norm()), "absmax", float(yc.abs().max()))


if __name__ == "__main__":
    torch.manual_seed(0)
    known_answers()
    small_case()
    full_case()
