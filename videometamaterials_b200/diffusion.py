"""GaussianDiffusion on the B200 kernels, keeping the reference's constructor / method surface.

Reference: VDDP:841-1067 (`GaussianDiffusion`), `cosine_beta_schedule` VDDP:829-839.
The network call, the guidance lerp + x0 prediction, the dynamic-threshold quantile and the posterior /
DDIM updates are each one kernel launch of libvmm_sm100.so; a whole `p_sample` step can be captured in a
CUDA graph and replayed (sampling is launch-latency bound otherwise).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import blocks, blocks_bwd, ops


def extract(a, t, x_shape):
    """VDDP:824-827."""
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def cosine_beta_schedule(timesteps, s=0.008):
    """VDDP:829-839."""
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps, dtype=torch.float64)
    alphas_cumprod = torch.cos(((x / timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
    alphas_cumprod = alphas_cumprod / alphas_cumprod[0]
    betas = 1 - (alphas_cumprod[1:] / alphas_cumprod[:-1])
    return torch.clip(betas, 0, 0.9999)


def normalize_img(t):
    return t * 2 - 1


def unnormalize_img(t):
    return (t + 1) * 0.5


def quantile_rank(n: int, q: float):
    """(k, frac) exactly as torch.quantile computes them for a float32 input: rank = q * (n - 1) in fp32."""
    rank = np.float32(q) * np.float32(n - 1)
    k = int(np.floor(rank))
    return k, float(np.float32(rank - np.float32(k)))


class GaussianDiffusion(nn.Module):
    def __init__(self, denoise_fn, *, image_size, num_frames, channels=4, timesteps=1000, loss_type='l1', use_dynamic_thres=False,
                 dynamic_thres_percentile=0.9, sampling_timesteps=1000, ddim_sampling_eta=0.):
        super().__init__()
        self.channels = channels
        self.image_size = image_size
        self.num_frames = num_frames
        self.denoise_fn = denoise_fn

        betas = cosine_beta_schedule(timesteps)
        alphas = 1. - betas
        alphas_cumprod = torch.cumprod(alphas, axis=0)
        alphas_cumprod_prev = F.pad(alphas_cumprod[:-1], (1, 0), value=1.)
        timesteps, = betas.shape
        self.num_timesteps = int(timesteps)
        self.loss_type = loss_type
        reg = lambda name, val: self.register_buffer(name, val.to(torch.float32))
        reg('betas', betas)
        reg('alphas_cumprod', alphas_cumprod)
        reg('alphas_cumprod_prev', alphas_cumprod_prev)
        reg('sqrt_alphas_cumprod', torch.sqrt(alphas_cumprod))
        reg('sqrt_one_minus_alphas_cumprod', torch.sqrt(1. - alphas_cumprod))
        reg('log_one_minus_alphas_cumprod', torch.log(1. - alphas_cumprod))
        reg('sqrt_recip_alphas_cumprod', torch.sqrt(1. / alphas_cumprod))
        reg('sqrt_recipm1_alphas_cumprod', torch.sqrt(1. / alphas_cumprod - 1))
        posterior_variance = betas * (1. - alphas_cumprod_prev) / (1. - alphas_cumprod)
        reg('posterior_variance', posterior_variance)
        reg('posterior_log_variance_clipped', torch.log(posterior_variance.clamp(min=1e-20)))
        reg('posterior_mean_coef1', betas * torch.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod))
        reg('posterior_mean_coef2', (1. - alphas_cumprod_prev) * torch.sqrt(alphas) / (1. - alphas_cumprod))
        self.use_dynamic_thres = use_dynamic_thres
        self.dynamic_thres_percentile = dynamic_thres_percentile
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else timesteps
        assert self.sampling_timesteps <= timesteps
        self.is_ddim_sampling = self.sampling_timesteps < timesteps
        self.ddim_sampling_eta = float(ddim_sampling_eta)
        if loss_type not in ('l1', 'l2'):
            raise NotImplementedError()
        self.use_cuda_graph = False        # bench / Trainer turn this on; parity tests run eagerly
        self._graphs = {}

    # ------------------------------------------------------------------ pieces kept for API parity
    def q_mean_variance(self, x_start, t):
        mean = extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
        variance = extract(1. - self.alphas_cumprod, t, x_start.shape)
        log_variance = extract(self.log_one_minus_alphas_cumprod, t, x_start.shape)
        return mean, variance, log_variance

    def predict_start_from_noise(self, x_t, t, noise):
        return extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise

    def q_posterior(self, x_start, x_t, t):
        mean = extract(self.posterior_mean_coef1, t, x_t.shape) * x_start + extract(self.posterior_mean_coef2, t, x_t.shape) * x_t
        return mean, extract(self.posterior_variance, t, x_t.shape), extract(self.posterior_log_variance_clipped, t, x_t.shape)

    def q_sample(self, x_start, t, noise=None):
        noise = noise if noise is not None else torch.randn_like(x_start)
        return extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start + extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise

    # ------------------------------------------------------------------ sampling
    def _eps_channels_last(self, x, t, cond, guidance_scale):
        """Network output for the guided step: fp32 channels-last [(2)b, f, h, w, c]; cond rows first, null rows second."""
        b = x.shape[0]
        if guidance_scale == 1:
            mask = torch.zeros(b, dtype=torch.bool, device=x.device)
            return blocks.unet_forward(self.denoise_fn, x, None, None, t, cond, mask), False
        mask = torch.cat((torch.zeros(b, dtype=torch.bool, device=x.device), torch.ones(b, dtype=torch.bool, device=x.device)))
        # x is passed ONCE: blocks.unet_forward runs the label-free stem (init_conv + init_temporal_attn) on b samples and duplicates its
        # output for the 2b conditional | unconditional rows (VMM_SHARED_STEM=0: the stem on the duplicated input, as before)
        xin = x if blocks.SHARED_STEM else torch.cat((x, x))
        eps = blocks.unet_forward(self.denoise_fn, xin, None, None, torch.cat((t, t)), torch.cat((cond, cond)), mask)
        return eps, True

    def _p_sample_core(self, x, t, cond, guidance_scale, noise, clip_denoised=True):
        b, c, f, h, w = x.shape
        eps_cl, has_null = self._eps_channels_last(x, t, cond, guidance_scale)
        sr = self.sqrt_recip_alphas_cumprod[t].contiguous()
        srm1 = self.sqrt_recipm1_alphas_cumprod[t].contiguous()
        x0 = torch.empty_like(x)
        ops.cfg_x0(x, eps_cl, has_null, guidance_scale, sr, srm1, x0, None, b, c, f, h, w)
        per = c * f * h * w
        s = None
        if clip_denoised:
            s = torch.ones(b, dtype=torch.float32, device=x.device)
            if self.use_dynamic_thres:
                k, frac = quantile_rank(per, self.dynamic_thres_percentile)
                ops.abs_quantile(x0, b, per, k, frac, 1.0, s)
        c1 = self.posterior_mean_coef1[t].contiguous()
        c2 = self.posterior_mean_coef2[t].contiguous()
        sig = ((t != 0).float() * (0.5 * self.posterior_log_variance_clipped[t]).exp()).contiguous()
        out = torch.empty_like(x)
        ops.posterior_step(x0, x, noise, s, c1, c2, sig, out, b, per)
        return out

    @torch.inference_mode()
    def p_sample(self, x, t, cond=None, clip_denoised=True, guidance_scale=1.):
        """VDDP:956-963.  Draws its noise with torch.randn_like exactly where the reference does."""
        x = x.contiguous().float()
        noise = torch.randn_like(x)
        return self._p_sample_core(x, t, cond, guidance_scale, noise, clip_denoised)

    @torch.inference_mode()
    def p_sample_loop(self, shape, cond=None, guidance_scale=1.):
        """VDDP:965-975."""
        device = self.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device)
        if self.use_cuda_graph:
            return unnormalize_img(self._graph_loop(img, cond, guidance_scale))
        for i in reversed(range(0, self.num_timesteps)):
            img = self.p_sample(img, torch.full((b,), i, device=device, dtype=torch.long), cond=cond, guidance_scale=guidance_scale)
        return unnormalize_img(img)

    def _step_graph(self, kind, img, cond, guidance_scale, make_state, step):
        """One CUDA graph per (step kind, shape, guidance scale, 16-bit format, packed-weight storage): `step(st)` is warmed up
        twice on a side stream and then captured; x, t, cond and the per-step scalars live in the static buffers of `st`."""
        weights = next(iter(self.denoise_fn.packed().values())).data_ptr()      # a repack into new storage invalidates the graph
        key = (kind, tuple(img.shape), float(guidance_scale), str(self.denoise_fn.compute_dtype), weights)
        g = self._graphs.get(key)
        if g is None:
            b = img.shape[0]
            st = dict(x=img.clone(), t=torch.zeros(b, dtype=torch.long, device=img.device), cond=cond.clone().float(),
                      noise=torch.zeros_like(img))
            st["weights"] = self.denoise_fn.packed()        # keeps the captured operand storage alive (and its address unique)
            make_state(st)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    st["out"] = step(st)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                st["out"] = step(st)
            g = (graph, st)
            self._graphs[key] = g
        graph, st = g
        st["x"].copy_(img)
        st["cond"].copy_(cond)
        return graph, st

    def _graph_loop(self, img, cond, guidance_scale):
        """The ancestral loop replayed from one CUDA graph per step shape: t, x and the noise live in static buffers."""
        graph, st = self._step_graph("ancestral", img, cond, guidance_scale, lambda st: None,
                                     lambda st: self._p_sample_core(st["x"], st["t"], st["cond"], guidance_scale, st["noise"]))
        for i in reversed(range(0, self.num_timesteps)):
            st["t"].fill_(i)
            st["noise"].normal_()
            graph.replay()
            st["x"].copy_(st["out"])
        return st["x"].clone()

    def _ddim_pairs(self):
        """(time, time_next) pairs of VDDP:990-992."""
        times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)
        times = list(reversed(times.int().tolist()))
        return list(zip(times[:-1], times[1:]))

    def _ddim_step_core(self, st, guidance_scale):
        """One eta = 0 DDIM update with every per-step scalar read on the device (graph-safe): the table `acp_next` holds
        alphas_cumprod shifted by one with 1.0 in front, so time_next = -1 gives (1, 0) and the step returns x0 (VDDP:1010-1012)."""
        x, t = st["x"], st["t"]
        b, c, f, h, w = x.shape
        eps_cl, has_null = self._eps_channels_last(x, t, st["cond"], guidance_scale)
        ops.cfg_x0(x, eps_cl, has_null, guidance_scale, self.sqrt_recip_alphas_cumprod[t].contiguous(),
                   self.sqrt_recipm1_alphas_cumprod[t].contiguous(), st["x0"], st["eps"], b, c, f, h, w)
        an = st["acp_next"][st["tn"] + 1]
        out = torch.empty_like(x)
        if self.ddim_sampling_eta == 0.:
            # x0 * sqrt(an) + eps * sqrt(1 - an): the posterior kernel's c1 * x0 + c2 * x + sig * noise with sig = 0 and no clamp
            ops.posterior_step(st["x0"], st["eps"], st["x0"], None, an.sqrt().contiguous(), (1. - an).sqrt().contiguous(), st["zeros"],
                               out, b, c * f * h * w)
        else:
            # eta > 0 (VDDP:1006-1016): sigma = eta sqrt((1 - a / an)(1 - an) / (1 - a)), and the step's noise enters through sig;
            # at time_next = -1 the table gives an = 1, hence sigma = 0 and sqrt(1 - an - sigma^2) = 0: the step returns x0
            a = self.alphas_cumprod[t]
            sigma = self.ddim_sampling_eta * ((1. - a / an) * (1. - an) / (1. - a)).sqrt()
            ops.posterior_step(st["x0"], st["eps"], st["noise"], None, an.sqrt().contiguous(),
                               (1. - an - sigma * sigma).sqrt().contiguous(), sigma.contiguous(), out, b, c * f * h * w)
        return out

    def _graph_ddim(self, img, cond, guidance_scale):
        def make_state(st):
            b = img.shape[0]
            st["tn"] = torch.zeros(b, dtype=torch.long, device=img.device)
            st["x0"], st["eps"] = torch.empty_like(img), torch.empty_like(img)
            st["zeros"] = torch.zeros(b, dtype=torch.float32, device=img.device)
            st["acp_next"] = torch.cat((torch.ones(1, device=img.device), self.alphas_cumprod)).contiguous()

        graph, st = self._step_graph(f"ddim eta={self.ddim_sampling_eta}", img, cond, guidance_scale, make_state,
                                     lambda st: self._ddim_step_core(st, guidance_scale))
        for time, time_next in self._ddim_pairs():
            st["t"].fill_(time)
            st["tn"].fill_(time_next)
            if time_next >= 0:
                st["noise"].normal_()                    # one draw per step as in the reference (multiplied by sigma = 0 when eta = 0)
            graph.replay()
            st["x"].copy_(st["out"])
        return st["x"].clone()

    @torch.inference_mode()
    def sample(self, cond=None, batch_size=16, guidance_scale=1.):
        """VDDP:977-984."""
        batch_size = cond.shape[0] if cond is not None else batch_size
        fn = self.p_sample_loop if not self.is_ddim_sampling else self.ddim_sample
        return fn((batch_size, self.channels, self.num_frames, self.image_size, self.image_size), cond=cond, guidance_scale=guidance_scale)

    @torch.inference_mode()
    def ddim_sample(self, shape, cond=None, guidance_scale=1.):
        """VDDP:986-1018: no clamp, no dynamic threshold; with the default eta = 0 the drawn noise is multiplied by sigma = 0."""
        batch, device = shape[0], self.betas.device
        pairs = self._ddim_pairs()
        img = torch.randn(shape, device=device)
        b, c, f, h, w = shape
        if self.use_cuda_graph:
            return unnormalize_img(self._graph_ddim(img, cond, guidance_scale))
        for time, time_next in pairs:
            t = torch.full((batch,), time, device=device, dtype=torch.long)
            eps_cl, has_null = self._eps_channels_last(img, t, cond, guidance_scale)
            x0 = torch.empty_like(img)
            eps = torch.empty_like(img)
            ops.cfg_x0(img, eps_cl, has_null, guidance_scale, self.sqrt_recip_alphas_cumprod[t].contiguous(),
                       self.sqrt_recipm1_alphas_cumprod[t].contiguous(), x0, eps, b, c, f, h, w)
            if time_next < 0:
                img = x0
                continue
            out = torch.empty_like(img)
            if self.ddim_sampling_eta == 0.:
                an = float(self.alphas_cumprod[time_next])
                torch.randn_like(img)                   # the reference draws (and discards, sigma = 0) one noise per step
                ops.axpby(x0, eps, an ** 0.5, (1 - an) ** 0.5, 0.0, out)
            else:                                       # VDDP:1003-1016 in fp32 as the reference computes it
                a, an = self.alphas_cumprod[time], self.alphas_cumprod[time_next]
                sigma = self.ddim_sampling_eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
                coef = lambda v: v.reshape(1).expand(b).contiguous()
                noise = torch.randn_like(img)
                ops.posterior_step(x0, eps, noise, None, coef(an.sqrt()), coef((1 - an - sigma ** 2).sqrt()), coef(sigma), out, b,
                                   c * f * h * w)
            img = out
        return unnormalize_img(img)

    @torch.inference_mode()
    def interpolate(self, x1, x2, t=None, lam=0.5, cond=None, guidance_scale=1.):
        """VDDP:1020-1034: noise both clips to step t (two q_sample draws), blend them, denoise from t - 1 down to 0.  The reference
        calls p_sample without a conditioning, which its own per-frame-conditioned network rejects (VDDP:753); `cond` and
        `guidance_scale` are therefore accepted here (keyword extensions) and `cond=None` fails the same way, with a ValueError."""
        b = x1.shape[0]
        t = self.num_timesteps - 1 if t is None else t
        assert x1.shape == x2.shape
        if cond is None:
            raise ValueError("cond is required (per_frame_cond=True): pass the (b, frames) conditioning of the interpolated clip")
        t_batched = torch.full((b,), t, device=x1.device, dtype=torch.long)
        xt1, xt2 = (self.q_sample(x, t=t_batched) for x in (x1, x2))
        img = (1 - lam) * xt1 + lam * xt2
        for i in reversed(range(0, t)):
            img = self.p_sample(img, torch.full((b,), i, device=x1.device, dtype=torch.long), cond=cond, guidance_scale=guidance_scale)
        return img

    # ------------------------------------------------------------------ training
    def p_losses(self, x_start, t, cond=None, noise=None, **kwargs):
        """VDDP:1044-1060: loss between the drawn noise and the network's prediction on q_sample(x_start, t, noise)."""
        noise = noise if noise is not None else torch.randn_like(x_start)
        null_cond_prob = kwargs.pop('null_cond_prob', 0.)
        kwargs.pop('prob_focus_present', None)
        kwargs.pop('focus_present_mask', None)
        b = x_start.shape[0]
        mask = blocks.prob_mask_like((b,), null_cond_prob, x_start.device)
        a = self.sqrt_alphas_cumprod[t].contiguous()
        s = self.sqrt_one_minus_alphas_cumprod[t].contiguous()
        return blocks_bwd.training_loss(self.denoise_fn, x_start.contiguous().float(), noise.contiguous().float(), (a, None, s), t, cond,
                                        mask, l2=(self.loss_type == 'l2'))

    def forward(self, x, *args, **kwargs):
        """VDDP:1062-1067."""
        b, device, img_size = x.shape[0], x.device, self.image_size
        if tuple(x.shape[1:]) != (self.channels, self.num_frames, img_size, img_size):
            raise ValueError(f"expected input of shape (b, {self.channels}, {self.num_frames}, {img_size}, {img_size}), got {tuple(x.shape)}")
        t = torch.randint(0, self.num_timesteps, (b,), device=device).long()
        x = normalize_img(x)
        return self.p_losses(x, t, *args, **kwargs)
