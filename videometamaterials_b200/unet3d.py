"""Unet3D on the B200 kernels, with the reference's constructor, parameter names and state_dict layout.

Reference: VDDP = denoising_diffusion_pytorch/video_denoising_diffusion_pytorch.py.
`Unet3D(...)` VDDP:574-713, `forward` VDDP:730-821, `forward_with_guidance_scale` VDDP:715-728.

The nn.Module tree exists only to own parameters under the reference's names (so checkpoints load both
ways); the arithmetic is issued block by block through `blocks.py` onto the C ABI.
Implemented: the shipped configuration (per_frame_cond=True -> 'self-stacked' attention with 11 per-frame tokens) with every
value of padding_mode, use_temporal_attention_cond and cond_to_time ('add' | 'concat') of the config surface; the
per_frame_cond=False ablation branches (CNN / GRU signal embedding, cross-attention), several of which are broken upstream
(SURVEY.md section 2), raise NotImplementedError.
"""
from __future__ import annotations

import math
from functools import partial
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import blocks
from . import ops


# ------------------------------------------------------------------------------------------------
# parameter containers mirroring the reference's module nesting (no forward of their own)
# ------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            setattr(self, k, v)


class RotaryEmbedding(nn.Module):
    """Parameter-compatible stand-in for rotary_embedding_torch.RotaryEmbedding(dim) (VDDP:612):
    a frozen `freqs` parameter; the rotation itself happens inside the attention kernels."""

    def __init__(self, dim: int, theta: float = 10000.0):
        super().__init__()
        self.freqs = nn.Parameter(1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim)), requires_grad=False)


class _LayerNorm(nn.Module):           # VDDP:245-254
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(1, dim, 1, 1, 1))


class _Attention(nn.Module):           # VDDP:396-423
    def __init__(self, dim, heads, dim_head, cond_dim, rotary_emb=None):
        super().__init__()
        hidden = heads * dim_head
        if rotary_emb is not None:
            self.rotary_emb = rotary_emb
        self.to_qkv = nn.Linear(dim, hidden * 3, bias=False)
        self.to_q = nn.Linear(dim, hidden, bias=False)          # cross-attention only; kept for the state_dict
        self.to_k = nn.Linear(cond_dim, hidden, bias=False)
        self.to_v = nn.Linear(cond_dim, hidden, bias=False)
        # the reference first builds a Conv2d here and immediately replaces it (VDDP:418 vs 421); doing the same
        # keeps the RNG stream, hence the random initialisation for a given seed, identical to the reference's
        self.to_out = nn.Conv2d(hidden, dim, 1)
        self.to_out = nn.Linear(hidden, dim, bias=False)


class _LinearAttention(nn.Module):     # VDDP:313-329
    def __init__(self, dim, heads, dim_head, cond_dim):
        super().__init__()
        hidden = heads * dim_head
        self.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
        self.to_q = nn.Conv2d(dim, hidden, 1, bias=False)
        self.to_k = nn.Linear(cond_dim, hidden, bias=False)
        self.to_v = nn.Linear(cond_dim, hidden, bias=False)
        self.to_out = nn.Conv2d(hidden, dim, 1)


def _conv_module(conv: nn.Module, padding_mode: str) -> nn.Module:
    """The reference wraps its convolutions in a module for 'circular_1d' (Circular_1d_Conv3d.conv, VDDP:219-236), which puts a
    `.conv.` level into their state_dict keys; the other modes register the convolution directly."""
    return _Holder(conv=conv) if padding_mode == 'circular_1d' else conv


class _Block(nn.Module):               # VDDP:267-275
    def __init__(self, dim, dim_out, groups, padding_mode='zeros'):
        super().__init__()
        self.proj = _conv_module(nn.Conv3d(dim, dim_out, (1, 3, 3), padding=(0, 1, 1)), padding_mode)
        self.norm = nn.GroupNorm(groups, dim_out)


class _ResnetBlock(nn.Module):         # VDDP:287-297
    def __init__(self, dim, dim_out, *, time_emb_dim=None, groups=8, padding_mode='zeros'):
        super().__init__()
        if time_emb_dim is not None:
            self.mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_emb_dim, dim_out * 2))
        else:
            self.mlp = None
        self.block1 = _Block(dim, dim_out, groups, padding_mode)
        self.block2 = _Block(dim_out, dim_out, groups, padding_mode)
        self.res_conv = nn.Conv3d(dim, dim_out, 1) if dim != dim_out else nn.Identity()
        self.dim, self.dim_out = dim, dim_out


class _SignalEmbeddingCNN(nn.Module):  # VDDP:538-561 (ablation only; parameters must exist)
    def __init__(self, init_channel, ups):
        super().__init__()
        chain = [init_channel, *ups]
        mods = []
        for a, b in zip(chain[:-1], chain[1:]):
            mods += [nn.Conv1d(a, b, kernel_size=4, stride=2, padding=1), nn.SiLU()]
        self.emb_model = nn.Sequential(*mods)


def _residual_prenorm(dim, fn, wrap_einops):
    inner = _Holder(fn=fn) if wrap_einops else fn       # EinopsToAndFrom adds one more `.fn` level
    pre = _Holder(fn=inner, norm=_LayerNorm(dim))       # PreNorm registers fn before norm (VDDP:257-260)
    return _Holder(fn=pre)                              # Residual


class Unet3D(nn.Module):
    def __init__(self, dim, out_dim=None, dim_mults=(1, 2, 4, 8), channels=3, attn_heads=8, attn_dim_head=32, init_dim=None,
                 init_kernel_size=7, use_sparse_linear_attn=True, resnet_groups=8, cond_bias=False, cond_attention='none',
                 cond_attention_tokens=6, cond_att_GRU=False, use_temporal_attention_cond=False, cond_to_time='add',
                 per_frame_cond=False, padding_mode='zeros', num_frames=11):
        super().__init__()
        unsupported = []
        if not per_frame_cond:
            unsupported.append("per_frame_cond=False (ablation path)")
        if cond_att_GRU:
            unsupported.append("cond_att_GRU=True")
        if cond_to_time not in ('add', 'concat'):
            # (the reference silently drops the label from the time embedding for any other string, VDDP:786-790)
            unsupported.append(f"cond_to_time={cond_to_time!r}")
        if padding_mode not in ('zeros', 'circular', 'circular_1d'):
            raise ValueError(f"unknown padding_mode {padding_mode!r}")
        if not use_sparse_linear_attn:
            unsupported.append("use_sparse_linear_attn=False")
        if attn_dim_head != 32:
            unsupported.append("attn_dim_head != 32")
        if attn_heads != 8:
            unsupported.append("attn_heads != 8 (the attention kernels map one warp to each of 8 heads)")
        if init_kernel_size != 7 or channels > 8:
            unsupported.append("init_kernel_size != 7 or channels > 8")
        if out_dim is not None and out_dim != channels:
            unsupported.append("out_dim != channels")
        if unsupported:
            raise NotImplementedError("videometamaterials_b200.Unet3D implements the shipped model.yaml configuration only; "
                                      "unsupported: " + ", ".join(unsupported))
        self.channels = channels
        time_dim = dim * 4
        self.dim = dim
        self.dim_mults = tuple(dim_mults)
        self.heads = attn_heads
        self.groups = resnet_groups
        self.cond_bias = cond_bias
        self.cond_attention = 'self-stacked'           # VDDP:602
        # VDDP:603 hard-codes 11 tokens (one per frame of the 11-frame clips); `num_frames` is an extension of this package for
        # BASELINE configs[4]'s 22-frame rows: tokens == frames (the reference itself raises at 22 frames, VDDP:777)
        if not 1 <= int(num_frames) <= 24:
            raise NotImplementedError("num_frames must be between 1 and 24 (11: tensor-core attention kernels; others: generic kernels)")
        self.cond_attention_tokens = int(num_frames)
        self.cond_att_GRU = cond_att_GRU
        self.cond_dim = time_dim
        self.use_temporal_attention_cond = use_temporal_attention_cond
        self.cond_to_time = cond_to_time
        self.per_frame_cond = per_frame_cond
        self.padding_mode = padding_mode

        rotary_emb = RotaryEmbedding(min(32, attn_dim_head))
        temporal = lambda d: _residual_prenorm(d, _Attention(d, attn_heads, attn_dim_head, time_dim, rotary_emb), True)
        self.time_rel_pos_bias = _Holder(relative_attention_bias=nn.Embedding(32, attn_heads))
        init_dim = init_dim if init_dim is not None else dim
        pad = init_kernel_size // 2
        self.init_conv = _conv_module(nn.Conv3d(channels, init_dim, (1, init_kernel_size, init_kernel_size), padding=(0, pad, pad)),
                                      padding_mode)
        self.init_temporal_attn = temporal(init_dim)
        dims = [init_dim, *[dim * m for m in dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))
        self.time_mlp = nn.Sequential(nn.Identity(), nn.Linear(dim, time_dim), nn.GELU(), nn.Linear(time_dim, time_dim))
        self.sign_emb_CNN = _SignalEmbeddingCNN(1, (16, 32, 64, 128, time_dim))
        self.sign_emb_GRU = None
        self.sign_emb = nn.Linear(1, time_dim)
        self.cond_token_to_hidden = nn.Sequential(nn.LayerNorm(time_dim), nn.Linear(time_dim, time_dim), nn.SiLU(),
                                                  nn.Linear(time_dim, time_dim))
        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])
        n_res = len(in_out)
        rb = partial(_ResnetBlock, groups=resnet_groups, padding_mode=padding_mode)
        # both circular upsamplers of the reference hold their transposed conv as `.conv_transpose` (VDDP:181, 204)
        up = lambda d: (nn.ConvTranspose3d(d, d, (1, 4, 4), (1, 2, 2), (0, 1, 1)) if padding_mode == 'zeros'
                        else _Holder(conv_transpose=nn.ConvTranspose3d(d, d, (1, 4, 4), (1, 2, 2), (0, 1, 1))))
        # VDDP:666: with 'concat' every ResnetBlock MLP reads time embedding | label embedding
        rbc = partial(rb, time_emb_dim=time_dim * 2 if cond_to_time == 'concat' else time_dim)
        lin = lambda d: _residual_prenorm(d, _LinearAttention(d, attn_heads, 32, time_dim), False)
        for ind, (di, do) in enumerate(in_out):
            last = ind >= n_res - 1
            self.downs.append(nn.ModuleList([
                rbc(di, do), rbc(do, do), lin(do), temporal(do),
                _conv_module(nn.Conv3d(do, do, (1, 4, 4), (1, 2, 2), (0, 1, 1)), padding_mode) if not last else nn.Identity()]))
        mid = dims[-1]
        self.mid_block1 = rbc(mid, mid)
        self.mid_spatial_attn = _residual_prenorm(mid, _Attention(mid, attn_heads, attn_dim_head, time_dim, None), True)
        self.mid_temporal_attn = temporal(mid)
        self.mid_block2 = rbc(mid, mid)
        for ind, (di, do) in enumerate(reversed(in_out)):
            last = ind >= n_res - 1
            self.ups.append(nn.ModuleList([
                rbc(do * 2, di), rbc(di, di), lin(di), temporal(di),
                up(di) if not last else nn.Identity()]))
        self.final_conv = nn.Sequential(rb(dim * 2, dim), nn.Conv3d(dim, channels, 1))
        self.null_text_token = nn.Parameter(torch.randn(1, self.cond_attention_tokens, time_dim))
        self.null_text_hidden = nn.Parameter(torch.randn(1, time_dim))

        self.compute_dtype = torch.bfloat16
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._packed_key = None

    # -------------------------------------------------------------------------------------------
    # packed 16-bit GEMM operands derived from the fp32 master parameters
    # -------------------------------------------------------------------------------------------
    def set_compute_dtype(self, dtype: torch.dtype) -> None:
        assert dtype in (torch.float16, torch.bfloat16)
        if dtype != self.compute_dtype:
            self.compute_dtype = dtype
            self._packed = None

    def repack(self) -> None:
        """Rebuild every packed weight (call after an optimizer step or a load_state_dict).  With a parameter
        arena (training) this is ONE gather+cast launch over the arena; otherwise the torch slicing path."""
        with torch.no_grad():
            dev = next(self.parameters()).device
            if dev.type == "cuda":
                # on a device the parameters always live in a flat arena (created here on first use, ONE concatenation), so that
                # packing is the one-launch gather also at model load and for sampling-only models
                from .blocks_bwd import get_arena
                get_arena(self)
            arena = getattr(self, "_vmm_arena", None)
            if arena is not None and arena.flat_param.device == dev and dev.type == "cuda":
                plan = getattr(self, "_pack_plan", None)
                if plan is None or plan[0].device != dev or plan[3] is not arena:
                    idx, layout = blocks.pack_plan(self, arena)
                    plan = (idx, layout, {}, arena)
                    self._pack_plan = plan
                idx, layout, bufs, _ = plan
                buf = bufs.get(self.compute_dtype)
                if buf is None:
                    buf = torch.empty(idx.numel(), dtype=self.compute_dtype, device=dev)
                    bufs[self.compute_dtype] = buf
                ops.gather_cast(arena.flat_param, idx, buf)
                self._packed = {k: buf[o:o + int(torch.Size(shp).numel())].view(shp) for k, (o, shp) in layout.items()}
            else:
                self._packed = blocks.pack_all(self, self.compute_dtype)
            self._packed_key = self._param_fingerprint()

    def _param_fingerprint(self):
        return tuple(p._version for p in self.parameters()) + (str(self.compute_dtype), str(next(self.parameters()).device))

    def packed(self) -> Dict[str, torch.Tensor]:
        if self._packed is None or self._packed_key != self._param_fingerprint():
            self.repack()
        return self._packed

    # -------------------------------------------------------------------------------------------
    # forward
    # -------------------------------------------------------------------------------------------
    def forward_with_guidance_scale(self, *args, **kwargs):
        """VDDP:715-728.  The conditional and unconditional passes run as ONE batch of 2b (legal: every
        normalisation is per sample; measured batch-invariance 4.9e-7)."""
        guidance_scale = kwargs.pop('guidance_scale', 5.)
        x, time = args[0], args[1]
        cond = kwargs.get('cond')
        if guidance_scale == 1:
            return self.forward(*args, null_cond_prob=0., **kwargs)
        b = x.shape[0]
        mask = torch.cat((torch.zeros(b, dtype=torch.bool, device=x.device), torch.ones(b, dtype=torch.bool, device=x.device)))
        xin = x if (blocks.SHARED_STEM and not (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()))) else torch.cat((x, x))
        eps = self._forward_masked(xin, torch.cat((time, time)), torch.cat((cond, cond)), mask)
        logits, null_logits = eps[:b], eps[b:]
        return null_logits + (logits - null_logits) * guidance_scale

    def forward(self, x, time, cond=None, null_cond_prob=0., focus_present_mask=None, prob_focus_present=0.):
        """VDDP:730-821.  x (b, c, f, h, w) fp32, time (b,) long, cond (b, f) -> (b, c, f, h, w) fp32."""
        # VDDP:738: an absent mask is drawn with prob_focus_present (0 -> all False, no RNG use).  All False is inert (VDDP:514).  With
        # any True entry the REFERENCE ITSELF raises in this configuration: its (n, n) self-mask is applied to the (n, 2n) similarity
        # of the stacked [cond | frame] keys (VDDP:514-524, "size of tensor a (11) must match the size of tensor b (22)"; measured
        # against the unmodified reference), so there is no behaviour to reproduce.
        if prob_focus_present != 0. or (focus_present_mask is not None and bool(torch.as_tensor(focus_present_mask).any())):
            raise RuntimeError("focus_present_mask with True entries / prob_focus_present > 0: the reference raises here under "
                               "cond_attention='self-stacked' (VDDP:514-524: (n, n) mask against the (n, 2n) similarity); only the "
                               "all-False mask of the shipped path is defined")
        if cond is None:
            raise ValueError("cond is required (per_frame_cond=True)")
        batch = x.shape[0]
        mask = blocks.prob_mask_like((batch,), null_cond_prob, x.device)       # VDDP:749 (same RNG consumption)
        return self._forward_masked(x, time, cond, mask)

    def _forward_masked(self, x, time, cond, null_mask):
        if torch.is_grad_enabled() and x.is_cuda and any(p.requires_grad for p in self.parameters()):
            # called under autograd (the reference's nn.Module supports `model(x, t, cond=c).backward()`, VDDP:730-821): route through
            # the block Functions of the training path so that the call is differentiable instead of silently detached
            from . import blocks_bwd
            eps_cl = blocks_bwd.unet_forward_autograd(self, x, time, cond, null_mask)
        else:
            eps_cl = blocks.unet_forward(self, x, None, None, time, cond, null_mask)   # (b, f, h, w, c) fp32
        return eps_cl.permute(0, 4, 1, 2, 3).contiguous()
