"""CPU oracle for the VideoMetamaterials hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional, plain-PyTorch fp32 restatement of the reference algorithm for the one path this
repository accelerates: `Unet3D.forward`, `forward_with_guidance_scale`, and the
`GaussianDiffusion` training loss / ancestral / DDIM sampling steps.  "VDDP:n" below means
/root/reference/denoising_diffusion_pytorch/video_denoising_diffusion_pytorch.py line n.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module, and only as the checker (or as the timed CPU arm).  The product package
`videometamaterials_b200` never imports it.

Pinning.  The reference ships no tests, golden vectors or fixtures, so the
oracle is pinned against outputs of the reference itself, produced in the build container by
`oracle/make_golden.py` (which imports the unmodified reference from /root/reference through
the stand-in packages in oracle/shims) and committed under tests/golden/.  One boundary stays
**parity unpinned**: `rotary_embedding_torch` is a third-party dependency that is not vendored
by the reference and not installed here; `rotary()` below restates its published 0.2.x
algorithm (interleaved pairs, theta=10000) and is only checked against the stand-in.

Everything is written as functions over a flat `state_dict` (the reference's own key names), in
the reference's `b c f h w` layout, covering the shipped configuration (model.yaml):
`per_frame_cond=True`, `cond_attention='self-stacked'`, `cond_to_time='add'`,
`use_temporal_attention_cond=True`, `padding_mode='zeros'`; plus the two other values of `padding_mode` in the config
surface ('circular', 'circular_1d'), pinned by tests/golden/padding_modes.pt.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------------------------
@dataclass
class UnetCfg:
    """The subset of Unet3D(...) arguments the shipped path depends on (VDDP:575-595)."""

    dim: int = 64
    dim_mults: Tuple[int, ...] = (1, 2, 4, 8)
    channels: int = 3
    heads: int = 8
    dim_head: int = 32
    groups: int = 8
    init_kernel: int = 7
    frames: int = 11
    padding_mode: str = "zeros"          # 'zeros' (shipped) | 'circular' | 'circular_1d'  (model.yaml:13, VDDP:153-243)
    temporal_cond: bool = True           # use_temporal_attention_cond (model.yaml:21, VDDP:792-795): False = no tokens in temporal attention
    cond_to_time: str = "add"            # 'add' (shipped) | 'concat' (model.yaml:22, VDDP:786-790, 666)

    @property
    def dims(self) -> List[int]:
        return [self.dim] + [self.dim * m for m in self.dim_mults]

    @property
    def levels(self) -> int:
        return len(self.dim_mults)

    @property
    def time_dim(self) -> int:
        return self.dim * 4


# ----------------------------------------------------------------------------------------------
# small pieces
# ----------------------------------------------------------------------------------------------
def rel_pos_buckets(n: int, num_buckets: int = 32, max_distance: int = 32) -> Tensor:
    """(n, n) long bucket index for rel = k_pos - q_pos.  VDDP:82-106 (T5 style, bidirectional)."""
    q = torch.arange(n)[:, None]
    k = torch.arange(n)[None, :]
    neg_rel = q - k                              # "n = -relative_position"  VDDP:85
    half = num_buckets // 2
    out = (neg_rel < 0).long() * half            # VDDP:87-88
    dist = neg_rel.abs()
    exact = half // 2
    log_part = torch.log(dist.float() / exact) / math.log(max_distance / exact) * (half - exact)
    large = (exact + log_part.long()).clamp(max=half - 1)   # VDDP:94-97
    return out + torch.where(dist < exact, dist, large)     # VDDP:99


def time_pos_bias(P: Params, n: int) -> Tensor:
    """(heads, n, n) additive bias.  VDDP:102-108."""
    table = P["time_rel_pos_bias.relative_attention_bias.weight"]     # (32, heads)
    return table[rel_pos_buckets(n).to(table.device)].permute(2, 0, 1)


def sinusoidal_embedding(t: Tensor, dim: int) -> Tensor:
    """VDDP:144-151: sin | cos of t * exp(-i ln(1e4)/(half-1))."""
    half = dim // 2
    rate = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, device=t.device) * -rate)
    arg = t[:, None] * freq[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def rotary(x: Tensor, freqs: Tensor) -> Tensor:
    """rotary_embedding_torch 0.2.x `rotate_queries_or_keys` (call sites VDDP:449,456,471,496).

    x: (..., n, d) with d == 2*len(freqs); position p = index along dim -2.  Pair (x[2i], x[2i+1]) is
    rotated by angle p*freqs[i].  UNPINNED against the real package (not installed here).
    """
    n = x.shape[-2]
    ang = torch.arange(n, device=x.device, dtype=freqs.dtype)[:, None] * freqs[None, :]
    c, s = ang.cos(), ang.sin()
    xe, xo = x[..., 0::2], x[..., 1::2]
    out = torch.empty_like(x)
    out[..., 0::2] = xe * c - xo * s
    out[..., 1::2] = xo * c + xe * s
    return out


def channel_layernorm(x: Tensor, gamma: Tensor, eps: float = 1e-5) -> Tensor:
    """VDDP:251-254: normalise over dim 1 (channels), biased variance, gain only."""
    mu = x.mean(dim=1, keepdim=True)
    var = x.var(dim=1, unbiased=False, keepdim=True)
    return (x - mu) / (var + eps).sqrt() * gamma


def pad_frames(x: Tensor, pad: int, mode: str) -> Tensor:
    """Explicit spatial padding of (..., H, W): 'circular' wraps both pixel axes (nn.Conv3d(padding_mode='circular'), VDDP:240,271),
    'circular_1d' wraps the horizontal axis and zero-fills the vertical one (Circular_1d_Conv3d, VDDP:219-236)."""
    if pad == 0:
        return x
    if mode == "circular":
        return F.pad(x, (pad, pad, pad, pad), mode="circular")
    if mode == "circular_1d":
        return F.pad(F.pad(x, (pad, pad, 0, 0), mode="circular"), (0, 0, pad, pad))
    raise ValueError(mode)


def conv_frames(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int = 1, pad: int = 0, mode: str = "zeros") -> Tensor:
    """nn.Conv3d with kernel depth 1 == the same 2-D conv on every frame (VDDP:271,241,626)."""
    B, C, Fr, H, W = x.shape
    x2 = x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W)
    if mode == "zeros":
        y = F.conv2d(x2, w[:, :, 0], b, stride=stride, padding=pad)
    else:
        y = F.conv2d(pad_frames(x2, pad, mode), w[:, :, 0], b, stride=stride, padding=0)
    return y.reshape(B, Fr, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def convT_frames(x: Tensor, w: Tensor, b: Tensor, mode: str = "zeros") -> Tensor:
    """nn.ConvTranspose3d(dim, dim, (1,4,4), (1,2,2), (0,1,1))  VDDP:155.  The circular variants (CircularUpsample /
    Circular_1d_Upsample, VDDP:164-216) pad the INPUT by k - 1 - p = 2 pixels and let the transposed conv crop
    k - 1 + s + p - 1 = 5 output pixels per side, which leaves exactly 2H x 2W."""
    B, C, Fr, H, W = x.shape
    x2 = x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W)
    if mode == "zeros":
        y = F.conv_transpose2d(x2, w[:, :, 0], b, stride=2, padding=1)
    else:
        y = F.conv_transpose2d(pad_frames(x2, 2, mode), w[:, :, 0], b, stride=2, padding=5)
    return y.reshape(B, Fr, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def conv_key(cfg_mode: str, kind: str) -> str:
    """Infix of a convolution's state_dict key: the reference wraps the conv in a module for 'circular_1d'
    (`.conv.`, VDDP:223) and for both circular upsamplers (`.conv_transpose.`, VDDP:181,204)."""
    if kind == "up":
        return "" if cfg_mode == "zeros" else "conv_transpose."
    return "conv." if cfg_mode == "circular_1d" else ""


# ----------------------------------------------------------------------------------------------
# ResnetBlock  (VDDP:267-311)
# ----------------------------------------------------------------------------------------------
def block(P: Params, pre: str, x: Tensor, groups: int, scale_shift=None, mode: str = "zeros") -> Tensor:
    """conv(1,3,3) -> GroupNorm(groups) over (C/g, f, h, w) -> x*(scale+1)+shift -> SiLU.  VDDP:277-285."""
    k = pre + "proj." + conv_key(mode, "conv")
    x = conv_frames(x, P[k + "weight"], P[k + "bias"], pad=1, mode=mode)
    x = F.group_norm(x, groups, P[pre + "norm.weight"], P[pre + "norm.bias"], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    return F.silu(x)


def resnet_block(P: Params, pre: str, x: Tensor, temb: Optional[Tensor], groups: int, mode: str = "zeros") -> Tensor:
    """VDDP:299-311."""
    ss = None
    if (pre + "mlp.1.weight") in P:
        e = F.linear(F.silu(temb), P[pre + "mlp.1.weight"], P[pre + "mlp.1.bias"])
        e = e[:, :, None, None, None]
        ss = e.chunk(2, dim=1)
    h = block(P, pre + "block1.", x, groups, ss, mode)
    h = block(P, pre + "block2.", h, groups, None, mode)
    if (pre + "res_conv.weight") in P:
        x = conv_frames(x, P[pre + "res_conv.weight"], P[pre + "res_conv.bias"])
    return h + x


# ----------------------------------------------------------------------------------------------
# attention blocks
# ----------------------------------------------------------------------------------------------
def linear_attention(P: Params, pre: str, x: Tensor, tokens: Tensor, heads: int) -> Tensor:
    """SpatialLinearAttention, 'self-stacked' branch with per_frame_cond=False at construction
    (VDDP:679,700 do not pass it) -> all cond tokens are prepended to every frame.  VDDP:331-378."""
    B, C, Fr, H, W = x.shape
    n = H * W
    xf = x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, n)                       # (bf, C, n)
    qkv = torch.einsum("oc,bcn->bon", P[pre + "to_qkv.weight"][:, :, 0, 0], xf)
    q, k, v = (t.reshape(B * Fr, heads, -1, n) for t in qkv.chunk(3, dim=1))  # (bf, h, d, n)
    ek = F.linear(tokens, P[pre + "to_k.weight"])                             # (b, T, h*d)
    ev = F.linear(tokens, P[pre + "to_v.weight"])
    T = tokens.shape[1]

    def to_cols(t):  # (b, T, h*d) -> (b f) h d T
        t = t.reshape(B, 1, T, heads, -1).expand(B, Fr, T, heads, t.shape[-1] // heads)
        return t.permute(0, 1, 3, 4, 2).reshape(B * Fr, heads, -1, T)

    k = torch.cat((to_cols(ek), k), dim=-1)
    v = torch.cat((to_cols(ev), v), dim=-1)
    d = q.shape[2]
    q = q.softmax(dim=-2) * d ** -0.5                                         # VDDP:367,370
    k = k.softmax(dim=-1)                                                     # VDDP:368
    v = v / n                                                                 # VDDP:371
    ctx = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", ctx, q).reshape(B * Fr, -1, n)
    out = torch.einsum("oc,bcn->bon", P[pre + "to_out.weight"][:, :, 0, 0], out) + P[pre + "to_out.bias"][None, :, None]
    return out.reshape(B, Fr, C, H, W).permute(0, 2, 1, 3, 4)


def attention(P: Params, pre: str, x: Tensor, heads: int, *, pos_bias: Optional[Tensor], tokens: Optional[Tensor],
              freqs: Optional[Tensor]) -> Tensor:
    """`Attention.forward` on x of shape (b, b2, n, c).  VDDP:425-535.

    temporal use: b2 = pixels, n = frames, freqs given, pos_bias (h,n,n) given.
    mid spatial use: b2 = frames, n = pixels, no rotary, no bias; one cond token per frame (VDDP:459-462).
    tokens=None reproduces init_temporal_attn (VDDP:436-449,512).
    """
    B, B2, n, _ = x.shape
    qkv = F.linear(x, P[pre + "to_qkv.weight"])
    q, k, v = (t.reshape(B, B2, n, heads, -1).transpose(2, 3) for t in qkv.chunk(3, dim=-1))  # (b,b2,h,n,d)
    d = q.shape[-1]
    if freqs is not None:
        k = rotary(k, freqs)
    if tokens is not None:
        ek = F.linear(tokens, P[pre + "to_k.weight"])          # (b, T, h*d)
        ev = F.linear(tokens, P[pre + "to_v.weight"])
        if pos_bias is None:                                   # spatial: token f belongs to frame f
            ek = ek.reshape(B, B2, 1, heads, d).transpose(2, 3)
            ev = ev.reshape(B, B2, 1, heads, d).transpose(2, 3)
        else:                                                  # temporal: all T tokens for every pixel
            T = tokens.shape[1]
            ek = ek.reshape(B, 1, T, heads, d).transpose(2, 3).expand(B, B2, heads, T, d)
            ev = ev.reshape(B, 1, T, heads, d).transpose(2, 3).expand(B, B2, heads, T, d)
        if freqs is not None:
            ek = rotary(ek, freqs)                             # VDDP:470-471
        k = torch.cat((ek, k), dim=-2)
        v = torch.cat((ev, v), dim=-2)
    q = q * d ** -0.5
    if freqs is not None:
        q = rotary(q, freqs)
    sim = torch.einsum("...id,...jd->...ij", q, k)
    if pos_bias is not None:
        if tokens is not None:
            sim = torch.cat((sim[..., :n] + pos_bias, sim[..., n:] + pos_bias), dim=-1)   # VDDP:503-510
        else:
            sim = sim + pos_bias
    attn = (sim - sim.amax(dim=-1, keepdim=True)).softmax(dim=-1)
    out = torch.einsum("...ij,...jd->...id", attn, v).transpose(2, 3).reshape(B, B2, n, heads * d)
    return F.linear(out, P[pre + "to_out.weight"])


def temporal_block(P: Params, pre: str, x: Tensor, cfg: UnetCfg, bias: Tensor, tokens: Optional[Tensor]) -> Tensor:
    """Residual(PreNorm(EinopsToAndFrom('b c f h w','b (h w) f c', Attention))).  VDDP:615,630,680."""
    B, C, Fr, H, W = x.shape
    xn = channel_layernorm(x, P[pre + "fn.norm.gamma"])
    seq = xn.permute(0, 3, 4, 2, 1).reshape(B, H * W, Fr, C)
    ap = pre + "fn.fn.fn."
    o = attention(P, ap, seq, cfg.heads, pos_bias=bias, tokens=tokens, freqs=P[ap + "rotary_emb.freqs"])
    return o.reshape(B, H, W, Fr, C).permute(0, 4, 3, 1, 2) + x


def mid_spatial_block(P: Params, pre: str, x: Tensor, cfg: UnetCfg, tokens: Tensor) -> Tensor:
    """Residual(PreNorm(EinopsToAndFrom('b c f h w','b f (h w) c', Attention))).  VDDP:687-689."""
    B, C, Fr, H, W = x.shape
    xn = channel_layernorm(x, P[pre + "fn.norm.gamma"])
    seq = xn.permute(0, 2, 3, 4, 1).reshape(B, Fr, H * W, C)
    o = attention(P, pre + "fn.fn.fn.", seq, cfg.heads, pos_bias=None, tokens=tokens, freqs=None)
    return o.reshape(B, Fr, H, W, C).permute(0, 4, 1, 2, 3) + x


def linear_block(P: Params, pre: str, x: Tensor, cfg: UnetCfg, tokens: Tensor) -> Tensor:
    """Residual(PreNorm(SpatialLinearAttention)).  VDDP:679,700."""
    xn = channel_layernorm(x, P[pre + "fn.norm.gamma"])
    return linear_attention(P, pre + "fn.fn.", xn, tokens, cfg.heads) + x


# ----------------------------------------------------------------------------------------------
# conditioning  (VDDP:745-788)
# ----------------------------------------------------------------------------------------------
def conditioning(P: Params, cfg: UnetCfg, time: Tensor, cond: Tensor, null_mask: Tensor) -> Tuple[Tensor, Tensor]:
    """returns (t (b, time_dim), tokens (b, frames, time_dim)).  null_mask: (b,) bool, True = drop the label."""
    e = sinusoidal_embedding(time, cfg.dim)
    e = F.linear(e, P["time_mlp.1.weight"], P["time_mlp.1.bias"])
    e = F.gelu(e)
    t = F.linear(e, P["time_mlp.3.weight"], P["time_mlp.3.bias"])
    tok = F.linear(cond[..., None], P["sign_emb.weight"], P["sign_emb.bias"])           # (b, f, D)
    hid = tok.mean(dim=-2)
    hid = F.layer_norm(hid, hid.shape[-1:], P["cond_token_to_hidden.0.weight"], P["cond_token_to_hidden.0.bias"])
    hid = F.linear(hid, P["cond_token_to_hidden.1.weight"], P["cond_token_to_hidden.1.bias"])
    hid = F.linear(F.silu(hid), P["cond_token_to_hidden.3.weight"], P["cond_token_to_hidden.3.bias"])
    tok = torch.where(null_mask[:, None, None], P["null_text_token"], tok)               # VDDP:772-777
    hid = torch.where(null_mask[:, None], P["null_text_hidden"], hid)                    # VDDP:780-784
    if cfg.cond_to_time == "concat":
        return torch.cat((t, hid), dim=-1), tok                                          # VDDP:789-790
    return t + hid, tok                                                                  # VDDP:788


# ----------------------------------------------------------------------------------------------
# Unet3D.forward  (VDDP:730-821)
# ----------------------------------------------------------------------------------------------
def unet_forward(P: Params, cfg: UnetCfg, x: Tensor, time: Tensor, cond: Tensor, null_mask: Tensor) -> Tensor:
    g, pm = cfg.groups, cfg.padding_mode
    ck, uk = conv_key(pm, "conv"), conv_key(pm, "up")
    bias = time_pos_bias(P, x.shape[2])
    x = conv_frames(x, P["init_conv." + ck + "weight"], P["init_conv." + ck + "bias"], pad=cfg.init_kernel // 2, mode=pm)
    x = temporal_block(P, "init_temporal_attn.", x, cfg, bias, None)       # no conditioning here  VDDP:743
    r = x
    t, tok_all = conditioning(P, cfg, time, cond, null_mask)
    tok = tok_all
    tok_t = tok_all if cfg.temporal_cond else None       # VDDP:792-795: label_emb_token_temporal
    skips = []
    L = cfg.levels
    for i in range(L):
        p = f"downs.{i}."
        x = resnet_block(P, p + "0.", x, t, g, pm)
        x = resnet_block(P, p + "1.", x, t, g, pm)
        x = linear_block(P, p + "2.", x, cfg, tok)
        x = temporal_block(P, p + "3.", x, cfg, bias, tok_t)
        skips.append(x)
        if i < L - 1:
            x = conv_frames(x, P[p + "4." + ck + "weight"], P[p + "4." + ck + "bias"], stride=2, pad=1, mode=pm)   # Downsample VDDP:238-243
    x = resnet_block(P, "mid_block1.", x, t, g, pm)
    x = mid_spatial_block(P, "mid_spatial_attn.", x, cfg, tok)
    x = temporal_block(P, "mid_temporal_attn.", x, cfg, bias, tok_t)
    x = resnet_block(P, "mid_block2.", x, t, g, pm)
    for i in range(L):
        p = f"ups.{i}."
        x = torch.cat((x, skips.pop()), dim=1)
        x = resnet_block(P, p + "0.", x, t, g, pm)
        x = resnet_block(P, p + "1.", x, t, g, pm)
        x = linear_block(P, p + "2.", x, cfg, tok)
        x = temporal_block(P, p + "3.", x, cfg, bias, tok_t)
        if i < L - 1:
            x = convT_frames(x, P[p + "4." + uk + "weight"], P[p + "4." + uk + "bias"], pm)      # Upsample VDDP:153-160
    x = torch.cat((x, r), dim=1)
    x = resnet_block(P, "final_conv.0.", x, None, g, pm)
    return conv_frames(x, P["final_conv.1.weight"], P["final_conv.1.bias"])


def unet_forward_guided(P: Params, cfg: UnetCfg, x: Tensor, time: Tensor, cond: Tensor, w: float) -> Tensor:
    """forward_with_guidance_scale: null + (cond - null) * w; w == 1 short-circuits.  VDDP:715-728."""
    b = x.shape[0]
    keep = torch.zeros(b, dtype=torch.bool, device=x.device)
    logits = unet_forward(P, cfg, x, time, cond, keep)
    if w == 1:
        return logits
    null = unet_forward(P, cfg, x, time, cond, ~keep)
    return null + (logits - null) * w


# ----------------------------------------------------------------------------------------------
# GaussianDiffusion  (VDDP:829-1067)
# ----------------------------------------------------------------------------------------------
SCHEDULE_KEYS = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
    "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
    "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2",
)


def cosine_betas(T: int, s: float = 0.008) -> Tensor:
    """fp64 cosine schedule, clipped to [0, 0.9999].  VDDP:829-839."""
    u = torch.linspace(0, T, T + 1, dtype=torch.float64)
    ac = torch.cos(((u / T) + s) / (1 + s) * torch.pi * 0.5) ** 2
    ac = ac / ac[0]
    return (1 - ac[1:] / ac[:-1]).clip(0, 0.9999)


def schedule(T: int) -> Dict[str, Tensor]:
    """The 12 fp32 buffers of GaussianDiffusion, computed in fp64 then cast.  VDDP:862-900."""
    betas = cosine_betas(T)
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = torch.cat((torch.ones(1, dtype=torch.float64), ac[:-1]))
    pv = betas * (1.0 - ac_prev) / (1.0 - ac)
    vals = dict(
        betas=betas, alphas_cumprod=ac, alphas_cumprod_prev=ac_prev,
        sqrt_alphas_cumprod=ac.sqrt(), sqrt_one_minus_alphas_cumprod=(1.0 - ac).sqrt(),
        log_one_minus_alphas_cumprod=(1.0 - ac).log(), sqrt_recip_alphas_cumprod=(1.0 / ac).sqrt(),
        sqrt_recipm1_alphas_cumprod=(1.0 / ac - 1).sqrt(), posterior_variance=pv,
        posterior_log_variance_clipped=pv.clamp(min=1e-20).log(),
        posterior_mean_coef1=betas * ac_prev.sqrt() / (1.0 - ac),
        posterior_mean_coef2=(1.0 - ac_prev) * alphas.sqrt() / (1.0 - ac),
    )
    return {k: v.to(torch.float32) for k, v in vals.items()}


def _at(buf: Tensor, t: Tensor, ndim: int) -> Tensor:
    """`extract`: gather per-sample coefficient and broadcast.  VDDP:824-827."""
    return buf.to(t.device)[t].reshape(-1, *([1] * (ndim - 1)))


def q_sample(S, x0: Tensor, t: Tensor, noise: Tensor) -> Tensor:
    """VDDP:1036-1042."""
    return _at(S["sqrt_alphas_cumprod"], t, x0.ndim) * x0 + _at(S["sqrt_one_minus_alphas_cumprod"], t, x0.ndim) * noise


def p_losses(P: Params, cfg: UnetCfg, S, x01: Tensor, t: Tensor, cond: Tensor, noise: Tensor, null_mask: Tensor,
             loss_type: str = "l1") -> Tensor:
    """GaussianDiffusion.forward + p_losses with the random draws passed in.  VDDP:1044-1067.
    x01 is the data in [0,1]; it is mapped to [-1,1] first (normalize_img VDDP:1109)."""
    x0 = x01 * 2 - 1
    pred = unet_forward(P, cfg, q_sample(S, x0, t, noise), t, cond, null_mask)
    if loss_type == "l1":
        return (noise - pred).abs().mean()
    if loss_type == "l2":
        return ((noise - pred) ** 2).mean()
    raise NotImplementedError(loss_type)


def predict_x0(S, x: Tensor, t: Tensor, eps: Tensor) -> Tensor:
    """VDDP:920-924."""
    return _at(S["sqrt_recip_alphas_cumprod"], t, x.ndim) * x - _at(S["sqrt_recipm1_alphas_cumprod"], t, x.ndim) * eps


def dynamic_threshold(x0: Tensor, q: float = 0.9) -> Tuple[Tensor, Tensor]:
    """s = max(quantile_q(|x0|) per sample, 1); clamp(x0,-s,s)/s.  torch.quantile 'linear'.  VDDP:938-951."""
    s = torch.quantile(x0.flatten(1).abs(), q, dim=-1).clamp(min=1.0)
    sv = s.reshape(-1, *([1] * (x0.ndim - 1)))
    return x0.clamp(-sv, sv) / sv, s


def p_sample_from_eps(S, x: Tensor, t: Tensor, eps: Tensor, noise: Tensor, dynamic: bool = True) -> Tensor:
    """Everything in p_sample after the network call.  VDDP:936-963."""
    x0 = predict_x0(S, x, t, eps)
    if dynamic:
        x0, _ = dynamic_threshold(x0)
    else:
        x0 = x0.clamp(-1.0, 1.0)
    mean = _at(S["posterior_mean_coef1"], t, x.ndim) * x0 + _at(S["posterior_mean_coef2"], t, x.ndim) * x
    logvar = _at(S["posterior_log_variance_clipped"], t, x.ndim)
    nz = (1 - (t == 0).float()).reshape(-1, *([1] * (x.ndim - 1)))
    return mean + nz * (0.5 * logvar).exp() * noise


def p_sample(P: Params, cfg: UnetCfg, S, x: Tensor, t: Tensor, cond: Tensor, w: float, noise: Tensor,
             dynamic: bool = True) -> Tensor:
    """One ancestral step with classifier-free guidance.  VDDP:956-963."""
    eps = unet_forward_guided(P, cfg, x, t, cond, w)
    return p_sample_from_eps(S, x, t, eps, noise, dynamic)


def p_sample_loop(P: Params, cfg: UnetCfg, S, x_T: Tensor, cond: Tensor, w: float, noises: Sequence[Tensor],
                  dynamic: bool = True) -> Tensor:
    """VDDP:965-975.  noises[k] is the k-th randn_like draw (one per step, including t=0)."""
    T = S["betas"].shape[0]
    img = x_T
    for k, i in enumerate(reversed(range(T))):
        t = torch.full((img.shape[0],), i, dtype=torch.long, device=img.device)
        img = p_sample(P, cfg, S, img, t, cond, w, noises[k], dynamic)
    return (img + 1) * 0.5


def ddim_time_pairs(T: int, steps: int) -> List[Tuple[int, int]]:
    """VDDP:990-992."""
    times = torch.linspace(-1, T - 1, steps=steps + 1)
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


def ddim_step_from_eps(S, x: Tensor, time: int, time_next: int, eps: Tensor, eta: float = 0.0,
                       noise: Optional[Tensor] = None) -> Tensor:
    """One DDIM update after the network call (no clamp / threshold in the reference).  VDDP:1000-1016."""
    t = torch.full((x.shape[0],), time, dtype=torch.long, device=x.device)
    x0 = predict_x0(S, x, t, eps)
    if time_next < 0:
        return x0
    a, an = S["alphas_cumprod"][time], S["alphas_cumprod"][time_next]
    sigma = eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
    c = (1 - an - sigma ** 2).sqrt()
    out = x0 * an.sqrt() + c * eps
    if noise is not None:
        out = out + sigma * noise
    return out


def ddim_sample(P: Params, cfg: UnetCfg, S, x_T: Tensor, cond: Tensor, w: float, steps: int, eta: float = 0.0,
                noises: Optional[Sequence[Tensor]] = None) -> Tensor:
    """VDDP:986-1018.  eta = 0 is the reference default (the drawn noise is multiplied by sigma = 0); with eta > 0, noises[k] is
    the k-th randn_like draw (one per step whose time_next >= 0)."""
    T = S["betas"].shape[0]
    img = x_T
    k = 0
    for time, time_next in ddim_time_pairs(T, steps):
        t = torch.full((img.shape[0],), time, dtype=torch.long, device=img.device)
        eps = unet_forward_guided(P, cfg, img, t, cond, w)
        noise = None
        if eta != 0.0 and time_next >= 0:
            noise = noises[k]
            k += 1
        img = ddim_step_from_eps(S, img, time, time_next, eps, eta, noise)
    return (img + 1) * 0.5


# ----------------------------------------------------------------------------------------------
# optimiser + EMA (Trainer.train, VDDP:1622-1640)
# ----------------------------------------------------------------------------------------------
def adam_ema_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, ema: Optional[Tensor], step: int, *, lr: float = 1e-4,
                  beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, grad_scale: float = 1.0, ema_mode: int = 0,
                  ema_beta: float = 0.995) -> None:
    """One `torch.optim.Adam(lr)` step (VDDP:1478, 1633: default betas / eps, no weight decay, no amsgrad) on flat fp32
    tensors, in place, followed by the reference's model average (VDDP:116-129, 1493-1497): `ema_mode` 1 = the
    `step < step_start_ema` phase (the average is a copy of the weights), 2 = `old * beta + (1 - beta) * new`.
    `step` counts from 1.  `grad_scale` multiplies the gradient first (1 / loss scale; the reference's GradScaler does it)."""
    g = g * grad_scale
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p.addcdiv_(m, denom, value=-lr / bc1)
    if ema is not None and ema_mode == 1:
        ema.copy_(p)
    elif ema is not None and ema_mode == 2:
        ema.copy_(ema * ema_beta + (1 - ema_beta) * p)


# ----------------------------------------------------------------------------------------------
# deterministic synthetic weights (shared by tests, bench and golden generation)
# ----------------------------------------------------------------------------------------------
def unet_param_shapes(cfg: UnetCfg) -> Dict[str, Tuple[int, ...]]:
    """Every entry of the reference `Unet3D.state_dict()` for the shipped configuration, in the
    reference's registration order (checked against the reference by tests/golden)."""
    D, hid, td = cfg.dim, cfg.heads * cfg.dim_head, cfg.time_dim
    S: Dict[str, Tuple[int, ...]] = {}
    S["null_text_token"] = (1, cfg.frames, td)
    S["null_text_hidden"] = (1, td)
    S["time_rel_pos_bias.relative_attention_bias.weight"] = (32, cfg.heads)
    ck, uk = conv_key(cfg.padding_mode, "conv"), conv_key(cfg.padding_mode, "up")
    S["init_conv." + ck + "weight"] = (D, cfg.channels, 1, cfg.init_kernel, cfg.init_kernel)
    S["init_conv." + ck + "bias"] = (D,)

    def temporal(pre, c):
        a = pre + "fn.fn.fn."
        S[a + "rotary_emb.freqs"] = (cfg.dim_head // 2,)
        S[a + "to_qkv.weight"] = (3 * hid, c)
        S[a + "to_q.weight"] = (hid, c)
        S[a + "to_k.weight"] = (hid, td)
        S[a + "to_v.weight"] = (hid, td)
        S[a + "to_out.weight"] = (c, hid)
        S[pre + "fn.norm.gamma"] = (1, c, 1, 1, 1)

    def spatial_full(pre, c):
        a = pre + "fn.fn.fn."
        S[a + "to_qkv.weight"] = (3 * hid, c)
        S[a + "to_q.weight"] = (hid, c)
        S[a + "to_k.weight"] = (hid, td)
        S[a + "to_v.weight"] = (hid, td)
        S[a + "to_out.weight"] = (c, hid)
        S[pre + "fn.norm.gamma"] = (1, c, 1, 1, 1)

    def linear(pre, c):
        a = pre + "fn.fn."
        S[a + "to_qkv.weight"] = (3 * hid, c, 1, 1)
        S[a + "to_q.weight"] = (hid, c, 1, 1)
        S[a + "to_k.weight"] = (hid, td)
        S[a + "to_v.weight"] = (hid, td)
        S[a + "to_out.weight"] = (c, hid, 1, 1)
        S[a + "to_out.bias"] = (c,)
        S[pre + "fn.norm.gamma"] = (1, c, 1, 1, 1)

    def resnet(pre, ci, co, with_time=True):
        if with_time:
            S[pre + "mlp.1.weight"] = (2 * co, 2 * td if cfg.cond_to_time == "concat" else td)        # VDDP:666
            S[pre + "mlp.1.bias"] = (2 * co,)
        for blk, c_in in (("block1.", ci), ("block2.", co)):
            S[pre + blk + "proj." + ck + "weight"] = (co, c_in, 1, 3, 3)
            S[pre + blk + "proj." + ck + "bias"] = (co,)
            S[pre + blk + "norm.weight"] = (co,)
            S[pre + blk + "norm.bias"] = (co,)
        if ci != co:
            S[pre + "res_conv.weight"] = (co, ci, 1, 1, 1)
            S[pre + "res_conv.bias"] = (co,)

    temporal("init_temporal_attn.", D)
    S["time_mlp.1.weight"] = (td, D)
    S["time_mlp.1.bias"] = (td,)
    S["time_mlp.3.weight"] = (td, td)
    S["time_mlp.3.bias"] = (td,)
    chain = (1, 16, 32, 64, 128, td)                     # SignalEmbedding CNN, unused by the shipped path
    for j, (a, b_) in enumerate(zip(chain[:-1], chain[1:])):
        S[f"sign_emb_CNN.emb_model.{2 * j}.weight"] = (b_, a, 4)
        S[f"sign_emb_CNN.emb_model.{2 * j}.bias"] = (b_,)
    S["sign_emb.weight"] = (td, 1)
    S["sign_emb.bias"] = (td,)
    S["cond_token_to_hidden.0.weight"] = (td,)
    S["cond_token_to_hidden.0.bias"] = (td,)
    S["cond_token_to_hidden.1.weight"] = (td, td)
    S["cond_token_to_hidden.1.bias"] = (td,)
    S["cond_token_to_hidden.3.weight"] = (td, td)
    S["cond_token_to_hidden.3.bias"] = (td,)
    dims = cfg.dims
    io = list(zip(dims[:-1], dims[1:]))
    L = len(io)
    for i, (ci, co) in enumerate(io):
        p = f"downs.{i}."
        resnet(p + "0.", ci, co)
        resnet(p + "1.", co, co)
        linear(p + "2.", co)
        temporal(p + "3.", co)
        if i < L - 1:
            S[p + "4." + ck + "weight"] = (co, co, 1, 4, 4)
            S[p + "4." + ck + "bias"] = (co,)
    for i, (ci, co) in enumerate(reversed(io)):
        p = f"ups.{i}."
        resnet(p + "0.", co * 2, ci)
        resnet(p + "1.", ci, ci)
        linear(p + "2.", ci)
        temporal(p + "3.", ci)
        if i < L - 1:
            S[p + "4." + uk + "weight"] = (ci, ci, 1, 4, 4)
            S[p + "4." + uk + "bias"] = (ci,)
    # the reference creates `downs` and `ups` (VDDP:664-665) before the mid blocks (VDDP:684-692)
    mid = dims[-1]
    resnet("mid_block1.", mid, mid)
    spatial_full("mid_spatial_attn.", mid)
    temporal("mid_temporal_attn.", mid)
    resnet("mid_block2.", mid, mid)
    resnet("final_conv.0.", 2 * D, D, with_time=False)
    S["final_conv.1.weight"] = (cfg.channels, D, 1, 1, 1)
    S["final_conv.1.bias"] = (cfg.channels,)
    return S


def _key_seed(key: str, seed: int) -> int:
    h = 1469598103934665603
    for ch in key.encode():
        h = ((h ^ ch) * 1099511628211) % (1 << 63)
    return (h + 7919 * seed) % (1 << 62)


def synthetic_state_dict(cfg: UnetCfg, seed: int = 0) -> Params:
    """Deterministic, machine-independent weights keyed by parameter NAME (not by construction
    order): each tensor is drawn from its own CPU generator seeded with a hash of its key.
    Magnitudes follow torch's default inits (uniform +-1/sqrt(fan_in); norm gains near 1) so that
    activations stay in a realistic range.  The reference ships no checkpoint in-repo; parity is on
    these weights loaded into both implementations."""
    out: Params = {}
    for key, shape in unet_param_shapes(cfg).items():
        g = torch.Generator().manual_seed(_key_seed(key, seed))
        if key.endswith("rotary_emb.freqs"):
            half = shape[0]
            out[key] = 1.0 / (10000 ** (torch.arange(0, 2 * half, 2).float() / (2 * half)))
        elif key.endswith("norm.gamma") or key.endswith("norm.weight") or key == "cond_token_to_hidden.0.weight":
            out[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif key.startswith("null_text") or key.startswith("time_rel_pos_bias"):
            out[key] = torch.randn(shape, generator=g)
        elif key.endswith("bias"):
            out[key] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(max(fan_in, 1))
            out[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return out
