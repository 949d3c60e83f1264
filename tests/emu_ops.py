"""CPU statements of the remaining forward entry points of the C ABI (include/vmm.h), in plain torch  --  TEST INFRASTRUCTURE.

Each function has the signature of its wrapper in videometamaterials_b200/ops.py and the arithmetic the `-m gpu` kernel tests
check the kernels against (tests/test_gpu_kernels.py).  Together with emu_cgemm they let a CPU test run the product's whole
forward GLUE (packing, views, tap tables, layouts, block wiring in blocks.py) against the oracle without a GPU.  Never imported
by the product; says nothing about the kernels.
"""
import torch
import torch.nn.functional as F

import emu_cgemm


def gn_silu_fwd(x, y, stats, gamma, beta, scale_shift, B, pix, C_, groups, act=True, eps=1e-5, res=None):
    cnt = pix * (C_ // groups)
    mean = (stats[..., 0] / cnt)
    var = stats[..., 1] / cnt - mean * mean
    xf = x.float().reshape(B, pix, groups, C_ // groups)
    v = (xf - mean[:, None, :, None].float()) * (var[:, None, :, None] + eps).rsqrt().float()
    v = v.reshape(B, pix, C_) * gamma + beta
    if scale_shift is not None:
        v = v * (scale_shift[:, None, :C_] + 1) + scale_shift[:, None, C_:]
    if act:
        v = F.silu(v)
    if res is not None:
        v = v + res.float().reshape(B, pix, C_)
    y.copy_(v.reshape(y.shape).to(y.dtype))


def ln_fwd(x2d, y2d, gamma, eps=1e-5):
    xf = x2d.float()
    y2d.copy_(((xf - xf.mean(1, keepdim=True)) / (xf.var(1, unbiased=False, keepdim=True) + eps).sqrt() * gamma).to(y2d.dtype))


def _turn(x, rot):
    """x (..., frames, 32) rotated in interleaved pairs by the per-frame angles rot (frames, 16, 2) = (cos, sin)."""
    xe, xo = x[..., 0::2], x[..., 1::2]
    c, s = rot[..., 0], rot[..., 1]
    return torch.stack((xe * c - xo * s, xo * c + xe * s), dim=-1).flatten(-2)


def tattn_fwd(qkv, ekv, bias, rot, out, B, frames, HW, heads, pre_rotated=False):
    hd = heads * 32
    q, k, v = (t.float().reshape(B, frames, HW, heads, 32).permute(0, 2, 3, 1, 4) for t in qkv.reshape(B * frames * HW, 3 * hd).chunk(3, dim=-1))
    if not pre_rotated:
        q, k = _turn(q * 32 ** -0.5, rot), _turn(k, rot)
    b2 = bias
    if ekv is not None:                                           # cond keys arrive rotated (blocks.conditioning)
        T = ekv.shape[1]
        ek = ekv[..., :hd].float().reshape(B, 1, T, heads, 32).transpose(2, 3).expand(B, HW, heads, T, 32)
        ev = ekv[..., hd:].float().reshape(B, 1, T, heads, 32).transpose(2, 3).expand(B, HW, heads, T, 32)
        k, v = torch.cat((ek, k), -2), torch.cat((ev, v), -2)
        b2 = torch.cat((bias, bias), -1)
    sim = torch.einsum("...id,...jd->...ij", q, k) + b2
    o = torch.einsum("...ij,...jd->...id", sim.softmax(-1), v)                                  # (B, HW, heads, frames, 32)
    out.copy_(o.permute(0, 3, 1, 2, 4).reshape(out.shape).to(out.dtype))


def lattn_fwd(qkv, ekv, T, out, ctx, kstat, BF, frames, HW, heads):
    hd = heads * 32
    B = BF // frames
    q, k, v = (t.float().reshape(BF, HW, heads, 32).permute(0, 2, 3, 1) for t in qkv.reshape(BF * HW, 3 * hd).chunk(3, dim=-1))
    ek = ekv[..., :hd].float().reshape(B, 1, T, heads, 32).expand(B, frames, T, heads, 32).permute(0, 1, 3, 4, 2).reshape(BF, heads, 32, T)
    ev = ekv[..., hd:].float().reshape(B, 1, T, heads, 32).expand(B, frames, T, heads, 32).permute(0, 1, 3, 4, 2).reshape(BF, heads, 32, T)
    k = torch.cat((ek, k), -1).softmax(-1)
    v = torch.cat((ev, v), -1) / HW
    q = q.softmax(-2) * 32 ** -0.5
    c = torch.einsum("bhdn,bhen->bhde", k, v)
    ctx.copy_(c)
    kstat.zero_()                                                 # only the backward kernels read it
    out.copy_(torch.einsum("bhde,bhdn->bhen", c, q).permute(0, 3, 1, 2).reshape(out.shape).to(out.dtype))


def sattn_fwd(qkv, ekv, out, lse, BF, frames, HW, heads):
    hd = heads * 32
    q, k, v = (t.float().reshape(BF, HW, heads, 32).transpose(1, 2) for t in qkv.reshape(BF * HW, 3 * hd).chunk(3, dim=-1))
    ek = ekv[..., :hd].float().reshape(BF, 1, heads, 32).transpose(1, 2)
    ev = ekv[..., hd:].float().reshape(BF, 1, heads, 32).transpose(1, 2)
    k, v = torch.cat((ek, k), -2), torch.cat((ev, v), -2)
    sim = torch.einsum("bhid,bhjd->bhij", q * 32 ** -0.5, k)
    lse.copy_(torch.logsumexp(sim, -1))
    out.copy_(torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), v).transpose(1, 2).reshape(out.shape).to(out.dtype))


def prep_input(x, noise, a, c, s, xin, B, C_, F_, H, W):
    v = x.float()
    if a is not None:
        v = v * a.view(-1, 1, 1, 1, 1)
    if c is not None:
        v = v + c.view(-1, 1, 1, 1, 1)
    if noise is not None:
        v = v + s.view(-1, 1, 1, 1, 1) * noise
    buf = torch.zeros(B * F_, H, W + 6, 8)
    buf[:, :, 3:3 + W, :C_] = v.permute(0, 2, 3, 4, 1).reshape(B * F_, H, W, C_)
    xin[: buf.numel()] = buf.reshape(-1).to(xin.dtype)
    xin[buf.numel():] = 0


def install(monkeypatch, ops):
    """Replace the kernel wrappers of `ops` by these statements (forward path only)."""
    monkeypatch.setattr(ops, "cgemm", emu_cgemm.cgemm)
    for name in ("gn_silu_fwd", "ln_fwd", "tattn_fwd", "lattn_fwd", "sattn_fwd", "prep_input"):
        monkeypatch.setattr(ops, name, globals()[name])
