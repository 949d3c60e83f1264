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


# ------------------------------------------------------------------------------------------------
# backward entry points: torch autograd of the forward statements above, written into the caller's buffers with the
# accumulate / overwrite conventions of include/vmm.h
# ------------------------------------------------------------------------------------------------
def colsum(x2d, out):
    out += x2d.float().sum(dim=0)[: out.numel()]


def gn_silu_bwd(x, dy, dx, stats, gamma, beta, scale_shift, B, pix, C_, groups, dgamma, dbeta, dss, act=True, eps=1e-5, dx_colsum=None):
    with torch.enable_grad():
        xf = x.detach().float().reshape(B, pix, C_).requires_grad_(True)
        g_, b_ = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
        ss_ = scale_shift.detach().clone().requires_grad_(True) if scale_shift is not None else None
        y = F.group_norm(xf.transpose(1, 2), groups, g_, b_, eps=eps).transpose(1, 2)
        if ss_ is not None:
            y = y * (ss_[:, None, :C_] + 1) + ss_[:, None, C_:]
        if act:
            y = F.silu(y)
        y.backward(dy.detach().float().reshape(B, pix, C_))
    dx.copy_(xf.grad.reshape(dx.shape).to(dx.dtype))
    dgamma += g_.grad
    dbeta += b_.grad
    if dss is not None:
        dss.copy_(ss_.grad)
    if dx_colsum is not None:
        dx_colsum += xf.grad.sum(dim=(0, 1))


def ln_bwd(x2d, dy2d, dres2d, dx2d, gamma, dgamma, eps=1e-5):
    with torch.enable_grad():
        xf = x2d.detach().float().requires_grad_(True)
        g_ = gamma.detach().clone().requires_grad_(True)
        y = (xf - xf.mean(1, keepdim=True)) / (xf.var(1, unbiased=False, keepdim=True) + eps).sqrt() * g_
        y.backward(dy2d.detach().float())
    d = xf.grad if dres2d is None else xf.grad + dres2d.float()
    dx2d.copy_(d.to(dx2d.dtype))
    dgamma += g_.grad


def qkv_bwd(dqkv2d, xn2d, wd, dxn2d, dw):
    """vmm_qkv_bwd: dxn = dqkv @ W (W^T arrives packed K-major as wd [C][768]) and dW += dqkv^T @ xn."""
    dq = dqkv2d.float()
    dxn2d.copy_((dq @ wd.float()[: xn2d.shape[1]].t()).to(dxn2d.dtype))
    dw += (dq.t() @ xn2d.float()).reshape(dw.shape)


def qkv_ln_bwd(dqkv2d, xn2d, wd, x2d, dres2d, gamma, dx2d, dw, dgamma, eps=1e-5):
    """vmm_qkv_ln_bwd: the same, with the PreNorm backward + residual add applied to the (unrounded) dxn rows."""
    dq = dqkv2d.float()
    dxn = dq @ wd.float()[: xn2d.shape[1]].t()
    dw += (dq.t() @ xn2d.float()).reshape(dw.shape)
    ln_bwd(x2d, dxn, dres2d, dx2d, gamma, dgamma, eps)


def _unturn(x, rot):
    """Inverse of _turn (rotation by the negative angle)."""
    return _turn(x, torch.stack((rot[..., 0], -rot[..., 1]), dim=-1))


def tattn_bwd(qkv, ekv, bias, rot, dout, dqkv, dekv, dbias, B, frames, HW, heads, pre_rotated=False):
    """dqkv is the gradient with respect to the PLAIN projection rows: with pre_rotated the stored q / k rows are rotated (q also
    scaled) and the kernel turns dq / dk back."""
    hd = heads * 32
    rows = qkv.detach().float().reshape(B * frames * HW, 3 * hd)
    if pre_rotated:
        q, k, v = (t.reshape(B, frames, HW, heads, 32) for t in rows.chunk(3, dim=-1))
        r = rot[None, :, None, None]                                  # frames on dim 1 here
        unq = _turn(q, torch.stack((r[..., 0], -r[..., 1]), -1)) / 32 ** -0.5
        unk = _turn(k, torch.stack((r[..., 0], -r[..., 1]), -1))
        rows = torch.cat((unq.reshape(-1, hd), unk.reshape(-1, hd), v.reshape(-1, hd)), dim=-1)
    with torch.enable_grad():
        rows = rows.requires_grad_(True)
        e_ = ekv.detach().float().clone().requires_grad_(True) if ekv is not None else None
        b_ = bias.detach().clone().requires_grad_(True)
        out = torch.empty(B, frames, HW, hd)
        o = _tattn_math(rows, e_, b_, rot, B, frames, HW, heads)
        o.backward(dout.detach().float().reshape(o.shape))
    dqkv.copy_(rows.grad.reshape(dqkv.shape).to(dqkv.dtype))
    if dekv is not None:
        dekv += e_.grad.to(dekv.dtype)
    dbias += b_.grad


def _tattn_math(rows, ekv, bias, rot, B, frames, HW, heads):
    hd = heads * 32
    q, k, v = (t.reshape(B, frames, HW, heads, 32).permute(0, 2, 3, 1, 4) for t in rows.chunk(3, dim=-1))
    q, k = _turn(q * 32 ** -0.5, rot), _turn(k, rot)
    b2 = bias
    if ekv is not None:
        T = ekv.shape[1]
        ek = ekv[..., :hd].reshape(B, 1, T, heads, 32).transpose(2, 3).expand(B, HW, heads, T, 32)
        ev = ekv[..., hd:].reshape(B, 1, T, heads, 32).transpose(2, 3).expand(B, HW, heads, T, 32)
        k, v = torch.cat((ek, k), -2), torch.cat((ev, v), -2)
        b2 = torch.cat((bias, bias), -1)
    sim = torch.einsum("...id,...jd->...ij", q, k) + b2
    o = torch.einsum("...ij,...jd->...id", sim.softmax(-1), v)       # (B, HW, heads, frames, 32)
    return o.permute(0, 3, 1, 2, 4).reshape(B, frames, HW, hd)


def lattn_bwd(qkv, ekv, T, dout, ctx, kstat, dctx, dqkv, dekv, BF, frames, HW, heads):
    hd = heads * 32
    B = BF // frames
    with torch.enable_grad():
        rows = qkv.detach().float().reshape(BF * HW, 3 * hd).requires_grad_(True)
        e_ = ekv.detach().float().clone().requires_grad_(True)
        q, k, v = (t.reshape(BF, HW, heads, 32).permute(0, 2, 3, 1) for t in rows.chunk(3, dim=-1))
        ek = e_[..., :hd].reshape(B, 1, T, heads, 32).expand(B, frames, T, heads, 32).permute(0, 1, 3, 4, 2).reshape(BF, heads, 32, T)
        ev = e_[..., hd:].reshape(B, 1, T, heads, 32).expand(B, frames, T, heads, 32).permute(0, 1, 3, 4, 2).reshape(BF, heads, 32, T)
        c = torch.einsum("bhdn,bhen->bhde", torch.cat((ek, k), -1).softmax(-1), torch.cat((ev, v), -1) / HW)
        o = torch.einsum("bhde,bhdn->bhen", c, q.softmax(-2) * 32 ** -0.5).permute(0, 3, 1, 2)
        o.backward(dout.detach().float().reshape(o.shape))
    dctx.zero_()
    dqkv.copy_(rows.grad.reshape(dqkv.shape).to(dqkv.dtype))
    dekv += e_.grad.to(dekv.dtype)


def sattn_bwd(qkv, ekv, aout, dout, lse, dqkv, dekv, BF, HW, heads):
    hd = heads * 32
    with torch.enable_grad():
        rows = qkv.detach().float().reshape(BF * HW, 3 * hd).requires_grad_(True)
        e_ = ekv.detach().float().clone().requires_grad_(True)
        q, k, v = (t.reshape(BF, HW, heads, 32).transpose(1, 2) for t in rows.chunk(3, dim=-1))
        ek = e_[..., :hd].reshape(BF, 1, heads, 32).transpose(1, 2)
        ev = e_[..., hd:].reshape(BF, 1, heads, 32).transpose(1, 2)
        sim = torch.einsum("bhid,bhjd->bhij", q * 32 ** -0.5, torch.cat((ek, k), -2))
        o = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), torch.cat((ev, v), -2)).transpose(1, 2)
        o.backward(dout.detach().float().reshape(o.shape))
    dqkv.copy_(rows.grad.reshape(dqkv.shape).to(dqkv.dtype))
    dekv += e_.grad.to(dekv.dtype)


def loss_fwd_bwd(pred, target, loss_sum, dpred, B, C_, F_, H, W, l2=False, grad_scale=1.0):
    d = pred.float().reshape(B, F_, H, W, C_) - target.float().permute(0, 2, 3, 4, 1)
    n = d.numel()
    loss_sum += ((d * d).sum() if l2 else d.abs().sum()) / n
    if dpred is not None:
        g = (2.0 * d if l2 else torch.sign(d)) * (grad_scale / n)
        dpred.zero_()
        dpred[:, :C_] = g.reshape(-1, C_).to(dpred.dtype)


def install_training(monkeypatch, ops):
    """Forward and backward wrappers (everything blocks_bwd.training_loss launches)."""
    install(monkeypatch, ops)
    monkeypatch.setattr(ops, "wgrad", emu_cgemm.wgrad)
    for name in ("colsum", "gn_silu_bwd", "ln_bwd", "tattn_bwd", "lattn_bwd", "sattn_bwd", "loss_fwd_bwd", "qkv_bwd", "qkv_ln_bwd"):
        monkeypatch.setattr(ops, name, globals()[name])


# ------------------------------------------------------------------------------------------------
# sampler entry points (fp32 tensors in the reference (B, C, F, H, W) layout; the network output arrives channels-last)
# ------------------------------------------------------------------------------------------------
def cfg_x0(x, eps_cl, has_null, w, sr, srm1, x0, eps_out, B, C_, F_, H, W):
    e = eps_cl.float().reshape(-1, F_, H, W, C_).permute(0, 4, 1, 2, 3)
    ec = e[:B]
    if has_null:
        en = e[B:2 * B]
        ec = en + (ec - en) * w
    if eps_out is not None:
        eps_out.copy_(ec)
    x0.copy_(sr.view(-1, 1, 1, 1, 1) * x - srm1.view(-1, 1, 1, 1, 1) * ec)


def abs_quantile(v, B, n, k, frac, floor_val, s_out):
    srt = v.reshape(B, n).abs().sort(dim=1).values
    lo, hi = srt[:, k], srt[:, min(k + 1, n - 1)]
    s_out.copy_(torch.lerp(lo, hi, torch.tensor(frac)).clamp(min=floor_val))


def posterior_step(x0, x, noise, s, c1, c2, sig, out, B, per):
    xc = x0.reshape(B, per)
    if s is not None:
        xc = torch.minimum(torch.maximum(xc, -s[:, None]), s[:, None]) / s[:, None]
    out.copy_((c1[:, None] * xc + c2[:, None] * x.reshape(B, per) + sig[:, None] * noise.reshape(B, per)).reshape(out.shape))


def axpby(a, b, ca, cb, cc, out):
    out.copy_(ca * a + cb * b + cc)


def install_sampler(monkeypatch, ops):
    install(monkeypatch, ops)
    for name in ("cfg_x0", "abs_quantile", "posterior_step", "axpby"):
        monkeypatch.setattr(ops, name, globals()[name])
