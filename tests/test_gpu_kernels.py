"""Parity of each CUDA kernel (called through the C ABI) against the CPU oracle / plain fp32 torch math.

Tolerances: activations are stored in 16 bit, so a kernel whose output is bf16 is checked to 6e-3 relative
L2 (bf16 has 8 mantissa bits: 2^-9 rounding ~ 2e-3 per element) and fp16 to 1e-3; fp32 outputs to 1e-5.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DT = [torch.bfloat16, torch.float16]
TOL = {torch.bfloat16: 6e-3, torch.float16: 1.5e-3}


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def ops():
    from videometamaterials_b200 import ops
    return ops


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("shape", [(2, 11 * 16 * 16, 64, True), (3, 11 * 8 * 8, 16, False), (2, 11 * 12 * 12, 512, True)])
def test_gn_silu_fwd(ops, dt, shape):
    B, pix, C, with_ss = shape
    torch.manual_seed(0)
    x = (torch.randn(B, pix, C, device="cuda") * 2 + 0.5).to(dt)
    gamma = torch.randn(C, device="cuda")
    beta = torch.randn(C, device="cuda")
    ss = torch.randn(B, 2 * C, device="cuda") if with_ss else None
    res = torch.randn(B, pix, C, device="cuda").to(dt)
    xf = x.double().view(B, pix, 8, C // 8)
    stats = torch.stack((xf.sum(dim=(1, 3)), (xf * xf).sum(dim=(1, 3))), dim=-1).contiguous()
    y = torch.empty_like(x)
    ops.gn_silu_fwd(x, y, stats, gamma, beta, ss, B, pix, C, 8, res=res)
    want = F.group_norm(x.float().transpose(1, 2), 8, gamma, beta, eps=1e-5).transpose(1, 2)
    if with_ss:
        want = want * (ss[:, None, :C] + 1) + ss[:, None, C:]
    want = F.silu(want) + res.float()
    assert rel(y, want) < TOL[dt]


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("C", [16, 64, 128, 256, 512])
def test_layernorm_fwd(ops, dt, C):
    torch.manual_seed(1)
    rows = 1000
    x = (torch.randn(rows, C, device="cuda") * 3 + 1).to(dt)
    gamma = torch.randn(C, device="cuda")
    y = torch.empty_like(x)
    ops.ln_fwd(x, y, gamma)
    xf = x.float()
    want = (xf - xf.mean(1, keepdim=True)) / (xf.var(1, unbiased=False, keepdim=True) + 1e-5).sqrt() * gamma
    assert rel(y, want) < TOL[dt]


def _attn_inputs(B, Fr, H, W, heads, dt, seed, T=11):
    torch.manual_seed(seed)
    hd = heads * 32
    qkv = torch.randn(B, Fr, H, W, 3 * hd, device="cuda").to(dt)
    ekv = torch.randn(B, T, 2 * hd, device="cuda")
    return qkv, ekv


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("with_cond", [True, False])
@pytest.mark.parametrize("Fr", [11, 22, 7])        # 11: tensor-core kernel; other frame counts (tokens == frames): csrc/tattn_generic.cu
def test_temporal_attention_core(ops, dt, with_cond, Fr):
    from oracle import vdm_oracle as O
    B, H, W, heads = 2, 5, 7, 8
    hd = heads * 32
    qkv, ekv = _attn_inputs(B, Fr, H, W, heads, dt, 2, T=Fr)
    bias = torch.randn(heads, Fr, Fr, device="cuda")
    freqs = 1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32)).cuda()
    ang = torch.arange(Fr, device="cuda").float()[:, None] * freqs[None, :]
    rot = torch.stack((ang.cos(), ang.sin()), -1).contiguous()
    out = torch.empty(B, Fr, H, W, hd, device="cuda", dtype=dt)
    ops.tattn_fwd(qkv, ekv if with_cond else None, bias, rot, out, B, Fr, H * W, heads)
    # oracle: identity projections; the oracle rotates cond keys itself, the kernel takes them pre-rotated
    q, k, v = (t.float().permute(0, 2, 3, 1, 4).reshape(B, H * W, Fr, heads, 32).transpose(2, 3) for t in qkv.chunk(3, dim=-1))
    k = O.rotary(k, freqs)
    if with_cond:
        ek = ekv[..., :hd].reshape(B, 1, Fr, heads, 32).transpose(2, 3).expand(B, H * W, heads, Fr, 32)
        ev = ekv[..., hd:].reshape(B, 1, Fr, heads, 32).transpose(2, 3).expand(B, H * W, heads, Fr, 32)
        k = torch.cat((ek, k), -2)
        v = torch.cat((ev, v), -2)
    qq = O.rotary(q * 32 ** -0.5, freqs)
    sim = torch.einsum("...id,...jd->...ij", qq, k)
    sim = sim + (torch.cat((bias, bias), -1) if with_cond else bias)
    want = torch.einsum("...ij,...jd->...id", sim.softmax(-1), v).transpose(2, 3).reshape(B, H, W, Fr, hd).permute(0, 3, 1, 2, 4)
    assert rel(out, want) < TOL[dt]


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("hw", [(6, 7), (24, 24)])
def test_linear_attention_core(ops, dt, hw):
    H, W = hw
    B, Fr, heads = 2, 11, 8
    hd = heads * 32
    qkv, ekv = _attn_inputs(B, Fr, H, W, heads, dt, 3)
    n = H * W
    out = torch.empty(B, Fr, H, W, hd, device="cuda", dtype=dt)
    ctx = torch.empty(B * Fr, heads, 32, 32, device="cuda")
    kstat = torch.empty(B * Fr, heads, 32, 2, device="cuda")
    ops.lattn_fwd(qkv, ekv, 11, out, ctx, kstat, B * Fr, Fr, n, heads)
    q, k, v = (t.float().reshape(B * Fr, n, heads, 32).permute(0, 2, 3, 1) for t in qkv.chunk(3, dim=-1))   # (bf, h, d, n)
    ek = ekv[..., :hd].reshape(B, 1, 11, heads, 32).expand(B, Fr, 11, heads, 32).permute(0, 1, 3, 4, 2).reshape(B * Fr, heads, 32, 11)
    ev = ekv[..., hd:].reshape(B, 1, 11, heads, 32).expand(B, Fr, 11, heads, 32).permute(0, 1, 3, 4, 2).reshape(B * Fr, heads, 32, 11)
    k = torch.cat((ek, k), -1).softmax(-1)
    v = torch.cat((ev, v), -1) / n
    q = q.softmax(-2) * 32 ** -0.5
    c = torch.einsum("bhdn,bhen->bhde", k, v)
    want = torch.einsum("bhde,bhdn->bhen", c, q).permute(0, 3, 1, 2).reshape(B, Fr, H, W, hd)
    # softmax weights and values enter the tensor-core MMAs in 16 bit (fp32 accumulation)
    assert rel(ctx, c) < TOL[dt]
    assert rel(out, want) < 1.5 * TOL[dt]


@pytest.mark.parametrize("dt", DT)
def test_spatial_attention_core(ops, dt):
    B, Fr, H, W, heads = 2, 11, 12, 12, 8
    hd = heads * 32
    qkv, ekv = _attn_inputs(B, Fr, H, W, heads, dt, 4)
    n = H * W
    out = torch.empty(B, Fr, H, W, hd, device="cuda", dtype=dt)
    lse = torch.empty(B * Fr, heads, n, device="cuda")
    ops.sattn_fwd(qkv, ekv, out, lse, B * Fr, Fr, n, heads)
    q, k, v = (t.float().reshape(B * Fr, n, heads, 32).transpose(1, 2) for t in qkv.chunk(3, dim=-1))      # (bf, h, n, d)
    ek = ekv[..., :hd].reshape(B * Fr, 1, heads, 32).transpose(1, 2)
    ev = ekv[..., hd:].reshape(B * Fr, 1, heads, 32).transpose(1, 2)
    k = torch.cat((ek, k), -2)
    v = torch.cat((ev, v), -2)
    sim = torch.einsum("bhid,bhjd->bhij", q * 32 ** -0.5, k)
    want = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), v).transpose(1, 2).reshape(B, Fr, H, W, hd)
    assert rel(out, want) < TOL[dt]
    assert rel(lse, torch.logsumexp(sim, -1)) < 1e-5


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("C", [16, 64, 128])
def test_down_up_convs(ops, dt, C):
    torch.manual_seed(5)
    bf, H, W = 5, 16, 16
    x = torch.randn(bf, H, W, C, device="cuda").to(dt)
    wd = (torch.randn(C, C, 4, 4, device="cuda") / (16 * C) ** 0.5).to(dt)
    b = torch.randn(C, device="cuda")
    out = torch.empty(bf, H // 2, W // 2, C, device="cuda", dtype=dt)
    ops.conv_down(x, ops.pack_conv_taps(wd.float(), [C], dt), C, out, bias=b)
    want = F.conv2d(x.float().permute(0, 3, 1, 2), wd.float(), b, stride=2, padding=1).permute(0, 2, 3, 1)
    assert rel(out, want) < TOL[dt]
    wu = (torch.randn(C, C, 1, 4, 4, device="cuda") / (4 * C) ** 0.5).to(dt)
    out2 = torch.empty(bf, 2 * H, 2 * W, C, device="cuda", dtype=dt)
    ops.conv_up(x, ops.pack_conv_up(wu.float(), dt), C, out2, bias=b)
    want2 = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wu[:, :, 0].float(), b, stride=2, padding=1).permute(0, 2, 3, 1)
    assert rel(out2, want2) < TOL[dt]


@pytest.mark.parametrize("dt", DT)
def test_init_conv_with_q_sample(ops, dt):
    torch.manual_seed(6)
    B, C, Fr, H, W, N = 2, 3, 11, 16, 16, 64
    x = torch.rand(B, C, Fr, H, W, device="cuda")
    noise = torch.randn_like(x)
    a = torch.tensor([0.9, 0.3], device="cuda")
    c = torch.tensor([0.1, -0.2], device="cuda")
    s = torch.tensor([0.4, 0.95], device="cuda")
    w = (torch.randn(N, C, 1, 7, 7, device="cuda") / (49 * C) ** 0.5)
    bias = torch.randn(N, device="cuda")
    xin = torch.zeros(B * Fr * H * (W + 6) * 8 + 8, device="cuda", dtype=dt)
    ops.prep_input(x, noise, a, c, s, xin, B, C, Fr, H, W)
    out = torch.empty(B * Fr, H, W, N, device="cuda", dtype=dt)
    ops.init_conv(xin, B * Fr, H, W, ops.pack_init_conv(w, dt), N, out, bias=bias)
    xt = (a.view(-1, 1, 1, 1, 1) * x + c.view(-1, 1, 1, 1, 1) + s.view(-1, 1, 1, 1, 1) * noise).to(dt).float()
    want = F.conv2d(xt.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W), w[:, :, 0].to(dt).float(), bias, padding=3).permute(0, 2, 3, 1)
    assert rel(out, want) < TOL[dt]


def test_sampler_elementwise_and_quantile(ops):
    from oracle import vdm_oracle as O
    from videometamaterials_b200.diffusion import quantile_rank
    torch.manual_seed(7)
    B, C, Fr, H, W = 3, 3, 11, 96, 96
    S = {k: v.cuda() for k, v in O.schedule(256).items()}
    x = torch.randn(B, C, Fr, H, W, device="cuda")
    eps_cl = torch.randn(2 * B, Fr, H, W, C, device="cuda")
    noise = torch.randn_like(x)
    t = torch.tensor([255, 17, 0], device="cuda")
    x0 = torch.empty_like(x)
    eps = torch.empty_like(x)
    ops.cfg_x0(x, eps_cl, True, 5.0, S["sqrt_recip_alphas_cumprod"][t].contiguous(), S["sqrt_recipm1_alphas_cumprod"][t].contiguous(),
               x0, eps, B, C, Fr, H, W)
    ec, en = (e.permute(0, 4, 1, 2, 3) for e in (eps_cl[:B], eps_cl[B:]))
    eps_want = en + (ec - en) * 5.0
    assert rel(eps, eps_want) < 1e-6          # fused multiply-add vs torch's separate ops: 1 ulp
    x0_want = O.predict_x0(S, x, t, eps)
    assert rel(x0, x0_want) < 1e-6
    per = C * Fr * H * W
    k, frac = quantile_rank(per, 0.9)
    s = torch.empty(B, device="cuda")
    ops.abs_quantile(x0, B, per, k, frac, 1.0, s)
    s_want = torch.quantile(x0.flatten(1).abs(), 0.9, dim=-1).clamp(min=1.0)
    assert torch.equal(s, s_want), (s, s_want)     # order statistics + the same lerp: bit exact
    out = torch.empty_like(x)
    c1 = S["posterior_mean_coef1"][t].contiguous()
    c2 = S["posterior_mean_coef2"][t].contiguous()
    sig = ((t != 0).float() * (0.5 * S["posterior_log_variance_clipped"][t]).exp()).contiguous()
    ops.posterior_step(x0, x, noise, s, c1, c2, sig, out, B, per)
    want = O.p_sample_from_eps(S, x, t, eps, noise, dynamic=True)
    assert rel(out, want) < 1e-6


def test_quantile_edge_cases(ops):
    from videometamaterials_b200.diffusion import quantile_rank
    for n in (1, 2, 10, 1000, 4097):
        torch.manual_seed(n)
        v = torch.randn(2, n, device="cuda")
        v[1] = torch.round(v[1] * 2) / 2          # many ties
        k, frac = quantile_rank(n, 0.9)
        s = torch.empty(2, device="cuda")
        ops.abs_quantile(v, 2, n, k, frac, 0.0, s)
        want = torch.quantile(v.abs(), 0.9, dim=-1)
        assert torch.equal(s, want), (n, s, want)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("rows", [148 * 128, 148 * 128 * 3 + 77, 811008 // 8])
def test_fused_qkv_backward(ops, dt, rows):
    """vmm_qkv_bwd (csrc/qkvbwd.cu: dxn = dqkv W and dW += dqkv^T xn in one pass over dqkv, two shared-memory descriptors on the same tile)
    against fp32 torch and against the two launches it replaces (vmm_cgemm + vmm_wgrad); ragged last tile; dW accumulates."""
    torch.manual_seed(rows % 1000)
    dqkv = (torch.randn(rows, 768, device="cuda") * 0.5).to(dt)
    xn = torch.randn(rows, 64, device="cuda").to(dt)
    w = (torch.randn(768, 64, device="cuda") * 0.1)
    wd = ops.pack_linear(w.t(), dt)                      # [64][768] K-major, the data-gradient pack of blocks.pack_all
    dxn = torch.empty(rows, 64, device="cuda", dtype=dt)
    dw0 = torch.randn(768, 64, device="cuda")
    dw = dw0.clone()
    ops.qkv_bwd(dqkv, xn, wd, dxn, dw)
    torch.cuda.synchronize()
    want_dx = dqkv.float() @ wd.float()[:64].t()
    want_dw = dqkv.float().t() @ xn.float()
    assert rel(dxn, want_dx) < (3e-3 if dt == torch.float16 else 6e-3)
    assert rel(dw - dw0, want_dw) < 2e-3
    # the two-kernel path on the same inputs
    dxn2 = torch.empty_like(dxn)
    dw2 = dw0.clone()
    ops.linear_rows([dqkv], wd, 64, dxn2)
    ops.wgrad_linear(dqkv, [xn], dw2)
    assert rel(dxn, dxn2) < 1e-3 and rel(dw - dw0, dw2 - dw0) < 1e-3
    # LayerNorm form: dx = LN'(x)(dxn) + dres and dgamma in the epilogue, against vmm_ln_bwd on the stored rows and against fp32 autograd
    x = torch.randn(rows, 64, device="cuda").to(dt)
    dres = torch.randn(rows, 64, device="cuda").to(dt)
    gamma = torch.rand(64, device="cuda") + 0.5
    dx = torch.empty_like(x)
    dw3, dg = dw0.clone(), torch.zeros(64, device="cuda")
    ops.qkv_ln_bwd(dqkv, xn, wd, x, dres, gamma, dx, dw3, dg)
    dx2, dg2 = torch.empty_like(x), torch.zeros(64, device="cuda")
    ops.ln_bwd(x, dxn2, dres, dx2, gamma, dg2)
    torch.cuda.synchronize()
    assert rel(dw3 - dw0, want_dw) < 2e-3
    xf = x.float().requires_grad_(True)
    gf = gamma.clone().requires_grad_(True)
    y = (xf - xf.mean(1, keepdim=True)) / (xf.var(1, unbiased=False, keepdim=True) + 1e-5).sqrt() * gf
    y.backward(want_dx)
    assert rel(dx, xf.grad + dres.float()) < (3e-3 if dt == torch.float16 else 8e-3)
    assert rel(dg, gf.grad) < 3e-3
    assert rel(dx, dx2) < 6e-3 and rel(dg, dg2) < 3e-3      # (the two-launch path rounds dxn to 16 bits in between)


@pytest.mark.parametrize("B,n", [(4, 304128), (1, 65536), (3, 70001), (8, 304128), (2, 1_200_000)])
def test_quantile_multi_cta_path_is_exact(ops, B, n):
    """n >= 65536: every radix pass is one launch over the whole GPU (csrc/sampler.cu quantile_hist_kernel / quantile_final_kernel).
    Bit-equal to torch.quantile for heavy-tailed values, massive ties, constants, a sample of zeros, several quantiles, and equal to the
    one-CTA kernel (workspace == NULL)."""
    from videometamaterials_b200 import _lib
    from videometamaterials_b200.diffusion import quantile_rank
    torch.manual_seed(n + B)
    v = torch.randn(B, n, device="cuda") * torch.exp(2 * torch.randn(B, 1, device="cuda"))
    if B > 1:
        v[1] = torch.round(v[1] * 4) / 4                       # a few dozen distinct values
    if B > 2:
        v[2] = 0.75
    if B > 3:
        v[3] = 0.0
    for q in (0.9, 0.5, 0.999, 0.0, 1.0):
        k, frac = quantile_rank(n, q)
        s = torch.empty(B, device="cuda")
        ops.abs_quantile(v, B, n, k, frac, 0.0, s)
        want = torch.quantile(v.abs(), q, dim=-1)
        assert torch.equal(s, want), (q, s, want)
        s1 = torch.empty(B, device="cuda")
        _lib.check(_lib.lib.vmm_abs_quantile(v.data_ptr(), B, n, k, frac, 0.0, s1.data_ptr(), None, 0, ops.stream_ptr()))
        assert torch.equal(s1, want)
    # the floor of the dynamic threshold (VDDP:947) and back-to-back calls on one workspace
    k, frac = quantile_rank(n, 0.9)
    s = torch.empty(B, device="cuda")
    for _ in range(3):
        ops.abs_quantile(v, B, n, k, frac, 1.0, s)
    assert torch.equal(s, torch.quantile(v.abs(), 0.9, dim=-1).clamp(min=1.0))


def test_adam_ema_step_matches_oracle(ops):
    """vmm_adam_ema_step (one launch over the flat arena) against the oracle's restatement of torch.optim.Adam + the reference's
    EMA (VDDP:116-129, 1633-1639; the oracle is pinned to torch.optim.Adam in tests/test_cpu_oracle.py), through the copy phase
    and the averaging phase, with a gradient scale.  fp32 in, fp32 out: 1e-6 (fused multiply-adds vs separate torch ops)."""
    from oracle import vdm_oracle as O
    torch.manual_seed(3)
    n = 148 * 256 * 8 + 13                       # more elements than the capped grid covers in one pass, ragged tail
    p0 = torch.randn(n)
    p, m, v, ema = p0.clone(), torch.zeros(n), torch.zeros(n), p0.clone()
    dp, dm, dv, dema = (t.clone().cuda() for t in (p, m, v, ema))
    for step in range(1, 26):
        g = torch.randn(n) * (1 + step % 4)
        mode = 0 if step % 5 else (1 if step < 12 else 2)
        scale = 1.0 if step % 2 else 0.25
        O.adam_ema_step(p, g, m, v, ema, step, lr=1e-3, grad_scale=scale, ema_mode=mode, ema_beta=0.995)
        ops.adam_ema_step(dp, g.cuda(), dm, dv, dema if mode else None, 1e-3, 0.9, 0.999, 1e-8, step, scale, mode, 0.995)
    assert rel(dp.cpu(), p) < 1e-6 and rel(dema.cpu(), ema) < 1e-6
    # beta2 crosses the C ABI as a float: 1.f - 0.999f = 0.00099998713 where torch rounds the double 1 - 0.999 once (0.001), so
    # the second moment runs 1.3e-5 low (measured 1.29e-5), i.e. 6e-6 on the update, far inside the parameter bound above
    assert rel(dm.cpu(), m) < 1e-6 and rel(dv.cpu(), v) < 1e-4
    assert float((dp.cpu() - p).abs().max()) < 1e-5
    with pytest.raises(RuntimeError):
        ops.adam_ema_step(dp, dp, dm, dv, None, 1e-3, 0.9, 0.999, 1e-8, 0, 1.0, 0, 0.995)       # step counts from 1


# ------------------------------------------------------------------------------------------------
# gradients of the GEMM-shaped layers: data gradient (vmm_cgemm with transformed weights) and weight
# gradient (vmm_wgrad), against torch autograd in fp32 on the same 16-bit-rounded operands
# ------------------------------------------------------------------------------------------------
def _conv_grads(x, w, dy, **kw):
    x = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    w = w.float().requires_grad_(True)
    y = F.conv2d(x, w, None, **kw)
    y.backward(dy.float().permute(0, 3, 1, 2))
    return x.grad.permute(0, 2, 3, 1), w.grad


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("shape", [(22, 16, 16, [64], 64), (11, 24, 24, [128, 64], 128), (22, 12, 12, [256], 512), (11, 8, 8, [16, 16], 16)])
def test_conv3x3_dgrad_wgrad(ops, dt, shape):
    bf, H, W, cins, cout = shape
    torch.manual_seed(8)
    cin = sum(cins)
    xs = [torch.randn(bf, H, W, c, device="cuda").to(dt) for c in cins]
    w = (torch.randn(cout, cin, 3, 3, device="cuda") / (9 * cin) ** 0.5).to(dt)
    dy = torch.randn(bf, H, W, cout, device="cuda").to(dt)
    dx_want, dw_want = _conv_grads(torch.cat(xs, -1), w, dy, padding=1)
    # data gradient: flipped taps, transposed channels, output split across the concat sources
    wd = ops.pack_conv_taps(w.float().flip(2, 3).permute(1, 0, 2, 3), [cout], dt)
    dxs = [torch.empty_like(x) for x in xs]
    taps, _ = ops.taps_conv(3, 3, [cout], 1)
    ops.cgemm([dy], [taps], wd, cin, dxs[0], (bf, H, W), out2=dxs[1] if len(xs) > 1 else None, nsplit=cins[0])
    assert rel(torch.cat(dxs, -1), dx_want) < TOL[dt]
    dw = torch.zeros(cout, cin, 1, 3, 3, device="cuda")
    ops.wgrad_conv3x3(dy, xs, dw)
    assert rel(dw[:, :, 0], dw_want) < 2e-3
    db = torch.zeros(cout, device="cuda")
    ops.colsum(dy.reshape(-1, cout), db)
    assert rel(db, dy.float().sum(dim=(0, 1, 2))) < 1e-4


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("mkn", [(1000, 64, 768), (5000, 256, 64), (300, [64, 64], 64), (777, 16, 48)])
def test_linear_dgrad_wgrad(ops, dt, mkn):
    M, K, N = mkn
    ks = K if isinstance(K, list) else [K]
    torch.manual_seed(9)
    xs = [torch.randn(M, k, device="cuda").to(dt) for k in ks]
    ktot = sum(ks)
    w = (torch.randn(N, ktot, device="cuda") / ktot ** 0.5).to(dt)
    dy = torch.randn(M, N, device="cuda").to(dt)
    dx_want = dy.float() @ w.float()
    dw_want = dy.float().t() @ torch.cat(xs, -1).float()
    dxs = [torch.empty_like(x) for x in xs]
    ops.linear_rows([dy], ops.pack_linear(w.float().t(), dt), ktot, dxs[0], out2=dxs[1] if len(xs) > 1 else None, nsplit=ks[0])
    assert rel(torch.cat(dxs, -1), dx_want) < TOL[dt]
    dw = torch.zeros(N, ktot, device="cuda")
    ops.wgrad_linear(dy, xs, dw)
    assert rel(dw, dw_want) < 2e-3


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("C", [16, 64, 128])
def test_down_up_grads(ops, dt, C):
    torch.manual_seed(10)
    bf, H, W = 5, 16, 16
    x = torch.randn(bf, H, W, C, device="cuda").to(dt)
    wd = (torch.randn(C, C, 4, 4, device="cuda") / (16 * C) ** 0.5).to(dt)
    dy = torch.randn(bf, H // 2, W // 2, C, device="cuda").to(dt)
    dx_want, dw_want = _conv_grads(x, wd, dy, stride=2, padding=1)
    # d/dx of the strided conv = transposed-conv form with weight viewed as (cin_t = cout, cout_t = cin)
    dx = torch.empty_like(x)
    ops.conv_up(dy, ops.pack_conv_taps(wd.float().permute(1, 0, 2, 3), [C], dt), C, dx)
    assert rel(dx, dx_want) < TOL[dt]
    dw = torch.zeros(C, C, 1, 4, 4, device="cuda")
    ops.wgrad_down(dy, x, dw)
    assert rel(dw[:, :, 0], dw_want) < 2e-3
    # transposed conv
    wu = (torch.randn(C, C, 1, 4, 4, device="cuda") / (4 * C) ** 0.5).to(dt)
    dyu = torch.randn(bf, 2 * H, 2 * W, C, device="cuda").to(dt)
    xg = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wg = wu[:, :, 0].float().requires_grad_(True)
    F.conv_transpose2d(xg, wg, None, stride=2, padding=1).backward(dyu.float().permute(0, 3, 1, 2))
    dxu = torch.empty_like(x)
    ops.conv_down(dyu, ops.pack_conv_taps(wu[:, :, 0].float(), [C], dt), C, dxu)
    assert rel(dxu, xg.grad.permute(0, 2, 3, 1)) < TOL[dt]
    dwu = torch.zeros(C, C, 1, 4, 4, device="cuda")
    ops.wgrad_up(dyu, x, dwu)
    assert rel(dwu[:, :, 0], wg.grad) < 2e-3


@pytest.mark.parametrize("dt", DT)
def test_init_conv_wgrad(ops, dt):
    torch.manual_seed(11)
    B, C, Fr, H, W, N = 2, 3, 11, 16, 16, 64
    x = torch.randn(B, C, Fr, H, W, device="cuda")
    xin = torch.zeros(B * Fr * H * (W + 6) * 8 + 8, device="cuda", dtype=dt)
    ops.prep_input(x, None, None, None, None, xin, B, C, Fr, H, W)
    dy = torch.randn(B * Fr, H, W, N, device="cuda").to(dt)
    xg = x.to(dt).float().permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W)
    w = torch.zeros(N, C, 7, 7, device="cuda", requires_grad=True)
    F.conv2d(xg, w, None, padding=3).backward(dy.float().permute(0, 3, 1, 2))
    dw = torch.zeros(N, C, 1, 7, 7, device="cuda")
    ops.wgrad_init_conv(dy, xin, dw, C)
    assert rel(dw[:, :, 0], w.grad) < 2e-3


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("shape", [(2, 11 * 16 * 16, 64, True), (3, 11 * 8 * 8, 16, False), (2, 11 * 12 * 12, 512, True)])
def test_gn_silu_bwd(ops, dt, shape):
    B, pix, C, with_ss = shape
    torch.manual_seed(12)
    x = (torch.randn(B, pix, C, device="cuda") * 2 + 0.5).to(dt)
    dy = torch.randn(B, pix, C, device="cuda").to(dt)
    gamma = torch.randn(C, device="cuda").requires_grad_(True)
    beta = torch.randn(C, device="cuda").requires_grad_(True)
    ss = torch.randn(B, 2 * C, device="cuda").requires_grad_(True) if with_ss else None
    xf = x.float().requires_grad_(True)
    y = F.group_norm(xf.transpose(1, 2), 8, gamma, beta, eps=1e-5).transpose(1, 2)
    if with_ss:
        y = y * (ss[:, None, :C] + 1) + ss[:, None, C:]
    F.silu(y).backward(dy.float())
    xd = x.double().view(B, pix, 8, C // 8)
    stats = torch.stack((xd.sum(dim=(1, 3)), (xd * xd).sum(dim=(1, 3))), dim=-1).contiguous()
    dx = torch.empty_like(x)
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    dss = torch.zeros(B, 2 * C, device="cuda") if with_ss else None
    dcs = torch.full((C,), 0.25, device="cuda")      # accumulated into (+=): the conv bias gradient = column sums of dx
    ops.gn_silu_bwd(x, dy, dx, stats, gamma.detach(), beta.detach(), ss.detach() if with_ss else None, B, pix, C, 8, dg, db, dss,
                    dx_colsum=dcs)
    assert rel(dx, xf.grad) < TOL[dt]
    # per channel the column sum of dx nearly cancels (GroupNorm removes the group mean): compare against the sum of the
    # kernel's own 16-bit dx plus an absolute tolerance from its rounding
    want = dx.float().sum(dim=(0, 1)) + 0.25
    assert float((dcs - want).abs().max()) < 2e-2 * float(dx.float().abs().sum(dim=(0, 1)).max()) * TOL[dt] + 1e-3
    assert rel(dg, gamma.grad) < 2e-3 and rel(db, beta.grad) < 2e-3
    if with_ss:
        assert rel(dss, ss.grad) < 2e-3


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("C", [16, 64, 256, 512])
def test_layernorm_bwd(ops, dt, C):
    torch.manual_seed(13)
    rows = 1003
    x = (torch.randn(rows, C, device="cuda") * 3 + 1).to(dt)
    dy = torch.randn(rows, C, device="cuda").to(dt)
    dres = torch.randn(rows, C, device="cuda").to(dt)
    gamma = torch.randn(C, device="cuda").requires_grad_(True)
    xf = x.float().requires_grad_(True)
    y = (xf - xf.mean(1, keepdim=True)) / (xf.var(1, unbiased=False, keepdim=True) + 1e-5).sqrt() * gamma
    y.backward(dy.float())
    dx = torch.empty_like(x)
    dg = torch.zeros(C, device="cuda")
    ops.ln_bwd(x, dy, dres, dx, gamma.detach(), dg)
    assert rel(dx, xf.grad + dres.float()) < TOL[dt]
    assert rel(dg, gamma.grad) < 2e-3


# ------------------------------------------------------------------------------------------------
# attention cores, backward: against torch autograd of the same fp32 math
# ------------------------------------------------------------------------------------------------
def _rot_tables(Fr):
    freqs = 1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32)).cuda()
    ang = torch.arange(Fr, device="cuda").float()[:, None] * freqs[None, :]
    return freqs, torch.stack((ang.cos(), ang.sin()), -1).contiguous()


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("with_cond", [True, False])
@pytest.mark.parametrize("Fr", [11, 22, 7])
def test_temporal_attention_bwd(ops, dt, with_cond, Fr):
    from oracle import vdm_oracle as O
    B, H, W, heads = 2, 5, 7, 8
    hd = heads * 32
    qkv, ekv = _attn_inputs(B, Fr, H, W, heads, dt, 20, T=Fr)
    dout = torch.randn(B, Fr, H, W, hd, device="cuda").to(dt)
    bias = torch.randn(heads, Fr, Fr, device="cuda")
    freqs, rot = _rot_tables(Fr)
    qf = qkv.float().requires_grad_(True)
    ef = ekv.clone().requires_grad_(True)
    bf_ = bias.clone().requires_grad_(True)
    q, k, v = (t.permute(0, 2, 3, 1, 4).reshape(B, H * W, Fr, heads, 32).transpose(2, 3) for t in qf.chunk(3, dim=-1))
    k = O.rotary(k, freqs)
    if with_cond:
        ek = ef[..., :hd].reshape(B, 1, Fr, heads, 32).transpose(2, 3).expand(B, H * W, heads, Fr, 32)
        ev = ef[..., hd:].reshape(B, 1, Fr, heads, 32).transpose(2, 3).expand(B, H * W, heads, Fr, 32)
        k = torch.cat((ek, k), -2)
        v = torch.cat((ev, v), -2)
    sim = torch.einsum("...id,...jd->...ij", O.rotary(q * 32 ** -0.5, freqs), k)
    sim = sim + (torch.cat((bf_, bf_), -1) if with_cond else bf_)
    out = torch.einsum("...ij,...jd->...id", sim.softmax(-1), v).transpose(2, 3).reshape(B, H, W, Fr, hd).permute(0, 3, 1, 2, 4)
    out.backward(dout.float())
    dqkv = torch.empty_like(qkv)
    dekv = torch.zeros_like(ekv)
    dbias = torch.zeros_like(bias)
    ops.tattn_bwd(qkv, ekv if with_cond else None, bias, rot, dout, dqkv, dekv if with_cond else None, dbias, B, Fr, H * W, heads)
    # the tensor-core kernel feeds rotated q / k, P and dS to the MMAs in 16 bit: two roundings more than the fp32 math
    assert rel(dqkv, qf.grad) < 2.5 * TOL[dt]
    assert rel(dbias, bf_.grad) < TOL[dt]
    if with_cond:
        assert rel(dekv, ef.grad) < TOL[dt]


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("hw", [(6, 7), (24, 24)])
def test_linear_attention_bwd(ops, dt, hw):
    H, W = hw
    B, Fr, heads = 2, 11, 8
    hd = heads * 32
    qkv, ekv = _attn_inputs(B, Fr, H, W, heads, dt, 21)
    n = H * W
    dout = torch.randn(B, Fr, H, W, hd, device="cuda").to(dt)
    qf = qkv.float().requires_grad_(True)
    ef = ekv.clone().requires_grad_(True)
    q, k, v = (t.reshape(B * Fr, n, heads, 32).permute(0, 2, 3, 1) for t in qf.chunk(3, dim=-1))
    ek = ef[..., :hd].reshape(B, 1, 11, heads, 32).expand(B, Fr, 11, heads, 32).permute(0, 1, 3, 4, 2).reshape(B * Fr, heads, 32, 11)
    ev = ef[..., hd:].reshape(B, 1, 11, heads, 32).expand(B, Fr, 11, heads, 32).permute(0, 1, 3, 4, 2).reshape(B * Fr, heads, 32, 11)
    kk = torch.cat((ek, k), -1).softmax(-1)
    vv = torch.cat((ev, v), -1) / n
    c = torch.einsum("bhdn,bhen->bhde", kk, vv)
    out = torch.einsum("bhde,bhdn->bhen", c, q.softmax(-2) * 32 ** -0.5).permute(0, 3, 1, 2).reshape(B, Fr, H, W, hd)
    out.backward(dout.float())
    o = torch.empty(B, Fr, H, W, hd, device="cuda", dtype=dt)
    ctx = torch.empty(B * Fr, heads, 32, 32, device="cuda")
    kstat = torch.empty(B * Fr, heads, 32, 2, device="cuda")
    ops.lattn_fwd(qkv, ekv, 11, o, ctx, kstat, B * Fr, Fr, n, heads)
    dctx = torch.empty_like(ctx)
    dqkv = torch.empty_like(qkv)
    dekv = torch.zeros_like(ekv)
    ops.lattn_bwd(qkv, ekv, 11, dout, ctx, kstat, dctx, dqkv, dekv, B * Fr, Fr, n, heads)
    # gradients here are ~1e-5 in magnitude: below fp16's normal range (6e-5), so fp16 storage rounds coarser
    assert rel(dqkv, qf.grad) < 2.5 * TOL[dt]
    assert rel(dekv, ef.grad) < TOL[dt]


@pytest.mark.parametrize("dt", DT)
def test_spatial_attention_bwd(ops, dt):
    B, Fr, H, W, heads = 2, 11, 12, 12, 8
    hd = heads * 32
    qkv, ekv = _attn_inputs(B, Fr, H, W, heads, dt, 22)
    n = H * W
    dout = torch.randn(B, Fr, H, W, hd, device="cuda").to(dt)
    qf = qkv.float().requires_grad_(True)
    ef = ekv.clone().requires_grad_(True)
    q, k, v = (t.reshape(B * Fr, n, heads, 32).transpose(1, 2) for t in qf.chunk(3, dim=-1))
    ek = ef[..., :hd].reshape(B * Fr, 1, heads, 32).transpose(1, 2)
    ev = ef[..., hd:].reshape(B * Fr, 1, heads, 32).transpose(1, 2)
    sim = torch.einsum("bhid,bhjd->bhij", q * 32 ** -0.5, torch.cat((ek, k), -2))
    out = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), torch.cat((ev, v), -2)).transpose(1, 2).reshape(B, Fr, H, W, hd)
    out.backward(dout.float())
    o = torch.empty(B, Fr, H, W, hd, device="cuda", dtype=dt)
    lse = torch.empty(B * Fr, heads, n, device="cuda")
    ops.sattn_fwd(qkv, ekv, o, lse, B * Fr, Fr, n, heads)
    dqkv = torch.empty_like(qkv)
    dekv = torch.zeros_like(ekv)
    ops.sattn_bwd(qkv, ekv, o, dout, lse, dqkv, dekv, B * Fr, n, heads)
    # the kernel uses the 16-bit forward output for rowsum(dO * O): allow the matching rounding
    assert rel(dqkv, qf.grad) < 2 * TOL[dt]
    assert rel(dekv, ef.grad) < 2 * TOL[dt]


# ------------------------------------------------------------------------------------------------
# 3x3 convolutions through ops.conv3x3: on H % 16 == 0, W % 8 == 0 grids vmm_cgemm runs its halo mode (one A slab per kx,
# weights resident in shared memory when they fit, two epilogue groups for 64 < N <= 128); other grids take the generic taps.
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("case", [
    (5, 32, 24, [64], 64, True, True),        # halo, weights resident (72 KB), GroupNorm statistics + residual
    (3, 16, 16, [64, 64], 64, True, False),   # halo, two sources (144 KB of weights: streamed)
    (3, 48, 48, [128], 128, False, True),     # halo, N = 128: two epilogue groups, 4 TMEM accumulators
    (11, 16, 8, [16, 8], 16, True, True),     # halo with partial 64-channel chunks, tiny GroupNorm groups (generic epilogue)
    (4, 24, 24, [64], 64, True, True),        # H % 16 != 0: generic taps
    (2, 96, 96, [64], 128, False, True),      # full-resolution rows
])
def test_conv3x3_forward_modes(ops, dt, case):
    bf, H, W, cins, n, with_res, with_gn = case
    torch.manual_seed(31)
    cin = sum(cins)
    xs = [torch.randn(bf, H, W, c, device="cuda").to(dt) for c in cins]
    w = (torch.randn(n, cin, 3, 3, device="cuda") / (9 * cin) ** 0.5).to(dt)
    b = torch.randn(n, device="cuda")
    r = torch.randn(bf, H, W, n, device="cuda").to(dt) if with_res else None
    fps, groups = 2, 8
    nsamp = -(-bf // fps)
    stats = torch.zeros(nsamp, groups, 2, device="cuda", dtype=torch.float64) if with_gn else None
    out = torch.empty(bf, H, W, n, device="cuda", dtype=dt)
    ops.conv3x3(xs, ops.pack_conv_taps(w.float(), cins, dt), n, out, bias=b, res=r, gn_stats=stats, gn_group=n // groups,
                frames_per_sample=fps)
    want = F.conv2d(torch.cat(xs, -1).float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1)
    if with_res:
        want = want + r.float()
    assert rel(out, want) < TOL[dt]
    if with_gn:
        o = want.double()
        if nsamp * fps != bf:
            o = torch.cat((o, o.new_zeros(nsamp * fps - bf, H, W, n)), 0)
        o = o.view(nsamp, fps, H, W, groups, n // groups)
        ws = torch.stack((o.sum(dim=(1, 2, 3, 5)), (o * o).sum(dim=(1, 2, 3, 5))), -1)
        assert float((stats[..., 1] - ws[..., 1]).abs().max() / ws[..., 1].abs().max()) < 1e-4      # sum of squares
        assert float((stats[..., 0] - ws[..., 0]).abs().max()) < 1e-4 * float(o.abs().sum(dim=(1, 2, 3, 5)).max())


@pytest.mark.parametrize("dt", DT)
def test_conv3x3_dgrad_halo_split(ops, dt):
    """Data gradient of a 3x3 conv over two concatenated sources on a halo-mode grid: flipped / transposed weights, output
    columns split over the two sources (out2 / nsplit), accumulated onto a residual (res / res2)."""
    bf, H, W, cins, cout = 3, 32, 16, [64, 64], 64
    torch.manual_seed(32)
    cin = sum(cins)
    w = (torch.randn(cout, cin, 3, 3, device="cuda") / (9 * cin) ** 0.5).to(dt)
    dy = torch.randn(bf, H, W, cout, device="cuda").to(dt)
    x0 = torch.randn(bf, H, W, cin, device="cuda").to(dt)
    dx_want, _ = _conv_grads(x0, w, dy, padding=1)
    acc = [torch.randn(bf, H, W, c, device="cuda").to(dt) for c in cins]
    wd = ops.pack_conv_taps(w.float().flip(2, 3).permute(1, 0, 2, 3), [cout], dt)
    dxs = [a.clone() for a in acc]
    ops.conv3x3([dy], wd, cin, dxs[0], out2=dxs[1], nsplit=cins[0], res=dxs[0], res2=dxs[1])
    assert rel(torch.cat(dxs, -1), dx_want + torch.cat(acc, -1).float()) < TOL[dt]


@pytest.mark.parametrize("dt", DT)
def test_qkv_projection_rotary_epilogue(ops, dt):
    """vmm_cgemm rotary epilogue (to_qkv of the temporal attention): q columns scaled and rotated, k columns rotated by the
    angle of the row's frame, v columns untouched; checked against the oracle's rotary on the plain projection, and the
    attention core on pre-rotated rows against the core that rotates in place."""
    from oracle import vdm_oracle as O
    B, Fr, HW, heads, C = 2, 11, 35, 8, 64
    hd = heads * 32
    torch.manual_seed(41)
    x = torch.randn(B * Fr * HW, C, device="cuda").to(dt)
    w = (torch.randn(3 * hd, C, device="cuda") / C ** 0.5).to(dt)
    wp = ops.pack_linear(w.float(), dt)
    freqs, rot = _rot_tables(Fr)
    tabs = ops.rotary_tables(rot, 32 ** -0.5)
    plain = torch.empty(B * Fr * HW, 3 * hd, device="cuda", dtype=dt)
    rotd = torch.empty_like(plain)
    ops.linear_rows([x], wp, 3 * hd, plain)
    ops.linear_rows([x], wp, 3 * hd, rotd, rot=(tabs, Fr, HW, 2 * hd, hd))
    y = (x.float() @ w.float().t()).view(B, Fr, HW, 3, heads, 32)
    q, k, v = y[:, :, :, 0], y[:, :, :, 1], y[:, :, :, 2]                       # (B, Fr, HW, heads, 32)
    qr = O.rotary((q * 32 ** -0.5).permute(0, 2, 3, 1, 4), freqs).permute(0, 3, 1, 2, 4)      # rotary over the frame axis
    kr = O.rotary(k.permute(0, 2, 3, 1, 4), freqs).permute(0, 3, 1, 2, 4)
    want = torch.stack((qr, kr, v), dim=3).reshape(B * Fr * HW, 3 * hd)
    assert rel(rotd, want) < TOL[dt]
    assert torch.equal(rotd[:, 2 * hd:], plain[:, 2 * hd:])
    # attention core: pre-rotated rows vs in-place rotary of the plain rows
    ekv = torch.randn(B, 11, 2 * hd, device="cuda")
    bias = torch.randn(heads, Fr, Fr, device="cuda")
    o1 = torch.empty(B * Fr * HW, hd, device="cuda", dtype=dt)
    o2 = torch.empty_like(o1)
    ops.tattn_fwd(plain, ekv, bias, rot, o1, B, Fr, HW, heads)
    ops.tattn_fwd(rotd, ekv, bias, rot, o2, B, Fr, HW, heads, pre_rotated=True)
    assert rel(o2, o1) < TOL[dt]
    dout = torch.randn(B * Fr * HW, hd, device="cuda").to(dt)
    g1, g2 = torch.empty_like(plain), torch.empty_like(plain)
    de1, de2 = torch.zeros_like(ekv), torch.zeros_like(ekv)
    db1, db2 = torch.zeros_like(bias), torch.zeros_like(bias)
    ops.tattn_bwd(plain, ekv, bias, rot, dout, g1, de1, db1, B, Fr, HW, heads)
    ops.tattn_bwd(rotd, ekv, bias, rot, dout, g2, de2, db2, B, Fr, HW, heads, pre_rotated=True)
    assert rel(g2, g1) < 2.5 * TOL[dt] and rel(de2, de1) < TOL[dt] and rel(db2, db1) < TOL[dt]


# ------------------------------------------------------------------------------------------------
# fused temporal-attention block (csrc/ftattn.cu): LayerNorm -> to_qkv (tcgen05) -> rotary -> attention (mma.sync) -> to_out
# (tcgen05) -> + residual, one kernel.  Checked against plain fp32 torch math of VDDP:131-137, 245-264, 396-535 and against the
# unfused kernel chain it replaces (ln_fwd -> cgemm with rotary epilogue -> tattn_fwd -> cgemm with residual).
# ------------------------------------------------------------------------------------------------
def _fused_block_case(ops, dt, B, H, W, with_cond, keep, seed):
    from oracle import vdm_oracle as O
    Fr, heads, Cc = 11, 8, 64
    hd = heads * 32
    torch.manual_seed(seed)
    x = torch.randn(B, Fr, H, W, Cc, device="cuda").to(dt)
    gamma = (1 + 0.2 * torch.randn(Cc, device="cuda")).contiguous()
    wq = torch.randn(3 * hd, Cc, device="cuda") * Cc ** -0.5
    wo = torch.randn(Cc, hd, device="cuda") * hd ** -0.5
    ekv = torch.randn(B, 11, 2 * hd, device="cuda") if with_cond else None
    bias = torch.randn(heads, Fr, Fr, device="cuda")
    freqs = 1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32)).cuda()
    ang = torch.arange(Fr, device="cuda").float()[:, None] * freqs[None, :]
    rot = ops.rotary_tables(torch.stack((ang.cos(), ang.sin()), -1).contiguous(), 32 ** -0.5)
    wqp, wop = ops.pack_linear(wq, dt), ops.pack_linear(wo, dt)
    rows = B * Fr * H * W
    out = torch.empty_like(x)
    xn = torch.empty(rows, Cc, device="cuda", dtype=dt) if keep else None
    qkv = torch.empty(rows, 3 * hd, device="cuda", dtype=dt) if keep else None
    ao = torch.empty(rows, hd, device="cuda", dtype=dt) if keep else None
    ops.ftattn_fwd(x, out, wqp, wop, gamma, ekv, bias, rot, xn, qkv, ao, B, Fr, H * W, heads)
    torch.cuda.synchronize()
    # ---- the unfused kernel chain
    x2 = x.reshape(-1, Cc)
    xn_u = torch.empty_like(x2)
    ops.ln_fwd(x2, xn_u, gamma)
    qkv_u = torch.empty(rows, 3 * hd, device="cuda", dtype=dt)
    ops.linear_rows([xn_u], wqp, 3 * hd, qkv_u, rot=(rot, Fr, H * W, 2 * hd, hd))
    ao_u = torch.empty(rows, hd, device="cuda", dtype=dt)
    ops.tattn_fwd(qkv_u, ekv, bias, rot[1], ao_u, B, Fr, H * W, heads, pre_rotated=True)
    out_u = torch.empty_like(x)
    ops.linear_rows([ao_u], wop, Cc, out_u.reshape(-1, Cc), res=x2)
    # ---- fp32 torch statement (weights rounded to the 16-bit operand format, everything else fp32)
    xf = x.float()
    mean = xf.mean(-1, keepdim=True)
    var = xf.var(-1, unbiased=False, keepdim=True)
    xnf = (xf - mean) / (var + 1e-5).sqrt() * gamma
    qkvf = xnf @ wq.to(dt).float().t()
    q, k, v = (t.permute(0, 2, 3, 1, 4).reshape(B, H * W, Fr, heads, 32).transpose(2, 3) for t in qkvf.chunk(3, dim=-1))
    k = O.rotary(k, freqs)
    if with_cond:
        ek = ekv[..., :hd].reshape(B, 1, 11, heads, 32).transpose(2, 3).expand(B, H * W, heads, 11, 32)
        ev = ekv[..., hd:].reshape(B, 1, 11, heads, 32).transpose(2, 3).expand(B, H * W, heads, 11, 32)
        k = torch.cat((ek, k), -2)
        v = torch.cat((ev, v), -2)
    sim = torch.einsum("...id,...jd->...ij", O.rotary(q * 32 ** -0.5, freqs), k)
    sim = sim + (torch.cat((bias, bias), -1) if with_cond else bias)
    aof = torch.einsum("...ij,...jd->...id", sim.softmax(-1), v).transpose(2, 3).reshape(B, H, W, Fr, hd).permute(0, 3, 1, 2, 4)
    want = aof @ wo.to(dt).float().t() + xf
    e = dict(out_vs_torch=rel(out, want), unfused_vs_torch=rel(out_u, want), out_vs_unfused=rel(out, out_u))
    # the block output is dominated by the residual x: also compare the attention branch alone
    e["branch_vs_torch"] = rel(out.float() - xf, want - xf)
    e["branch_unfused_vs_torch"] = rel(out_u.float() - xf, want - xf)
    if keep:
        e["xn"] = rel(xn, xn_u)
        e["qkv"] = rel(qkv, qkv_u)
        e["ao"] = rel(ao, ao_u)
    return e


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("case", [(2, 5, 7, True, True), (1, 16, 16, True, False), (3, 3, 3, False, True), (2, 12, 11, False, False),
                                  (1, 1, 1, True, True)])
def test_fused_temporal_attention_block(ops, dt, case):
    """Ragged pixel counts (35, 9, 132 = 12 tiles exactly, 256, 1) against 11-pixel tiles; with / without conditioning keys;
    with / without the rows kept for the backward kernels."""
    B, H, W, with_cond, keep = case
    e = _fused_block_case(ops, dt, B, H, W, with_cond, keep, seed=31)
    print("fused temporal block:", dt, case, {k: round(v, 5) for k, v in e.items()})
    assert e["out_vs_torch"] < TOL[dt], e
    # the attention branch alone: qkv and the attention rows pass through 16 bit twice (as in the unfused chain)
    assert e["branch_vs_torch"] < 3 * TOL[dt], e
    assert e["branch_vs_torch"] < 1.5 * e["branch_unfused_vs_torch"] + 1e-4, e
    if keep:
        assert e["xn"] < 1e-6 + (2e-3 if dt == torch.bfloat16 else 3e-4), e        # same formula; rsqrt / rounding ties only
        assert e["qkv"] < TOL[dt] and e["ao"] < TOL[dt], e


def test_fused_temporal_attention_occupancy(ops):
    """One 512-thread CTA per SM (two independent 8-warp groups, 112 KB of shared memory and 256 TMEM columns each): the kernel
    must be launchable with its 225 KB of dynamic shared memory."""
    from videometamaterials_b200 import _lib
    assert _lib.lib.vmm_ftattn_ctas_per_sm() == 1


# ------------------------------------------------------------------------------------------------
# conditioning / time path (csrc/cond.cu): one forward kernel + two backward kernels against the torch statement of the same
# path (blocks.conditioning(_torch_path=True), itself pinned to the oracle on the CPU in tests/test_cpu_host.py)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [(64, (1, 2, 4, 8), 8, [False, True, False, False, True, False, False, False]),
                                  (64, (1, 2, 4, 8), 32, None), (16, (1, 2), 3, [True, False, False]), (64, (1, 2), 1, [False])])
def test_conditioning_kernels_match_the_torch_path(case, monkeypatch):
    from videometamaterials_b200 import Unet3D, blocks
    from videometamaterials_b200.blocks_bwd import get_arena
    dim, mults, B, mask = case
    monkeypatch.setattr(blocks, "COND_KERNEL_MAX_B", 32)      # the product routes batches above 8 to the torch statement (speed); the kernels handle 32
    torch.manual_seed(7)
    model = Unet3D(dim=dim, dim_mults=mults, channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
                   cond_bias=True, cond_attention='self-stacked', cond_attention_tokens=16, use_temporal_attention_cond=True,
                   cond_to_time='add', per_frame_cond=True, padding_mode='zeros').cuda()
    with torch.no_grad():        # parameters that start at zero / one in the reference's init would hide errors
        for k, p in model.named_parameters():
            if "null_text" in k or "cond_token_to_hidden.0" in k or "relative_attention_bias" in k:
                p.add_(0.3 * torch.randn_like(p))
    arena = get_arena(model)
    time = torch.randint(0, 256, (B,), device="cuda")
    cond = torch.rand(B, 11, device="cuda") * 2 - 1
    null = torch.tensor(mask, device="cuda") if mask is not None else (torch.rand(B, device="cuda") < 0.25)
    assert blocks.cond_kernel_eligible(model, time, cond)
    # torch path with autograd
    ss_t, ekv_t, bias_t, rot_t = blocks.conditioning(model, time, cond, null, 11, _torch_path=True)
    gen = torch.Generator(device="cuda").manual_seed(3)
    g_ss = {k: torch.randn(v.shape, device="cuda", generator=gen) for k, v in ss_t.items()}
    g_ekv = {k: torch.randn(v.shape, device="cuda", generator=gen) for k, v in ekv_t.items()}
    g_bias = torch.randn(bias_t.shape, device="cuda", generator=gen)
    arena.zero_grad()
    loss = sum((ss_t[k] * g_ss[k]).sum() for k in ss_t) + sum((ekv_t[k] * g_ekv[k]).sum() for k in ekv_t) + (bias_t * g_bias).sum()
    loss.backward()
    want = arena.flat_grad.clone()
    # kernel path
    arena.zero_grad()
    ss_k, ekv_k, bias_k, rot_k, st = blocks.conditioning_state(model, time, cond, null, 11)
    assert st is not None
    worst = max([rel(ss_k[k], ss_t[k].detach()) for k in ss_t] + [rel(ekv_k[k], ekv_t[k].detach()) for k in ekv_t])
    assert worst < 2e-5, worst
    assert rel(bias_k, bias_t.detach()) < 1e-6 and rel(rot_k, rot_t.detach()) < 1e-6
    d_ss, d_ekv, d_bias = st.grad_views()
    for k in g_ss:
        d_ss[k].copy_(g_ss[k])
    for k in g_ekv:
        d_ekv[k].copy_(g_ekv[k])
    d_bias.copy_(g_bias)
    st.backward()
    torch.cuda.synchronize()
    got = arena.flat_grad
    # per parameter, so that a wrong small tensor cannot hide behind the large ones
    o, errs = 0, {}
    names = {id(p): k for k, p in model.named_parameters()}
    for p in arena.params:
        n = p.numel()
        w = want[o:o + n]
        if float(w.abs().max()) > 0 or float(got[o:o + n].abs().max()) > 0:
            errs[names[id(p)]] = rel(got[o:o + n], w)
        o += n
    bad = {k: v for k, v in errs.items() if v > 2e-4}
    print("conditioning kernels: forward worst", worst, "| parameters with gradient", len(errs), "| worst gradient",
          max(errs.items(), key=lambda kv: kv[1]))
    assert len(errs) > 20 and not bad, bad


# ------------------------------------------------------------------------------------------------
# fused linear-attention block, inference form (csrc/flattn.cu): three kernels, qkv never in HBM.  Against plain fp32 torch math of
# VDDP:131-137, 245-264, 313-378 and against the unfused kernel chain (ln_fwd -> cgemm -> lattn_fwd (5 kernels) -> cgemm).
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("case", [(1, 11, 16, 16), (2, 11, 16, 24), (1, 2, 48, 48)])
def test_fused_linear_attention_block(ops, dt, case):
    B, Fr, H, W = case
    heads, Cc, T = 8, 64, 11
    hd = heads * 32
    n = H * W
    torch.manual_seed(41)
    x = torch.randn(B, Fr, H, W, Cc, device="cuda").to(dt)
    gamma = (1 + 0.2 * torch.randn(Cc, device="cuda")).contiguous()
    wq = torch.randn(3 * hd, Cc, device="cuda") * Cc ** -0.5
    wo = torch.randn(Cc, hd, device="cuda") * hd ** -0.5
    bo = torch.randn(Cc, device="cuda") * 0.1
    ekv = torch.randn(B, T, 2 * hd, device="cuda")
    wqp, wop = ops.pack_linear(wq, dt), ops.pack_linear(wo, dt)
    rows = B * Fr * n
    out = torch.empty_like(x)
    ctx = torch.empty(B * Fr, heads, 32, 32, device="cuda")
    kstat = torch.empty(B * Fr, heads, 32, 2, device="cuda")
    ops.flattn_fwd(x, out, wqp, wop, gamma, bo, ekv, ctx, kstat, B * Fr, Fr, n, heads)
    torch.cuda.synchronize()
    # unfused chain
    x2 = x.reshape(-1, Cc)
    xn = torch.empty_like(x2)
    ops.ln_fwd(x2, xn, gamma)
    qkv = torch.empty(rows, 3 * hd, device="cuda", dtype=dt)
    ops.linear_rows([xn], wqp, 3 * hd, qkv)
    ao = torch.empty(rows, hd, device="cuda", dtype=dt)
    ctx_u = torch.empty_like(ctx)
    kstat_u = torch.empty_like(kstat)
    ops.lattn_fwd(qkv, ekv, T, ao, ctx_u, kstat_u, B * Fr, Fr, n, heads)
    out_u = torch.empty_like(x)
    ops.linear_rows([ao], wop, Cc, out_u.reshape(-1, Cc), bias=bo, res=x2)
    # fp32 torch statement (16-bit weights, everything else fp32)
    xf = x.float()
    mean = xf.mean(-1, keepdim=True)
    var = xf.var(-1, unbiased=False, keepdim=True)
    xnf = (xf - mean) / (var + 1e-5).sqrt() * gamma
    qkvf = (xnf @ wq.to(dt).float().t()).reshape(B * Fr, n, 3 * hd)
    q, k, v = (t.reshape(B * Fr, n, heads, 32).permute(0, 2, 3, 1) for t in qkvf.chunk(3, dim=-1))          # (bf, h, d, n)
    ek = ekv[..., :hd].reshape(B, 1, T, heads, 32).expand(B, Fr, T, heads, 32).permute(0, 1, 3, 4, 2).reshape(B * Fr, heads, 32, T)
    ev = ekv[..., hd:].reshape(B, 1, T, heads, 32).expand(B, Fr, T, heads, 32).permute(0, 1, 3, 4, 2).reshape(B * Fr, heads, 32, T)
    kk = torch.cat((ek, k), -1).softmax(-1)
    vv = torch.cat((ev, v), -1) / n
    c = torch.einsum("bhdn,bhen->bhde", kk, vv)
    aof = torch.einsum("bhde,bhdn->bhen", c, q.softmax(-2) * 32 ** -0.5).permute(0, 3, 1, 2).reshape(B, Fr, H, W, hd)
    want = aof @ wo.to(dt).float().t() + bo + xf
    e = dict(ctx=rel(ctx, c), ctx_unfused=rel(ctx_u, c), out=rel(out, want), branch=rel(out.float() - xf, want - xf),
             branch_unfused=rel(out_u.float() - xf, want - xf), kstat_max=float((kstat[..., 0] - kstat_u[..., 0]).abs().max()))
    print("fused linear block:", dt, case, {k_: round(v_, 5) for k_, v_ in e.items()})
    assert e["ctx"] < 1.5 * TOL[dt] and e["out"] < TOL[dt], e
    assert e["branch"] < 3 * TOL[dt] and e["branch"] < 1.5 * e["branch_unfused"] + 1e-4, e
    # the column maxima come from the fp32 accumulators here and from the 16-bit k rows in the unfused chain
    assert e["kstat_max"] < (0.05 if dt == torch.bfloat16 else 0.01), e
