"""The host side of the implicit-GEMM formulation on the CPU: every convolution of the network is expressed for `vmm_cgemm` as
views + tap tables + a packed weight (ops.py / blocks.pack_all).  With `ops.cgemm` replaced by a torch statement of the kernel's
documented contract (tests/emu_cgemm.py) the front-ends must reproduce torch's own convolutions, forward and data gradient:
3x3 with concatenated sources (VDDP:271, 813), 1x1 / linear rows, Downsample (VDDP:241) through parity views, Upsample
(VDDP:155) through four output phases, init_conv (VDDP:626) through the 8 px x 8 ch row taps of the prepared input."""
import pytest
import torch
import torch.nn.functional as F

import emu_cgemm


@pytest.fixture
def ops(monkeypatch):
    from videometamaterials_b200 import ops
    monkeypatch.setattr(ops, "cgemm", emu_cgemm.cgemm)
    return ops


def cl(x):          # (n, c, h, w) -> channels-last (n, h, w, c), contiguous
    return x.permute(0, 2, 3, 1).contiguous()


def close(a, b, tol=1e-5):
    a, b = a.detach(), b.detach()
    return float((a - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


@pytest.mark.parametrize("bf,h,w,cins,cout", [(2, 16, 8, [64], 64), (3, 10, 12, [24, 40], 32), (1, 5, 7, [3, 70, 9], 20)])
def test_conv3x3_forward_and_data_gradient(ops, bf, h, w, cins, cout):
    g = torch.Generator().manual_seed(bf * 100 + h)
    xs = [torch.randn(bf, c, h, w, generator=g) for c in cins]
    wt = torch.randn(cout, sum(cins), 3, 3, generator=g) * 0.1
    bias = torch.randn(cout, generator=g)
    x = torch.cat(xs, dim=1).requires_grad_(True)
    y_ref = F.conv2d(x, wt, bias, padding=1)
    out = torch.zeros(bf, h, w, cout)
    stats = torch.zeros(bf, cout // 4, 2, dtype=torch.float64)
    ops.conv3x3([cl(t) for t in xs], ops.pack_conv_taps(wt, cins, torch.float32), cout, out, bias=bias, gn_stats=stats, gn_group=4,
                frames_per_sample=1)
    assert close(out, cl(y_ref))
    # GroupNorm sums of the epilogue: per (frame-sample, group of 4 channels)
    yr = cl(y_ref).reshape(bf, h * w, cout // 4, 4).double()
    assert torch.allclose(stats[..., 0], yr.sum(dim=(1, 3)), rtol=1e-5, atol=1e-3) and torch.allclose(stats[..., 1], (yr * yr).sum(dim=(1, 3)), rtol=1e-5)
    # data gradient = the same front-end on dy with the flipped / transposed pack, columns split over the concatenated sources
    dy = torch.randn(bf, cout, h, w, generator=g)
    y_ref.backward(dy)
    wd = ops.pack_conv_taps(wt.flip(2, 3).permute(1, 0, 2, 3), [cout], torch.float32)
    dx = torch.zeros(bf, h, w, sum(cins))
    ops.conv3x3([cl(dy)], wd, sum(cins), dx)
    assert close(dx, cl(x.grad))
    if len(cins) == 2 and cins[0] % 16 == 0:          # the fused concat gradient: columns >= nsplit go to a second tensor
        d0, d1 = torch.zeros(bf, h, w, cins[0]), torch.zeros(bf, h, w, cins[1])
        ops.conv3x3([cl(dy)], wd, sum(cins), d0, out2=d1, nsplit=cins[0])
        assert close(torch.cat((d0, d1), dim=-1), cl(x.grad))


def test_linear_rows_with_two_sources_and_residual(ops):
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn(50, 24, generator=g), torch.randn(50, 100, generator=g)
    wt = torch.randn(40, 124, generator=g)
    res = torch.randn(50, 40, generator=g)
    bias = torch.randn(40, generator=g)
    out = torch.zeros(50, 40)
    wp = ops.pack_conv_taps(wt[:, :, None, None], [24, 100], torch.float32)
    ops.linear_rows([a, b], wp, 40, out, bias=bias, res=res, alpha=0.5)
    assert close(out, 0.5 * (torch.cat((a, b), 1) @ wt.t()) + bias + res)
    out1 = torch.zeros(50, 40)
    ops.linear_rows([torch.cat((a, b), 1)], ops.pack_linear(wt, torch.float32), 40, out1)
    assert close(out1, torch.cat((a, b), 1) @ wt.t())


@pytest.mark.parametrize("bf,h,w,c", [(2, 8, 8, 16), (1, 12, 20, 70)])
def test_downsample_and_upsample_with_their_gradients(ops, bf, h, w, c):
    g = torch.Generator().manual_seed(h)
    x = torch.randn(bf, c, h, w, generator=g, requires_grad=True)
    wd_ = torch.randn(c, c, 4, 4, generator=g) * 0.1          # Conv3d weight [:, :, 0]
    wu_ = torch.randn(c, c, 1, 4, 4, generator=g) * 0.1       # ConvTranspose3d weight (cin, cout, 1, 4, 4)
    bias = torch.randn(c, generator=g)
    # Downsample forward; its data gradient is the Upsample front-end on the (cin <-> cout) pack
    y = F.conv2d(x, wd_, bias, stride=2, padding=1)
    out = torch.zeros(bf, h // 2, w // 2, c)
    ops.conv_down(cl(x.detach()), ops.pack_conv_taps(wd_, [c], torch.float32), c, out, bias=bias)
    assert close(out, cl(y))
    dy = torch.randn_like(y)
    y.backward(dy)
    dx = torch.zeros(bf, h, w, c)
    ops.conv_up(cl(dy), ops.pack_conv_taps(wd_.permute(1, 0, 2, 3), [c], torch.float32), c, dx)
    assert close(dx, cl(x.grad))
    # Upsample forward; its data gradient is the Downsample front-end
    x2 = torch.randn(bf, c, h, w, generator=g, requires_grad=True)
    y2 = F.conv_transpose2d(x2, wu_[:, :, 0], bias, stride=2, padding=1)
    out2 = torch.zeros(bf, 2 * h, 2 * w, c)
    ops.conv_up(cl(x2.detach()), ops.pack_conv_up(wu_, torch.float32), c, out2, bias=bias)
    assert close(out2, cl(y2))
    dy2 = torch.randn_like(y2)
    y2.backward(dy2)
    dx2 = torch.zeros(bf, h, w, c)
    ops.conv_down(cl(dy2), ops.pack_conv_taps(wu_[:, :, 0], [c], torch.float32), c, dx2)
    assert close(dx2, cl(x2.grad))


def test_init_conv_through_the_prepared_input_layout(ops):
    """vmm_prep_input lays the clip out as [bf][h][w + 6][8] (3-pixel zero border in x, channels padded to 8); the (1,7,7) conv is
    then 7 row taps of 8 px x 8 ch (K = 64 each, the 8th pixel has zero weights)."""
    g = torch.Generator().manual_seed(5)
    b, c, f, h, w, n = 2, 3, 2, 9, 11, 16
    x = torch.randn(b, c, f, h, w, generator=g)
    wt = torch.randn(n, c, 1, 7, 7, generator=g) * 0.1
    bias = torch.randn(n, generator=g)
    frames = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    y = F.conv2d(frames, wt[:, :, 0], bias, padding=3)
    xin = torch.zeros(b * f, h, w + 6, 8)
    xin[:, :, 3:3 + w, :c] = cl(frames)
    xin = torch.cat((xin.reshape(-1), torch.zeros(8)))
    out = torch.zeros(b * f, h, w, n)
    ops.init_conv(xin, b * f, h, w, ops.pack_init_conv(wt, torch.float32), n, out, bias=bias)
    assert close(out, cl(y))


def test_wrap_padded_view_gives_circular_convolution(ops):
    """What section 8f N3 needs from the kernel's contract: views are independent of the output grid, so a wrap-padded copy of the
    activation with non-negative tap offsets is a circular convolution (nn.Conv3d(padding_mode='circular'), VDDP:271), with the
    GroupNorm sums still taken over the unpadded output grid."""
    g = torch.Generator().manual_seed(8)
    bf, c, h, w, n = 2, 16, 6, 10, 8
    x = torch.randn(bf, c, h, w, generator=g)
    wt = torch.randn(n, c, 3, 3, generator=g) * 0.1
    y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="circular"), wt)
    xp = cl(F.pad(x, (1, 1, 1, 1), mode="circular"))
    taps, _ = ops.taps_conv(3, 3, [c], 0)                      # offsets 0..2 into the padded view
    out = torch.zeros(bf, h, w, n)
    stats = torch.zeros(bf, 1, 2, dtype=torch.float64)
    ops.cgemm([xp], [taps], ops.pack_conv_taps(wt, [c], torch.float32), n, out, (bf, h, w), gn_stats=stats, gn_group=n, frames_per_sample=1)
    assert close(out, cl(y))
    assert torch.allclose(stats[:, 0, 0], cl(y).double().sum(dim=(1, 2, 3)), rtol=1e-6, atol=1e-4)


# ------------------------------------------------------------------------------------------------
# weight gradients: the tap tables of ops.wgrad_* must scatter dY^T X into the master (torch) weight layout
# ------------------------------------------------------------------------------------------------
@pytest.fixture
def ops_w(monkeypatch):
    from videometamaterials_b200 import ops
    monkeypatch.setattr(ops, "cgemm", emu_cgemm.cgemm)
    monkeypatch.setattr(ops, "wgrad", emu_cgemm.wgrad)
    return ops


def test_weight_gradient_tap_tables(ops_w):
    ops = ops_w
    g = torch.Generator().manual_seed(12)
    bf, h, w = 2, 6, 10
    # 3x3 conv over two concatenated sources (VDDP:271 after the skip concat VDDP:813)
    cins, cout = [8, 24], 16
    xs = [torch.randn(bf, c, h, w, generator=g) for c in cins]
    wt = (torch.randn(cout, sum(cins), 1, 3, 3, generator=g) * 0.1).requires_grad_(True)
    dy = torch.randn(bf, cout, h, w, generator=g)
    F.conv2d(torch.cat(xs, 1), wt[:, :, 0], None, padding=1).backward(dy)
    dw = torch.full_like(wt.detach(), 0.5)                        # accumulated into, not overwritten
    ops.wgrad_conv3x3(cl(dy), [cl(t) for t in xs], dw)
    assert close(dw - 0.5, wt.grad, 1e-4)
    # linear layer, both operand orders of wgrad_linear (n > k puts the narrow operand on the row side)
    for n, k in ((40, 24), (16, 72)):
        x2, dy2 = torch.randn(90, k, generator=g), torch.randn(90, n, generator=g)
        dwl = torch.zeros(n, k)
        ops.wgrad_linear(dy2, [x2], dwl)
        assert close(dwl, dy2.t() @ x2, 1e-4)
    xa, xb, dy2 = torch.randn(90, 8, generator=g), torch.randn(90, 24, generator=g), torch.randn(90, 16, generator=g)
    dwl = torch.zeros(16, 32, 1, 1, 1)                            # res_conv weight shape
    ops.wgrad_linear(dy2, [xa, xb], dwl)
    assert close(dwl.reshape(16, 32), dy2.t() @ torch.cat((xa, xb), 1), 1e-4)
    # Downsample (strided 4x4) and Upsample (transposed 4x4)
    c = 12
    x = torch.randn(bf, c, h, w, generator=g)
    wd_ = (torch.randn(c, c, 1, 4, 4, generator=g) * 0.1).requires_grad_(True)
    yd = F.conv2d(x, wd_[:, :, 0], None, stride=2, padding=1)
    dyd = torch.randn_like(yd)
    yd.backward(dyd)
    dwd = torch.zeros_like(wd_.detach())
    ops.wgrad_down(cl(dyd), cl(x), dwd)
    assert close(dwd, wd_.grad, 1e-4)
    wu_ = (torch.randn(c, c, 1, 4, 4, generator=g) * 0.1).requires_grad_(True)
    yu = F.conv_transpose2d(x, wu_[:, :, 0], None, stride=2, padding=1)
    dyu = torch.randn_like(yu)
    yu.backward(dyu)
    dwu = torch.zeros_like(wu_.detach())
    ops.wgrad_up(cl(dyu), cl(x), dwu)
    assert close(dwu, wu_.grad, 1e-4)
    # init_conv (1,7,7) on the prepared 8-channel layout: only (kx < 7, ch < 3) columns of the 64-wide window land in dw
    ch, n = 3, 16
    xi = torch.randn(bf, ch, h, w, generator=g)
    wi = (torch.randn(n, ch, 1, 7, 7, generator=g) * 0.1).requires_grad_(True)
    yi = F.conv2d(xi, wi[:, :, 0], None, padding=3)
    dyi = torch.randn_like(yi)
    yi.backward(dyi)
    xin = torch.zeros(bf, h, w + 6, 8)
    xin[:, :, 3:3 + w, :ch] = cl(xi)
    xin = torch.cat((xin.reshape(-1), torch.zeros(8)))
    dwi = torch.zeros_like(wi.detach())
    ops.wgrad_init_conv(cl(dyi), xin, dwi, ch)
    assert close(dwi, wi.grad, 1e-4)


# ------------------------------------------------------------------------------------------------
# padding_mode = 'circular' / 'circular_1d' (model.yaml:13, VDDP:153-243): wrap-padded copies read with non-negative taps
# ------------------------------------------------------------------------------------------------
def _ref_pad(x, p, mode):          # (n, c, h, w), the reference's padding
    if mode == "circular":
        return F.pad(x, (p, p, p, p), mode="circular")
    return F.pad(F.pad(x, (p, p, 0, 0), mode="circular"), (0, 0, p, p))


@pytest.mark.parametrize("mode", ["circular", "circular_1d"])
def test_wrap_mode_front_ends_forward_data_gradient_weight_gradient(ops_w, mode):
    """Every convolution kind of the network in the two wrap modes: forward against the reference's construction (explicit pad +
    unpadded conv; for the upsampler pad 2 + transposed conv cropping 5, VDDP:164-216), the data gradient through the ADJOINT
    front-end (3x3 -> 3x3 with the flipped pack, down -> up, up -> down: the adjoint of a wrap-mode conv is a wrap-mode conv),
    and the weight gradient tap tables."""
    ops = ops_w
    g = torch.Generator().manual_seed(21)
    bf, h, w = 2, 8, 12
    # 3x3, two concatenated sources
    cins, cout = [8, 24], 16
    xs = [torch.randn(bf, c, h, w, generator=g) for c in cins]
    wt = (torch.randn(cout, sum(cins), 1, 3, 3, generator=g) * 0.1).requires_grad_(True)
    x = torch.cat(xs, 1).requires_grad_(True)
    y = F.conv2d(_ref_pad(x, 1, mode), wt[:, :, 0])
    out = torch.zeros(bf, h, w, cout)
    ops.conv3x3([cl(t) for t in xs], ops.pack_conv_taps(wt.detach()[:, :, 0], cins, torch.float32), cout, out, mode=mode)
    assert close(out, cl(y))
    dy = torch.randn_like(y)
    y.backward(dy)
    dx = torch.zeros(bf, h, w, sum(cins))
    ops.conv3x3([cl(dy)], ops.pack_conv_taps(wt.detach()[:, :, 0].flip(2, 3).permute(1, 0, 2, 3), [cout], torch.float32), sum(cins), dx, mode=mode)
    assert close(dx, cl(x.grad), 1e-4)
    dw = torch.zeros_like(wt.detach())
    ops.wgrad_conv3x3(cl(dy), [cl(t) for t in xs], dw, mode=mode)
    assert close(dw, wt.grad, 1e-4)
    # Downsample / Upsample
    c = 12
    x = torch.randn(bf, c, h, w, generator=g, requires_grad=True)
    wd_ = (torch.randn(c, c, 1, 4, 4, generator=g) * 0.1).requires_grad_(True)
    yd = F.conv2d(_ref_pad(x, 1, mode), wd_[:, :, 0], None, stride=2)
    out = torch.zeros(bf, h // 2, w // 2, c)
    ops.conv_down(cl(x.detach()), ops.pack_conv_taps(wd_.detach()[:, :, 0], [c], torch.float32), c, out, mode=mode)
    assert close(out, cl(yd))
    dyd = torch.randn_like(yd)
    yd.backward(dyd)
    dx = torch.zeros(bf, h, w, c)
    ops.conv_up(cl(dyd), ops.pack_conv_taps(wd_.detach()[:, :, 0].permute(1, 0, 2, 3), [c], torch.float32), c, dx, mode=mode)
    assert close(dx, cl(x.grad), 1e-4)
    dwd = torch.zeros_like(wd_.detach())
    ops.wgrad_down(cl(dyd), cl(x.detach()), dwd, mode=mode)
    assert close(dwd, wd_.grad, 1e-4)
    x2 = torch.randn(bf, c, h, w, generator=g, requires_grad=True)
    wu_ = (torch.randn(c, c, 1, 4, 4, generator=g) * 0.1).requires_grad_(True)
    yu = F.conv_transpose2d(_ref_pad(x2, 2, mode), wu_[:, :, 0], None, stride=2, padding=5)
    assert yu.shape[-2:] == (2 * h, 2 * w)
    out = torch.zeros(bf, 2 * h, 2 * w, c)
    ops.conv_up(cl(x2.detach()), ops.pack_conv_up(wu_.detach(), torch.float32), c, out, mode=mode)
    assert close(out, cl(yu))
    dyu = torch.randn_like(yu)
    yu.backward(dyu)
    dx2 = torch.zeros(bf, h, w, c)
    ops.conv_down(cl(dyu), ops.pack_conv_taps(wu_.detach()[:, :, 0], [c], torch.float32), c, dx2, mode=mode)
    assert close(dx2, cl(x2.grad), 1e-4)
    dwu = torch.zeros_like(wu_.detach())
    ops.wgrad_up(cl(dyu), cl(x2.detach()), dwu, mode=mode)
    assert close(dwu, wu_.grad, 1e-4)
    # init_conv (1,7,7) through the wrapped prepared layout
    ch, n = 3, 16
    xi = torch.randn(bf, ch, h, w, generator=g)
    wi = (torch.randn(n, ch, 1, 7, 7, generator=g) * 0.1).requires_grad_(True)
    yi = F.conv2d(_ref_pad(xi, 3, mode), wi[:, :, 0])
    xin = torch.zeros(bf, h, w + 6, 8)
    xin[:, :, 3:3 + w, :ch] = cl(xi)
    xin = ops.wrap_prepared_input(torch.cat((xin.reshape(-1), torch.zeros(8))), bf, h, w, mode)
    out = torch.zeros(bf, h, w, n)
    ops.init_conv(xin, bf, h, w, ops.pack_init_conv(wi.detach(), torch.float32), n, out, mode=mode)
    assert close(out, cl(yi))
    dyi = torch.randn_like(yi)
    yi.backward(dyi)
    dwi = torch.zeros_like(wi.detach())
    ops.wgrad_init_conv(cl(dyi), xin, dwi, ch, mode=mode)
    assert close(dwi, wi.grad, 1e-4)
