"""Python-side launch helpers over the C ABI (include/vmm.h).

Tensors are torch CUDA tensors used purely as device memory: every function turns them into raw
pointers + sizes and calls one `vmm_*` entry point on torch's current stream.  Activations are
channels-last `(b, f, h, w, c)` 16-bit.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import CgemmParams, FMT_BF16, FMT_F16, WgradParams, check, lib  # noqa: F401  (C, check, lib are used by blocks.py)


# bench.py sets PROFILE = [] to collect (kernel, algorithmic flops, start event, end event) per GEMM launch
PROFILE = None
PROFILE_TAGS = False        # probes: append the GEMM shape to the kernel name
HALO_MAX_N = int(os.environ.get("VMM_HALO_MAX_N", "128"))   # widest output of a 3x3 conv that takes the halo tile


def _prof_begin():
    if PROFILE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _prof_end(name: str, flops: float, e0, nbytes: float = 0.0) -> None:
    """(kernel, algorithmic FLOPs, start, end, algorithmic bytes = every operand read once + the output written once)."""
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    PROFILE.append((name, flops, e0, e1, nbytes))


def fmt_of(t: torch.Tensor) -> int:
    if t.dtype == torch.float16:
        return FMT_F16
    if t.dtype == torch.bfloat16:
        return FMT_BF16
    raise TypeError(f"expected a 16-bit activation tensor, got {t.dtype}")


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ceil64(c: int) -> int:
    return (c + 63) // 64 * 64


def _require_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("videometamaterials_b200 kernels need CUDA tensors (there is no CPU path)")


# ------------------------------------------------------------------------------------------------
# tile choice for the implicit GEMM
# ------------------------------------------------------------------------------------------------
def choose_tile(bf: int, oh: int, ow: int, max_tf: int = 8) -> Tuple[int, int, int]:
    """(tf, th, tw), powers of two with product 128, minimising zero-padded work; ties -> wider rows."""
    best = None
    for tw_l in range(8):
        tw = 1 << tw_l
        for th_l in range(8 - tw_l):
            th = 1 << th_l
            tf = 128 // (tw * th)
            if tf > max_tf:
                continue
            pad = (-(-bf // tf) * tf) * (-(-oh // th) * th) * (-(-ow // tw) * tw)
            key = (pad, -tw, -th)
            if best is None or key < best[0]:
                best = (key, (tf, th, tw))
    assert best is not None
    return best[1]


def _view_bfhwc(p: CgemmParams, idx: int, t: torch.Tensor) -> None:
    """Fill view idx from a (bf, h, w, c) tensor whose channel stride is 1."""
    assert t.dim() == 4 and t.stride(3) == 1, "view must be (bf, h, w, c) with contiguous channels"
    v = p.a[idx]
    v.ptr = t.data_ptr()
    bf, h, w, c = t.shape
    v.dims[0], v.dims[1], v.dims[2], v.dims[3] = c, w, h, bf
    sw, sh, sb = t.stride(2), t.stride(1), t.stride(0)
    # size-1 dims may carry arbitrary strides; give TMA something legal
    v.strides[0] = sw if w > 1 else max(8, c)
    v.strides[1] = sh if h > 1 else max(8, v.strides[0] * w)
    v.strides[2] = sb if bf > 1 else max(8, v.strides[1] * h)


def as_bfhwc(x: torch.Tensor) -> torch.Tensor:
    """(b, f, h, w, c) -> (b*f, h, w, c) view (no copy)."""
    b, f, h, w, c = x.shape
    return x.reshape(b * f, h, w, c) if x.is_contiguous() else x.view(b * f, h, w, c)


def cgemm(views: Sequence[torch.Tensor], taps: Sequence[Sequence[Tuple[int, int, int, int, int]]], w: torch.Tensor, n: int,
          out: torch.Tensor, grid: Tuple[int, int, int], *, out_geom: Optional[Tuple[int, int, int, int]] = None,
          phase_off: Optional[Sequence[Tuple[int, int]]] = None, bias: Optional[torch.Tensor] = None,
          res: Optional[torch.Tensor] = None, gn_stats: Optional[torch.Tensor] = None, gn_group: int = 0,
          frames_per_sample: int = 1, out2: Optional[torch.Tensor] = None, nsplit: int = 0,
          tile: Optional[Tuple[int, int, int]] = None, res2: Optional[torch.Tensor] = None, alpha: float = 1.0,
          rot: Optional[Tuple[torch.Tensor, int, int, int, int]] = None) -> None:
    """Launch vmm_cgemm.

    views : list of (bf, h, w, c) 16-bit tensors (A operand sources)
    taps  : per phase, list of (src, dy, dx, kofs, c)
    w     : packed weights [n_pad16, ktot] (same 16-bit dtype)
    out   : tensor whose last dim is the channel dim; rows are addressed by the pixel index formula
    grid  : (BF, OH, OW) of the GEMM-M space;  out_geom = (ohs, ows, sy, sx) (default: same grid, unit stride)
    """
    _require_cuda(w, out, *views)
    p = CgemmParams()
    p.fmt = fmt_of(views[0])
    p.n_views = len(views)
    for i, t in enumerate(views):
        assert t.dtype == views[0].dtype
        _view_bfhwc(p, i, t)
    p.n_phases = len(taps)
    for ph, tl in enumerate(taps):
        p.n_taps[ph] = len(tl)
        for i, (src, dy, dx, kofs, c) in enumerate(tl):
            T = p.taps[ph][i]
            T.src, T.dy, T.dx, T.kofs, T.c = src, dy, dx, kofs, c
        oy, ox = phase_off[ph] if phase_off is not None else (0, 0)
        p.phase_oy[ph], p.phase_ox[ph] = oy, ox
    assert w.dtype == views[0].dtype and w.is_contiguous() and w.dim() == 2
    p.w = w.data_ptr()
    p.n = n
    p.ktot = w.shape[1]
    assert w.shape[0] >= (n + 15) // 16 * 16
    bf, oh, ow = grid
    p.bf, p.oh, p.ow = bf, oh, ow
    p.tf, p.th, p.tw = tile if tile is not None else choose_tile(bf, oh, ow, max_tf=min(8, frames_per_sample) if gn_stats is not None else 8)
    p.out = out.data_ptr()
    assert out.stride(-1) == 1
    p.ldo = out.stride(-2)
    p.out_fp32 = 1 if out.dtype == torch.float32 else 0
    if not p.out_fp32:
        assert out.dtype == views[0].dtype
    ohs, ows, sy, sx = out_geom if out_geom is not None else (oh, ow, 1, 1)
    p.ohs, p.ows, p.sy, p.sx = ohs, ows, sy, sx
    if out2 is not None:
        p.out2 = out2.data_ptr()
        p.ldo2 = out2.stride(-2)
        p.nsplit = nsplit
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() >= n
        p.bias = bias.data_ptr()
    if res is not None:
        assert res.dtype == views[0].dtype and res.stride(-1) == 1
        p.res = res.data_ptr()
        p.ldr = res.stride(-2)
        if res2 is not None:
            assert res2.dtype == views[0].dtype and res2.stride(-1) == 1
            p.res2 = res2.data_ptr()
            p.ldr2 = res2.stride(-2)
    p.alpha = alpha
    if gn_stats is not None:
        assert gn_stats.dtype == torch.float64 and gn_stats.is_contiguous()
        p.gn_stats = gn_stats.data_ptr()
        p.gn_group = gn_group
        p.frames_per_sample = frames_per_sample
    if rot is not None:
        tab, p.rot_frames, p.rot_hw, p.rot_cols, p.rot_qcols = rot      # tab: fp32 (2, frames, 16, 2), see rotary_tables
        assert tab.dtype == torch.float32 and tab.is_contiguous() and tab.shape == (2, p.rot_frames, 16, 2)
        p.rot = tab.data_ptr()
    e0 = _prof_begin()
    check(lib.vmm_cgemm(C.byref(p), stream_ptr()), "vmm_cgemm")
    if e0 is not None:
        ksum = sum(c for tl in taps for (_, _, _, _, c) in tl)
        tag = f"|M={bf * oh * ow} N={n} K={ksum} taps={len(taps[0])}x{len(taps)} tile={p.tf}x{p.th}x{p.tw}" if PROFILE_TAGS else ""
        m = bf * oh * ow
        nbytes = sum(t.shape[0] * t.shape[1] * t.shape[2] * t.shape[3] * 2 for t in views) + w.numel() * 2 + m * n * out.element_size() * len(taps)
        if res is not None:
            nbytes += m * n * 2 * len(taps)
        _prof_end("cgemm" + tag, 2.0 * m * n * ksum, e0, nbytes)


# ------------------------------------------------------------------------------------------------
# weight packing (fp32 master weights -> 16-bit K-major GEMM operands)
# ------------------------------------------------------------------------------------------------
def _pad_rows16(w2d: torch.Tensor) -> torch.Tensor:
    n = w2d.shape[0]
    n_pad = (n + 15) // 16 * 16
    if n_pad != n:
        w2d = torch.cat((w2d, w2d.new_zeros(n_pad - n, w2d.shape[1])), dim=0)
    return w2d.contiguous()


def pack_linear(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """(N, K) -> [N_pad16, ceil64(K)] (one zero fill + one strided copy; the copy casts)."""
    n, k = w.shape
    out = torch.zeros((n + 15) // 16 * 16, ceil64(k), dtype=dtype, device=w.device)
    out[:n, :k] = w
    return out


def pack_conv_taps(w: torch.Tensor, splits: Sequence[int], dtype: torch.dtype) -> torch.Tensor:
    """(N, Cin, kh, kw) with Cin = sum(splits) -> [N_pad16, kh*kw*sum(ceil64(split))], tap-major then source-major
    (one zero fill + one strided copy per source; the copy casts)."""
    n, cin, kh, kw = w.shape
    assert sum(splits) == cin
    widths = [ceil64(c) for c in splits]
    per_tap = sum(widths)
    out = torch.zeros((n + 15) // 16 * 16, kh * kw * per_tap, dtype=dtype, device=w.device)
    o3 = out[:n].view(n, kh * kw, per_tap)
    c0 = k0 = 0
    for c, wd in zip(splits, widths):
        o3[:, :, k0:k0 + c] = w[:, c0:c0 + c].reshape(n, c, kh * kw).permute(0, 2, 1)
        c0 += c
        k0 += wd
    return out


def taps_conv(kh: int, kw: int, splits: Sequence[int], pad: int) -> Tuple[List[Tuple[int, int, int, int, int]], int]:
    taps, kofs = [], 0
    for ky in range(kh):
        for kx in range(kw):
            for s, c in enumerate(splits):
                taps.append((s, ky - pad, kx - pad, kofs, c))
                kofs += ceil64(c)
    return taps, kofs


# ------------------------------------------------------------------------------------------------
# convolution front-ends over cgemm (forward geometry; the same helpers serve the data gradients)
# ------------------------------------------------------------------------------------------------
def rows_view(x2d: torch.Tensor) -> torch.Tensor:
    """[M, C] (row stride may exceed C) -> (1, 1, M, C) view for the 'rows' GEMM."""
    return x2d.as_strided((1, 1, x2d.shape[0], x2d.shape[1]), (x2d.stride(0) * x2d.shape[0], x2d.stride(0) * x2d.shape[0], x2d.stride(0), 1))


def rotary_tables(rot: torch.Tensor, scale: float) -> torch.Tensor:
    """(frames, 16, 2) cos/sin -> (2, frames, 16, 2): table 0 pre-multiplied by the attention scale (queries), table 1 plain (keys)."""
    return torch.stack((rot * scale, rot), dim=0).contiguous()


def linear_rows(xs: Sequence[torch.Tensor], wp: torch.Tensor, n: int, out: torch.Tensor, *, bias=None, res=None,
                gn_stats=None, gn_group=0, frames_per_sample=1, out2=None, nsplit=0, res2=None, alpha=1.0, rot=None) -> None:
    """out[m, :n] = cat(xs, dim=1)[m] @ W^T  for 2-D operands [M, C_i]; wp = pack_linear / pack_conv_taps(1x1)."""
    taps, kofs = [], 0
    for s, x in enumerate(xs):
        taps.append((s, 0, 0, kofs, x.shape[1]))
        kofs += ceil64(x.shape[1])
    m = xs[0].shape[0]
    cgemm([rows_view(x) for x in xs], [taps], wp, n, out, (1, 1, m), bias=bias, res=res, gn_stats=gn_stats,
          gn_group=gn_group, frames_per_sample=frames_per_sample, out2=out2, nsplit=nsplit, res2=res2, alpha=alpha, rot=rot)


def wrap_pad(x: torch.Tensor, p: int, mode: str) -> torch.Tensor:
    """(bf, h, w, c) -> (bf, h + 2p, w + 2p, c) for the reference's non-zero padding modes (model.yaml:13): 'circular' wraps both
    pixel axes (nn.Conv3d(padding_mode='circular'), VDDP:240,271), 'circular_1d' wraps the horizontal axis and zero-fills the
    vertical one (Circular_1d_Conv3d, VDDP:219-236).  vmm_cgemm / vmm_wgrad read such a copy with non-negative tap offsets: their
    views are independent of the output grid, so the convolution (and its GroupNorm sums) still runs over the h x w grid."""
    xw = torch.cat((x[:, :, -p:], x, x[:, :, :p]), dim=2)
    if mode == "circular":
        return torch.cat((xw[:, -p:], xw, xw[:, :p]), dim=1).contiguous()
    if mode == "circular_1d":
        z = xw.new_zeros(xw.shape[0], p, xw.shape[2], xw.shape[3])
        return torch.cat((z, xw, z), dim=1).contiguous()
    raise ValueError(f"unknown padding mode {mode!r}")


def conv3x3(xs: Sequence[torch.Tensor], wp: torch.Tensor, n: int, out: torch.Tensor, mode: str = "zeros", **kw) -> None:
    """xs: list of (bf, h, w, c_i) sources (implicit channel concat); padding 1, 'zeros' or a wrap mode (see wrap_pad)."""
    bf, h, w, _ = xs[0].shape
    if mode != "zeros":
        taps, _ = taps_conv(3, 3, [x.shape[3] for x in xs], 0)
        cgemm([wrap_pad(x, 1, mode) for x in xs], [taps], wp, n, out, (bf, h, w), **kw)
        return
    taps, _ = taps_conv(3, 3, [x.shape[3] for x in xs], 1)
    if kw.get("tile") is None and h % 16 == 0 and w % 8 == 0 and n <= HALO_MAX_N:
        kw["tile"] = (1, 16, 8)      # vmm_cgemm switches to its halo mode on this tile (one A slab per kx instead of one tile per tap)
    cgemm(list(xs), [taps], wp, n, out, (bf, h, w), **kw)


def down_taps(c: int):
    taps = []
    for ky in range(4):
        for kx in range(4):
            py, px = (ky + 1) % 2, (kx + 1) % 2
            taps.append((py * 2 + px, (ky - 1) // 2, (kx - 1) // 2, (ky * 4 + kx) * ceil64(c), c))
    return taps


def parity_views(x: torch.Tensor):
    return [x[:, py::2, px::2, :] for py in range(2) for px in range(2)]


def down_taps_padded(c: int):
    """Taps of the strided 4x4 conv over the parity views of a copy padded by one pixel: padded row 2y + ky = parity ky % 2, row y + ky // 2."""
    return [((ky % 2) * 2 + kx % 2, ky // 2, kx // 2, (ky * 4 + kx) * ceil64(c), c) for ky in range(4) for kx in range(4)]


def conv_down(x: torch.Tensor, wp: torch.Tensor, n: int, out: torch.Tensor, mode: str = "zeros", **kw) -> None:
    """(1,4,4) stride (1,2,2) pad (0,1,1) conv, VDDP:238-243.  x (bf, h, w, c) -> out (bf, h/2, w/2, n)."""
    bf, h, w, c = x.shape
    if mode != "zeros":
        cgemm(parity_views(wrap_pad(x, 1, mode)), [down_taps_padded(c)], wp, n, out, (bf, h // 2, w // 2), **kw)
        return
    cgemm(parity_views(x), [down_taps(c)], wp, n, out, (bf, h // 2, w // 2), **kw)


def up_taps(c: int):
    phases, offs = [], []
    for py in range(2):
        for px in range(2):
            taps = []
            for ky, dy in (((1, 0), (3, -1)) if py == 0 else ((0, 1), (2, 0))):
                for kx, dx in (((1, 0), (3, -1)) if px == 0 else ((0, 1), (2, 0))):
                    taps.append((0, dy, dx, (ky * 4 + kx) * ceil64(c), c))
            phases.append(taps)
            offs.append((py, px))
    return phases, offs


def conv_up(x: torch.Tensor, wp: torch.Tensor, n: int, out: torch.Tensor, mode: str = "zeros", **kw) -> None:
    """ConvTranspose (1,4,4)/(1,2,2)/(0,1,1), VDDP:153-160, as 4 output phases of 2x2 taps.  out (bf, 2h, 2w, n).  The circular
    upsamplers of the reference (pad the input by 2, crop 5 output pixels per side, VDDP:164-216) are the same phase formula read
    from the wrapped input; a one-pixel pad covers the tap offsets -1 .. 1."""
    bf, h, w, c = x.shape
    phases, offs = up_taps(c)
    if mode != "zeros":
        phases = [[(src, dy + 1, dx + 1, kofs, cc) for (src, dy, dx, kofs, cc) in ph] for ph in phases]
        cgemm([wrap_pad(x, 1, mode)], phases, wp, n, out, (bf, h, w), out_geom=(2 * h, 2 * w, 2, 2), phase_off=offs, **kw)
        return
    cgemm([x], phases, wp, n, out, (bf, h, w), out_geom=(2 * h, 2 * w, 2, 2), phase_off=offs, **kw)


def pack_conv_up(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """ConvTranspose3d weight (Cin, Cout, 1, 4, 4) -> [Cout_pad, 16*ceil64(Cin)]."""
    return pack_conv_taps(w[:, :, 0].permute(1, 0, 2, 3), [w.shape[0]], dtype)


def pack_init_conv(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """(N, C<=8, 1, 7, 7) -> [N_pad, 7*64]: per ky a block of 8 px x 8 ch (px 7 and ch >= C are zero)."""
    n, c, _, kh, kw = w.shape
    assert kw <= 8 and c <= 8
    out = w.new_zeros(n, kh, 8, 8)
    out[:, :, :kw, :c] = w[:, :, 0].permute(0, 2, 3, 1)
    return _pad_rows16(out.reshape(n, kh * 64).to(dtype))


def wrap_prepared_input(xin: torch.Tensor, bf: int, h: int, w: int, mode: str) -> torch.Tensor:
    """The prepared network input [bf][h][w+6][8] (zero border in x, zero fill above / below by the TMA) re-laid as
    [bf][h+6][w+6][8] with the three-pixel border of a wrap mode filled in (+8 slack elements)."""
    inner = xin[: bf * h * (w + 6) * 8].view(bf, h, w + 6, 8)[:, :, 3:3 + w]
    return torch.cat((wrap_pad(inner, 3, mode).reshape(-1), xin.new_zeros(8)))


def init_conv(xin: torch.Tensor, bf: int, h: int, w: int, wp: torch.Tensor, n: int, out: torch.Tensor, ksize: int = 7,
              mode: str = "zeros", **kw) -> None:
    """xin: flat 16-bit buffer [bf][h][w+6][8] (+8 slack) from prep_input; (1,7,7) conv VDDP:624-628.  With a wrap mode xin is
    the [bf][h+6][w+6][8] buffer of wrap_prepared_input and the seven row taps start at the padded row y + ky."""
    assert ksize == 7
    if mode != "zeros":
        view = xin.as_strided((bf, h + 6, w, 64), ((h + 6) * (w + 6) * 8, (w + 6) * 8, 8, 1))
        cgemm([view], [[(0, ky, 0, ky * 64, 64) for ky in range(7)]], wp, n, out, (bf, h, w), **kw)
        return
    view = xin.as_strided((bf, h, w, 64), (h * (w + 6) * 8, (w + 6) * 8, 8, 1))
    taps = [(0, ky - 3, 0, ky * 64, 64) for ky in range(7)]
    cgemm([view], [taps], wp, n, out, (bf, h, w), **kw)


# ------------------------------------------------------------------------------------------------
# norms / attention cores / sampler wrappers
# ------------------------------------------------------------------------------------------------
def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def gn_silu_fwd(x, y, stats, gamma, beta, scale_shift, B, pix, C_, groups, act=True, eps=1e-5, res=None):
    e0 = _prof_begin()
    check(lib.vmm_gn_silu_fwd(_p(x), _p(res), _p(y), fmt_of(x), B, pix, C_, groups, _p(stats), _p(gamma), _p(beta), _p(scale_shift),
                              eps, 1 if act else 0, stream_ptr()), "vmm_gn_silu_fwd")
    _prof_end("gn_silu_fwd", 0.0, e0, 2.0 * B * pix * C_ * (3 if res is not None else 2))


def gn_silu_bwd(x, dy, dx, stats, gamma, beta, scale_shift, B, pix, C_, groups, dgamma, dbeta, dss, act=True, eps=1e-5, dx_colsum=None):
    nbytes = int(lib.vmm_gn_silu_bwd_workspace(B, C_, groups))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    e0 = _prof_begin()
    check(lib.vmm_gn_silu_bwd(_p(x), _p(dy), _p(dx), fmt_of(x), B, pix, C_, groups, _p(stats), _p(gamma), _p(beta),
                              _p(scale_shift), eps, 1 if act else 0, _p(dgamma), _p(dbeta), _p(dss), _p(dx_colsum), _p(ws), nbytes,
                              stream_ptr()), "vmm_gn_silu_bwd")
    _prof_end("gn_silu_bwd", 0.0, e0, 2.0 * B * pix * C_ * 5)        # two passes: (x, dy) read twice, dx written once


def ln_fwd(x2d, y2d, gamma, eps=1e-5):
    rows, C_ = x2d.shape
    e0 = _prof_begin()
    check(lib.vmm_ln_fwd(_p(x2d), _p(y2d), fmt_of(x2d), rows, C_, _p(gamma), eps, None, stream_ptr()), "vmm_ln_fwd")
    _prof_end("ln_fwd", 0.0, e0, 2.0 * rows * C_ * 2)


def ln_bwd(x2d, dy2d, dres2d, dx2d, gamma, dgamma, eps=1e-5):
    rows, C_ = x2d.shape
    e0 = _prof_begin()
    check(lib.vmm_ln_bwd(_p(x2d), _p(dy2d), _p(dres2d), _p(dx2d), fmt_of(x2d), rows, C_, _p(gamma), eps, _p(dgamma),
                         stream_ptr()), "vmm_ln_bwd")
    _prof_end("ln_bwd", 0.0, e0, 2.0 * rows * C_ * 4)


def tattn_fwd(qkv, ekv, bias, rot, out, B, frames, HW, heads, pre_rotated=False):
    e0 = _prof_begin()
    check(lib.vmm_tattn_fwd(_p(qkv), _p(ekv), _p(bias), _p(rot), _p(out), fmt_of(qkv), B, frames, HW, heads, 32 ** -0.5,
                            1 if pre_rotated else 0, stream_ptr()), "vmm_tattn_fwd")
    rows = B * frames * HW
    _prof_end("tattn_fwd", 2.0 * rows * heads * 32 * (2 * frames if ekv is not None else frames) * 2, e0, 2.0 * rows * heads * 32 * 4)


def ftattn_fwd(x, out, wqkv, wout, gamma, ekv, bias, rot, xn, qkv, ao, B, frames, HW, heads, eps=1e-5):
    """Fused Residual(PreNorm(temporal attention)) forward of a 64-channel level; xn / qkv / ao = None when nothing is kept."""
    C_ = x.shape[-1]
    nbytes_ws = int(lib.vmm_ftattn_workspace(B))
    ws = torch.empty(nbytes_ws, dtype=torch.uint8, device=x.device)
    e0 = _prof_begin()
    check(lib.vmm_ftattn_fwd(_p(x), _p(out), _p(wqkv), _p(wout), _p(gamma), _p(ekv), _p(bias), _p(rot), _p(xn), _p(qkv), _p(ao), _p(ws),
                             nbytes_ws, fmt_of(x), B, frames, HW, C_, heads, eps, stream_ptr()), "vmm_ftattn_fwd")
    rows, hd = B * frames * HW, heads * 32
    keys = 2 * frames if ekv is not None else frames
    flops = 2.0 * rows * (C_ * 3 * hd + hd * C_) + 2.0 * rows * hd * keys * 2
    nbytes = 2.0 * rows * (2 * C_ + (C_ + 4 * hd if qkv is not None else 0))
    _prof_end("ftattn_fwd", flops, e0, nbytes)


def flattn_fwd(x, out, wqkv, wout, gamma, bias_out, ekv, ctx, kstat, BF, frames, HW, heads, eps=1e-5):
    """Fused Residual(PreNorm(linear attention)) forward of a 64-channel level (inference form, nothing kept)."""
    C_ = x.shape[-1]
    nbytes_ws = int(lib.vmm_flattn_workspace(BF))
    ws = torch.empty(nbytes_ws, dtype=torch.uint8, device=x.device)
    e0 = _prof_begin()
    check(lib.vmm_flattn_fwd(_p(x), _p(out), _p(wqkv), _p(wout), _p(gamma), _p(bias_out), _p(ekv), ekv.shape[1], _p(ctx), _p(kstat), _p(ws),
                             nbytes_ws, fmt_of(x), BF, frames, HW, C_, heads, 32 ** -0.5, eps, stream_ptr()), "vmm_flattn_fwd")
    rows, hd = BF * HW, heads * 32
    _prof_end("flattn_fwd", 2.0 * rows * (C_ * 4 * hd + hd * C_) + 2.0 * rows * hd * 32 * 2, e0, 2.0 * rows * 3 * C_)


def lattn_fwd(qkv, ekv, T, out, ctx, kstat, BF, frames, HW, heads):
    e0 = _prof_begin()
    check(lib.vmm_lattn_fwd(_p(qkv), _p(ekv), T, _p(out), _p(ctx), _p(kstat), fmt_of(qkv), BF, frames, HW, heads, 32 ** -0.5,
                            stream_ptr()), "vmm_lattn_fwd")
    rows = BF * HW
    _prof_end("lattn_fwd", 2.0 * rows * heads * 32 * 32 * 2, e0, 2.0 * rows * heads * 32 * 4)


def sattn_fwd(qkv, ekv, out, lse, BF, frames, HW, heads):
    e0 = _prof_begin()
    check(lib.vmm_sattn_fwd(_p(qkv), _p(ekv), _p(out), _p(lse), fmt_of(qkv), BF, frames, HW, heads, 32 ** -0.5, stream_ptr()),
          "vmm_sattn_fwd")
    _prof_end("sattn_fwd", 2.0 * BF * HW * heads * 32 * (HW + 1) * 2, e0, 2.0 * BF * HW * heads * 32 * 4)


def prep_input(x, noise, a, c, s, xin, B, C_, F_, H, W):
    fmt = FMT_F16 if xin.dtype == torch.float16 else FMT_BF16
    check(lib.vmm_prep_input(_p(x), _p(noise), _p(a), _p(c), _p(s), _p(xin), fmt, B, C_, F_, H, W, stream_ptr()), "vmm_prep_input")


def loss_fwd_bwd(pred, target, loss_sum, dpred, B, C_, F_, H, W, l2=False, grad_scale=1.0):
    fmt = FMT_BF16 if (dpred is None or dpred.dtype == torch.bfloat16) else FMT_F16
    check(lib.vmm_loss(_p(pred), _p(target), _p(loss_sum), _p(dpred), fmt, B, C_, F_, H, W, 1 if l2 else 0, grad_scale,
                       stream_ptr()), "vmm_loss")


def cfg_x0(x, eps_cl, has_null, w, sr, srm1, x0, eps_out, B, C_, F_, H, W):
    check(lib.vmm_cfg_x0(_p(x), _p(eps_cl), 1 if has_null else 0, float(w), _p(sr), _p(srm1), _p(x0), _p(eps_out), B, C_, F_, H, W,
                         stream_ptr()), "vmm_cfg_x0")


_quantile_ws = {}


def abs_quantile(v, B, n, k, frac, floor_val, s_out):
    # per-(device, batch) workspace of the multi-CTA selection (4 x 256 counters per sample), created on first use and zeroed by the library
    # in stream order on every call; one buffer per key is enough because calls on a device are stream-ordered
    key = (v.device, int(B))
    ws = _quantile_ws.get(key)
    if ws is None:
        ws = torch.empty(max(int(lib.vmm_abs_quantile_workspace(B)), 16), dtype=torch.uint8, device=v.device)
        if not torch.cuda.is_current_stream_capturing():
            _quantile_ws[key] = ws          # (a buffer created during a graph capture belongs to that graph's pool: it is not kept for others)
    check(lib.vmm_abs_quantile(_p(v), B, n, k, frac, floor_val, _p(s_out), _p(ws), ws.numel(), stream_ptr()), "vmm_abs_quantile")


def posterior_step(x0, x, noise, s, c1, c2, sig, out, B, per):
    check(lib.vmm_posterior_step(_p(x0), _p(x), _p(noise), _p(s), _p(c1), _p(c2), _p(sig), _p(out), B, per, stream_ptr()),
          "vmm_posterior_step")


def axpby(a, b, ca, cb, cc, out):
    check(lib.vmm_axpby(_p(a), _p(b), float(ca), float(cb), float(cc), _p(out), out.numel(), stream_ptr()), "vmm_axpby")


def adam_ema_step(p, g, m, v, ema, lr, beta1, beta2, eps, step, grad_scale, ema_mode, ema_beta):
    check(lib.vmm_adam_ema_step(_p(p), _p(g), _p(m), _p(v), _p(ema), p.numel(), lr, beta1, beta2, eps, step, grad_scale, ema_mode,
                                ema_beta, stream_ptr()), "vmm_adam_ema_step")


# ------------------------------------------------------------------------------------------------
# weight gradients
# ------------------------------------------------------------------------------------------------
def _fill_view(v, t: torch.Tensor) -> None:
    assert t.dim() == 4 and t.stride(3) == 1
    v.ptr = t.data_ptr()
    bf, h, w, c = t.shape
    v.dims[0], v.dims[1], v.dims[2], v.dims[3] = c, w, h, bf
    sw, sh, sb = t.stride(2), t.stride(1), t.stride(0)
    v.strides[0] = sw if w > 1 else max(8, c)
    v.strides[1] = sh if h > 1 else max(8, v.strides[0] * w)
    v.strides[2] = sb if bf > 1 else max(8, v.strides[1] * h)


def wgrad(a_views: Sequence[torch.Tensor], b_views: Sequence[torch.Tensor], taps: Sequence[Tuple[int, int, int, int, int, int]],
          n: int, dw: torch.Tensor, s_m: int, s_c: int, grid: Tuple[int, int, int], *, s_c2: int = 0, cmod: int = 0,
          c_valid: int = 0, k_valid: int = 0, tile: Optional[Tuple[int, int, int]] = None) -> None:
    """dw (fp32, master layout, accumulated) += dY^T X.  taps: (a_src, b_src, dy, dx, c, wofs)."""
    _require_cuda(dw, *a_views, *b_views)
    assert dw.dtype == torch.float32
    p = WgradParams()
    p.fmt = fmt_of(a_views[0])
    p.n_a_views, p.n_b_views = len(a_views), len(b_views)
    for i, t in enumerate(a_views):
        _fill_view(p.a[i], t)
    for i, t in enumerate(b_views):
        assert t.dtype == a_views[0].dtype
        _fill_view(p.b[i], t)
    p.n_taps = len(taps)
    for i, (a_src, b_src, dy, dx, c, wofs) in enumerate(taps):
        T = p.taps[i]
        T.a_src, T.b_src, T.dy, T.dx, T.c, T.wofs = a_src, b_src, dy, dx, c, wofs
    p.n = n
    bf, oh, ow = grid
    p.bf, p.oh, p.ow = bf, oh, ow
    p.tf, p.th, p.tw = tile if tile is not None else choose_tile(bf, oh, ow)
    p.dw = dw.data_ptr()
    p.s_m, p.s_c, p.s_c2 = s_m, s_c, s_c2
    p.cmod, p.c_valid, p.k_valid = cmod, c_valid, k_valid
    e0 = _prof_begin()
    check(lib.vmm_wgrad(C.byref(p), stream_ptr()), "vmm_wgrad")
    if e0 is not None:
        tag = f"|px={bf * oh * ow} M={n} C={sum(tp[4] for tp in taps)} taps={len(taps)}" if PROFILE_TAGS else ""
        nbytes = sum(t.shape[0] * t.shape[1] * t.shape[2] * t.shape[3] * 2 for t in list(a_views) + list(b_views)) + n * sum(tp[4] for tp in taps) * 4
        _prof_end("wgrad" + tag, 2.0 * bf * oh * ow * n * sum(tp[4] for tp in taps), e0, nbytes)


def qkv_bwd(dqkv2d: torch.Tensor, xn2d: torch.Tensor, wd: torch.Tensor, dxn2d: torch.Tensor, dw: torch.Tensor) -> None:
    """dxn = dqkv @ W and dW += dqkv^T @ xn of a to_qkv projection with 64 input channels, reading dqkv once (csrc/qkvbwd.cu)."""
    _require_cuda(dqkv2d, xn2d, wd, dxn2d, dw)
    rows = dqkv2d.shape[0]
    assert dqkv2d.shape[1] == 768 and xn2d.shape == (rows, 64) and dxn2d.shape == (rows, 64) and wd.shape == (64, 768) and dw.numel() == 768 * 64
    assert dqkv2d.is_contiguous() and xn2d.is_contiguous() and dxn2d.is_contiguous() and wd.is_contiguous() and dw.is_contiguous() and dw.dtype == torch.float32
    e0 = _prof_begin()
    check(lib.vmm_qkv_bwd(_p(dqkv2d), _p(xn2d), _p(wd), _p(dxn2d), _p(dw), rows, fmt_of(dqkv2d), stream_ptr()), "vmm_qkv_bwd")
    _prof_end("qkv_bwd", 4.0 * rows * 768 * 64, e0, nbytes=rows * (768 + 64 + 64) * 2.0)


def qkv_ln_bwd(dqkv2d, xn2d, wd, x2d, dres2d, gamma, dx2d, dw, dgamma, eps=1e-5) -> None:
    """qkv_bwd with the LayerNorm backward + residual add of the block in its epilogue: dx = LN'(x)(dqkv @ W) + dres, dgamma += ..., dW += ..."""
    _require_cuda(dqkv2d, xn2d, wd, x2d, dres2d, dx2d, dw, dgamma)
    rows = dqkv2d.shape[0]
    assert dqkv2d.shape[1] == 768 and xn2d.shape == (rows, 64) and x2d.shape == (rows, 64) and dres2d.shape == (rows, 64) and dx2d.shape == (rows, 64)
    assert wd.shape == (64, 768) and dw.numel() == 768 * 64 and gamma.numel() == 64 and dgamma.numel() == 64
    for t in (dqkv2d, xn2d, wd, x2d, dres2d, dx2d, dw, gamma, dgamma):
        assert t.is_contiguous()
    assert dw.dtype == torch.float32 and gamma.dtype == torch.float32 and dgamma.dtype == torch.float32
    e0 = _prof_begin()
    check(lib.vmm_qkv_ln_bwd(_p(dqkv2d), _p(xn2d), _p(wd), _p(x2d), _p(dres2d), _p(gamma), eps, _p(dx2d), _p(dw), _p(dgamma), rows,
                             fmt_of(dqkv2d), stream_ptr()), "vmm_qkv_ln_bwd")
    _prof_end("qkv_bwd", 4.0 * rows * 768 * 64, e0, nbytes=rows * (768 + 4 * 64) * 2.0)


def qkv_bwd_eligible(dqkv2d: torch.Tensor, xn2d: torch.Tensor) -> bool:
    return FUSED_QKV_BWD and dqkv2d.shape[1] == 768 and xn2d.shape[1] == 64 and dqkv2d.shape[0] >= 128 * 148


FUSED_QKV_BWD = os.environ.get("VMM_FUSED_QKV_BWD", "1") != "0"
FUSED_QKV_LN_BWD = os.environ.get("VMM_FUSED_QKV_LN_BWD", "1") != "0"


def colsum(x2d: torch.Tensor, out: torch.Tensor) -> None:
    """out[n] (fp32, accumulated) += sum over rows of x2d[:, n]."""
    rows, n = x2d.shape
    check(lib.vmm_colsum(_p(x2d), rows, n, x2d.stride(0), fmt_of(x2d), _p(out), stream_ptr()), "vmm_colsum")


def wgrad_conv3x3(dy: torch.Tensor, xs: Sequence[torch.Tensor], dw: torch.Tensor, mode: str = "zeros") -> None:
    """dw: (Cout, Cin_total, 1, 3, 3) fp32.  dy (bf,h,w,Cout); xs: concat sources (bf,h,w,c_i)."""
    cout, cin_tot = dw.shape[0], dw.shape[1]
    taps, coff = [], 0
    for s, x in enumerate(xs):
        for ky in range(3):
            for kx in range(3):
                taps.append((0, s, ky - 1, kx - 1, x.shape[3], coff * 9 + ky * 3 + kx))
        coff += x.shape[3]
    bf, h, w, _ = dy.shape
    # 1 x 16 x 8 pixel tiles switch vmm_wgrad to its slab mode (one 18-row box of X per (kx, 64-channel chunk) serves the three ky taps)
    tile = (1, 16, 8) if (h % 16 == 0 and w % 8 == 0) else None
    if mode != "zeros":
        taps = [(a, b, ddy + 1, ddx + 1, c, wofs) for (a, b, ddy, ddx, c, wofs) in taps]
        wgrad([dy], [wrap_pad(x, 1, mode) for x in xs], taps, cout, dw, cin_tot * 9, 9, (bf, h, w), tile=tile)
        return
    wgrad([dy], list(xs), taps, cout, dw, cin_tot * 9, 9, (bf, h, w), tile=tile)


def wgrad_linear(dy2d: torch.Tensor, xs2d: Sequence[torch.Tensor], dw: torch.Tensor) -> None:
    """dw: (N, K_total) fp32 (any trailing singleton dims);  dy2d [M, N]; xs2d concat sources [M, K_i]."""
    n = dw.shape[0]
    ktot = dw.numel() // n
    m = dy2d.shape[0]
    if len(xs2d) == 1 and n > ktot:
        # dW^T = X^T dY: put the narrower operand on the 128-row M side and pack the wide one into accumulator columns
        # (3 chunks of 64 per block), which cuts the re-reads of the narrow operand (the kernel is L2-fill bound)
        wgrad([rows_view(xs2d[0])], [rows_view(dy2d)], [(0, 0, 0, 0, n, 0)], ktot, dw, 1, ktot, (1, 1, m))
        return
    taps, coff = [], 0
    for s, x in enumerate(xs2d):
        taps.append((0, s, 0, 0, x.shape[1], coff))
        coff += x.shape[1]
    wgrad([rows_view(dy2d)], [rows_view(x) for x in xs2d], taps, n, dw, ktot, 1, (1, 1, m))


def wgrad_down(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, mode: str = "zeros") -> None:
    """Strided (1,4,4) conv: dw (Cout, Cin, 1, 4, 4); dy (bf, h/2, w/2, Cout); x (bf, h, w, Cin)."""
    cout, cin = dw.shape[0], dw.shape[1]
    taps = []
    for ky in range(4):
        for kx in range(4):
            py, px = (ky + 1) % 2, (kx + 1) % 2
            taps.append((0, py * 2 + px, (ky - 1) // 2, (kx - 1) // 2, cin, ky * 4 + kx))
    bf, h2, w2, _ = dy.shape
    if mode != "zeros":
        taps = [(0, (ky % 2) * 2 + kx % 2, ky // 2, kx // 2, cin, ky * 4 + kx) for ky in range(4) for kx in range(4)]
        wgrad([dy], parity_views(wrap_pad(x, 1, mode)), taps, cout, dw, cin * 16, 16, (bf, h2, w2))
        return
    wgrad([dy], parity_views(x), taps, cout, dw, cin * 16, 16, (bf, h2, w2))


def wgrad_up(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, mode: str = "zeros") -> None:
    """Transposed conv: dw (Cin, Cout, 1, 4, 4); dy (bf, 2h, 2w, Cout); x (bf, h, w, Cin)."""
    cin, cout = dw.shape[0], dw.shape[1]
    taps = []
    for py in range(2):
        for px in range(2):
            for ky, ddy in (((1, 0), (3, -1)) if py == 0 else ((0, 1), (2, 0))):
                for kx, ddx in (((1, 0), (3, -1)) if px == 0 else ((0, 1), (2, 0))):
                    taps.append((py * 2 + px, 0, ddy, ddx, cin, ky * 4 + kx))
    bf, h, w, _ = x.shape
    if mode != "zeros":
        taps = [(a, b, ddy + 1, ddx + 1, c, wofs) for (a, b, ddy, ddx, c, wofs) in taps]
        wgrad(parity_views(dy), [wrap_pad(x, 1, mode)], taps, cout, dw, 16, cout * 16, (bf, h, w))
        return
    wgrad(parity_views(dy), [x], taps, cout, dw, 16, cout * 16, (bf, h, w))


def wgrad_init_conv(dy: torch.Tensor, xin: torch.Tensor, dw: torch.Tensor, channels: int, mode: str = "zeros") -> None:
    """init_conv: dw (N, C, 1, 7, 7); dy (bf, h, w, N); xin the padded 8-channel buffer of prep_input (of wrap_prepared_input
    with a wrap mode)."""
    n = dw.shape[0]
    bf, h, w, _ = dy.shape
    if mode != "zeros":
        view = xin.as_strided((bf, h + 6, w, 64), ((h + 6) * (w + 6) * 8, (w + 6) * 8, 8, 1))
        wgrad([dy], [view], [(0, 0, ky, 0, 64, ky * 7) for ky in range(7)], n, dw, channels * 49, 49, (bf, h, w), s_c2=1, cmod=8,
              c_valid=channels, k_valid=7)
        return
    view = xin.as_strided((bf, h, w, 64), (h * (w + 6) * 8, (w + 6) * 8, 8, 1))
    taps = [(0, 0, ky - 3, 0, 64, ky * 7) for ky in range(7)]
    # column j of the 64-wide window = (kx = j // 8, ch = j % 8) -> dw[n][ch][ky][kx]
    wgrad([dy], [view], taps, n, dw, channels * 49, 49, (bf, h, w), s_c2=1, cmod=8, c_valid=channels, k_valid=7)


def tattn_bwd(qkv, ekv, bias, rot, dout, dqkv, dekv, dbias, B, frames, HW, heads, pre_rotated=False):
    e0 = _prof_begin()
    check(lib.vmm_tattn_bwd(_p(qkv), _p(ekv), _p(bias), _p(rot), _p(dout), _p(dqkv), _p(dekv), _p(dbias), fmt_of(qkv), B, frames, HW,
                            heads, 32 ** -0.5, 1 if pre_rotated else 0, stream_ptr()), "vmm_tattn_bwd")
    rows = B * frames * HW
    _prof_end("tattn_bwd", 2.0 * rows * heads * 32 * (2 * frames if ekv is not None else frames) * 5, e0, 2.0 * rows * heads * 32 * 7)


def lattn_bwd(qkv, ekv, T, dout, ctx, kstat, dctx, dqkv, dekv, BF, frames, HW, heads):
    e0 = _prof_begin()
    check(lib.vmm_lattn_bwd(_p(qkv), _p(ekv), T, _p(dout), _p(ctx), _p(kstat), _p(dctx), _p(dqkv), _p(dekv), fmt_of(qkv), BF, frames,
                            HW, heads, 32 ** -0.5, 1.0 / HW, stream_ptr()), "vmm_lattn_bwd")
    rows = BF * HW
    _prof_end("lattn_bwd", 2.0 * rows * heads * 32 * 32 * 5, e0, 2.0 * rows * heads * 32 * 7)


def sattn_bwd(qkv, ekv, aout, dout, lse, dqkv, dekv, BF, HW, heads):
    e0 = _prof_begin()
    check(lib.vmm_sattn_bwd(_p(qkv), _p(ekv), _p(aout), _p(dout), _p(lse), _p(dqkv), _p(dekv), fmt_of(qkv), BF, HW, heads, 32 ** -0.5,
                            stream_ptr()), "vmm_sattn_bwd")
    _prof_end("sattn_bwd", 2.0 * BF * HW * heads * 32 * (HW + 1) * 5, e0, 2.0 * BF * HW * heads * 32 * 8)


def gather_cast(src: torch.Tensor, idx: torch.Tensor, dst: torch.Tensor) -> None:
    assert src.dtype == torch.float32 and idx.dtype == torch.int32 and idx.numel() == dst.numel()
    check(lib.vmm_gather_cast(_p(src), _p(idx), _p(dst), dst.numel(), fmt_of(dst), stream_ptr()), "vmm_gather_cast")
