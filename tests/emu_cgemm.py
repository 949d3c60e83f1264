"""CPU statement of the `vmm_cgemm` CONTRACT (include/vmm.h) in plain torch  --  TEST INFRASTRUCTURE, never imported by the product.

    out[pix, n] = epilogue( sum_{tap} sum_{c} A_tap[pix + (dy, dx), c] * W[n, tap.kofs + c] )

with out-of-range reads of a view returning zero, output rows addressed by
`((bf * ohs) + y * sy + oy) * ows + x * sx + ox`, optional alpha / bias / residual, GroupNorm sums, and the column split into
`out2`.  It lets the CPU tests check the host side of the GEMM formulation -- tap tables, weight packs, parity views, output phases,
data-gradient packs -- against torch's own convolutions without a GPU.  It says nothing about the kernel; the `-m gpu` tests do.
"""
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


def _window(view: torch.Tensor, dy: int, dx: int, oh: int, ow: int) -> torch.Tensor:
    """view (bf, h, w, c) -> (bf, oh, ow, c) with element [y, x] = view[y + dy, x + dx], zero outside the view."""
    bf, h, w, c = view.shape
    top, left = max(0, -dy), max(0, -dx)
    bottom, right = max(0, dy + oh - h), max(0, dx + ow - w)
    v = F.pad(view.float(), (0, 0, left, right, top, bottom))
    return v[:, top + dy: top + dy + oh, left + dx: left + dx + ow, :]


def cgemm(views: Sequence[torch.Tensor], taps, w: torch.Tensor, n: int, out: torch.Tensor, grid: Tuple[int, int, int], *,
          out_geom: Optional[Tuple[int, int, int, int]] = None, phase_off=None, bias=None, res=None, gn_stats=None, gn_group: int = 0,
          frames_per_sample: int = 1, out2=None, nsplit: int = 0, tile=None, res2=None, alpha: float = 1.0, rot=None) -> None:
    bf, oh, ow = grid
    ohs, ows, sy, sx = out_geom if out_geom is not None else (oh, ow, 1, 1)
    W = w.float()
    yy = torch.arange(oh)[:, None]
    xx = torch.arange(ow)[None, :]
    for ph, tl in enumerate(taps):
        acc = torch.zeros(bf, oh, ow, n)
        for (src, dy, dx, kofs, c) in tl:
            a = _window(views[src], dy, dx, oh, ow)[..., :c]
            acc += a @ W[:n, kofs:kofs + c].t()
        acc = acc * (alpha if alpha != 0 else 1.0)
        if bias is not None:
            acc = acc + bias[:n].float()
        oy, ox = phase_off[ph] if phase_off is not None else (0, 0)
        rows = ((torch.arange(bf)[:, None, None] * ohs) + (yy * sy + oy)[None]) * ows + (xx * sx + ox)[None]       # (bf, oh, ow)
        rows = rows.reshape(-1)
        acc2 = acc.reshape(-1, n)
        if res is not None:
            r = res.reshape(-1, res.shape[-1]).float()
            if out2 is None:
                acc2 = acc2 + r[rows][:, :n]
            else:
                acc2 = torch.cat((acc2[:, :nsplit] + r[rows][:, :nsplit],
                                  acc2[:, nsplit:] + res2.reshape(-1, res2.shape[-1]).float()[rows][:, :n - nsplit]), dim=1)
        if rot is not None:
            # rotary epilogue (vmm_cgemm_params.rot): columns [0, rot_cols) are rotated in interleaved pairs of every 32-wide head
            # slice by the angle of the row's frame f = (row / rot_hw) % rot_frames; columns < rot_qcols use table 0 (pre-scaled)
            tab, rot_frames, rot_hw, rot_cols, rot_qcols = rot
            f = (rows // rot_hw) % rot_frames
            head = acc2[:, :rot_cols].reshape(-1, rot_cols // 32, 16, 2)
            which = (torch.arange(rot_cols // 32) * 32 >= rot_qcols).long()                    # per head slice: table 0 or 1
            cs = tab[which[None, :], f[:, None]]                                                # (rows, heads, 16, 2)
            xe, xo = head[..., 0], head[..., 1]
            turned = torch.stack((xe * cs[..., 0] - xo * cs[..., 1], xo * cs[..., 0] + xe * cs[..., 1]), dim=-1)
            acc2 = torch.cat((turned.reshape(-1, rot_cols), acc2[:, rot_cols:]), dim=1)
        if gn_stats is not None:
            g = acc2.reshape(bf // frames_per_sample, -1, n // gn_group, gn_group).double()
            gn_stats[..., 0] += g.sum(dim=(1, 3))
            gn_stats[..., 1] += (g * g).sum(dim=(1, 3))
        o = out.reshape(-1, out.shape[-1])
        if out2 is None:
            o[rows, :n] = acc2.to(out.dtype)
        else:
            o[rows, :nsplit] = acc2[:, :nsplit].to(out.dtype)
            out2.reshape(-1, out2.shape[-1])[rows, :n - nsplit] = acc2[:, nsplit:].to(out2.dtype)


def wgrad(a_views, b_views, taps, n: int, dw: torch.Tensor, s_m: int, s_c: int, grid: Tuple[int, int, int], *, s_c2: int = 0, cmod: int = 0,
          c_valid: int = 0, k_valid: int = 0, tile=None) -> None:
    """The `vmm_wgrad` contract: for every tap (a_src, b_src, dy, dx, c, wofs)
        dw.flat[wofs + m * s_m + col(j)] += sum_pix A[pix, m] * B[pix + (dy, dx), j],   m < n, j < c
    over the pixel grid of the a views, reads of B outside its view being zero; col(j) = j * s_c, or with cmod > 0
    (j % cmod) * s_c + (j // cmod) * s_c2 restricted to j % cmod < c_valid and j // cmod < k_valid."""
    bf, oh, ow = grid
    flat = dw.reshape(-1)
    for (a_src, b_src, dy, dx, c, wofs) in taps:
        a = a_views[a_src].float().reshape(-1, a_views[a_src].shape[-1])[:, :n] if a_views[a_src].shape[:3] == (bf, oh, ow) \
            else _window(a_views[a_src], 0, 0, oh, ow).reshape(-1, a_views[a_src].shape[-1])[:, :n]
        b = _window(b_views[b_src], dy, dx, oh, ow)[..., :c].reshape(-1, c)
        g = a.t() @ b                                              # (n, c)
        j = torch.arange(c)
        if cmod > 0:
            col = (j % cmod) * s_c + (j // cmod) * s_c2
            ok = ((j % cmod) < c_valid) & ((j // cmod) < k_valid)
        else:
            col, ok = j * s_c, torch.ones(c, dtype=torch.bool)
        idx = wofs + torch.arange(n)[:, None] * s_m + col[None, :]
        flat.index_put_((idx[:, ok].reshape(-1),), g[:, ok].reshape(-1), accumulate=True)
