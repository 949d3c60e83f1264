"""Accuracy of the two rotary placements of the temporal attention against fp32 autograd on the unrounded projection."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vdm_oracle as O
from videometamaterials_b200 import ops

def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
B, Fr, HW, heads, C = 2, 11, 64, 8, 64
hd = heads * 32
for dt in (torch.float16, torch.bfloat16):
    torch.manual_seed(41)
    x = torch.randn(B * Fr * HW, C, device="cuda").to(dt)
    w = (torch.randn(3 * hd, C, device="cuda") / C ** 0.5).to(dt)
    wp = ops.pack_linear(w.float(), dt)
    freqs = 1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32)).cuda()
    ang = torch.arange(Fr, device="cuda").float()[:, None] * freqs[None, :]
    rot = torch.stack((ang.cos(), ang.sin()), -1).contiguous()
    tabs = ops.rotary_tables(rot, 32 ** -0.5)
    plain = torch.empty(B * Fr * HW, 3 * hd, device="cuda", dtype=dt); rotd = torch.empty_like(plain)
    ops.linear_rows([x], wp, 3 * hd, plain)
    ops.linear_rows([x], wp, 3 * hd, rotd, rot=(tabs, Fr, HW, 2 * hd, hd))
    ekv = torch.randn(B, 11, 2 * hd, device="cuda"); bias = torch.randn(heads, Fr, Fr, device="cuda")
    dout = torch.randn(B * Fr * HW, hd, device="cuda").to(dt)
    # fp32 reference on the unrounded projection
    y = (x.float() @ w.float().t()).requires_grad_(True)
    q, k, v = (t.reshape(B, Fr, HW, heads, 32).permute(0, 2, 3, 1, 4) for t in y.chunk(3, dim=-1))   # (B, HW, heads, Fr, 32)
    ek = ekv[..., :hd].reshape(B, 1, 11, heads, 32).transpose(2, 3).expand(B, HW, heads, 11, 32)
    ev = ekv[..., hd:].reshape(B, 1, 11, heads, 32).transpose(2, 3).expand(B, HW, heads, 11, 32)
    kk = torch.cat((ek, O.rotary(k, freqs)), -2); vv = torch.cat((ev, v), -2)
    sim = torch.einsum("...id,...jd->...ij", O.rotary(q * 32 ** -0.5, freqs), kk) + torch.cat((bias, bias), -1)
    out = torch.einsum("...ij,...jd->...id", sim.softmax(-1), vv).permute(0, 3, 1, 2, 4).reshape(B * Fr * HW, hd)
    out.backward(dout.float())
    res = {}
    for name, rows, pre in (("in-kernel rotary", plain, False), ("epilogue rotary", rotd, True)):
        o = torch.empty(B * Fr * HW, hd, device="cuda", dtype=dt); g = torch.empty_like(plain)
        de, db = torch.zeros_like(ekv), torch.zeros_like(bias)
        ops.tattn_fwd(rows, ekv, bias, rot, o, B, Fr, HW, heads, pre_rotated=pre)
        ops.tattn_bwd(rows, ekv, bias, rot, dout, g, de, db, B, Fr, HW, heads, pre_rotated=pre)
        print(f"{dt} {name:18s}: out {rel(o, out):.3e}  dq {rel(g[:, :hd], y.grad[:, :hd]):.3e}  dk {rel(g[:, hd:2*hd], y.grad[:, hd:2*hd]):.3e}  "
              f"dv {rel(g[:, 2*hd:], y.grad[:, 2*hd:]):.3e}  dbias {rel(db, torch.zeros_like(db) + db):.1e}")
