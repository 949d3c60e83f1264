"""Level-0 timing of the fused temporal-attention block against the unfused kernel chain (b = 8, 96 x 96 x 11, C = 64).
CUDA events, L2 flushed between repetitions.  Usage: PYTHONPATH=. python tests/probe_ftattn.py [B]"""
import sys

import torch

from videometamaterials_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
Fr, H, W, Cc, heads = 11, 96, 96, 64, 8
hd = heads * 32
rows = B * Fr * H * W
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts), sorted(ts)[len(ts) // 2]


for dt in (torch.bfloat16, torch.float16):
    torch.manual_seed(0)
    x = torch.randn(B, Fr, H, W, Cc, device="cuda").to(dt)
    gamma = torch.ones(Cc, device="cuda")
    wqp = ops.pack_linear(torch.randn(3 * hd, Cc, device="cuda") * Cc ** -0.5, dt)
    wop = ops.pack_linear(torch.randn(Cc, hd, device="cuda") * hd ** -0.5, dt)
    ekv = torch.randn(B, 11, 2 * hd, device="cuda")
    bias = torch.randn(heads, Fr, Fr, device="cuda")
    freqs = 1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32)).cuda()
    ang = torch.arange(Fr, device="cuda").float()[:, None] * freqs[None, :]
    rot = ops.rotary_tables(torch.stack((ang.cos(), ang.sin()), -1).contiguous(), 32 ** -0.5)
    out = torch.empty_like(x)
    xn = torch.empty(rows, Cc, device="cuda", dtype=dt)
    qkv = torch.empty(rows, 3 * hd, device="cuda", dtype=dt)
    ao = torch.empty(rows, hd, device="cuda", dtype=dt)
    x2 = x.reshape(-1, Cc)

    def unfused():
        ops.ln_fwd(x2, xn, gamma)
        ops.linear_rows([xn], wqp, 3 * hd, qkv, rot=(rot, Fr, H * W, 2 * hd, hd))
        ops.tattn_fwd(qkv, ekv, bias, rot[1], ao, B, Fr, H * W, heads, pre_rotated=True)
        ops.linear_rows([ao], wop, Cc, out.reshape(-1, Cc), res=x2)

    t_u = timed(unfused)
    t_f = timed(lambda: ops.ftattn_fwd(x, out, wqp, wop, gamma, ekv, bias, rot, None, None, None, B, Fr, H * W, heads))
    t_k = timed(lambda: ops.ftattn_fwd(x, out, wqp, wop, gamma, ekv, bias, rot, xn, qkv, ao, B, Fr, H * W, heads))
    t_n = timed(lambda: ops.ftattn_fwd(x, out, wqp, wop, gamma, None, bias, rot, None, None, None, B, Fr, H * W, heads))
    io = rows * Cc * 2 * 2
    print(f"{dt}: unfused chain {t_u[0]:.0f} us (median {t_u[1]:.0f}) | fused sampling {t_f[0]:.0f} us ({io / t_f[0] / 1e3:.0f} GB/s of x+out) | "
          f"fused keeping xn/qkv/ao {t_k[0]:.0f} us ({(io + rows * (Cc + 4 * hd) * 2) / t_k[0] / 1e3:.0f} GB/s) | fused without cond keys {t_n[0]:.0f} us")
