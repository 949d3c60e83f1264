"""Level-0 timing of the fused linear-attention block (inference form) against the unfused kernel chain (b = 8, 96 x 96 x 11, C = 64)."""
import sys
import torch
from videometamaterials_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
Fr, H, W, Cc, heads, T = 11, 96, 96, 64, 8, 11
hd, n = heads * 32, H * W
rows = B * Fr * n
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
    return min(ts)


for dt in (torch.float16, torch.bfloat16):
    torch.manual_seed(0)
    x = torch.randn(B, Fr, H, W, Cc, device="cuda").to(dt)
    gamma = torch.ones(Cc, device="cuda")
    wqp = ops.pack_linear(torch.randn(3 * hd, Cc, device="cuda") * Cc ** -0.5, dt)
    wop = ops.pack_linear(torch.randn(Cc, hd, device="cuda") * hd ** -0.5, dt)
    bo = torch.zeros(Cc, device="cuda")
    ekv = torch.randn(B, T, 2 * hd, device="cuda")
    out = torch.empty_like(x)
    ctx = torch.empty(B * Fr, heads, 32, 32, device="cuda")
    kstat = torch.empty(B * Fr, heads, 32, 2, device="cuda")
    x2 = x.reshape(-1, Cc)
    xn = torch.empty_like(x2)
    qkv = torch.empty(rows, 3 * hd, device="cuda", dtype=dt)
    ao = torch.empty(rows, hd, device="cuda", dtype=dt)

    def unfused():
        ops.ln_fwd(x2, xn, gamma)
        ops.linear_rows([xn], wqp, 3 * hd, qkv)
        ops.lattn_fwd(qkv, ekv, T, ao, ctx, kstat, B * Fr, Fr, n, heads)
        ops.linear_rows([ao], wop, Cc, out.reshape(-1, Cc), bias=bo, res=x2)

    t_u = timed(unfused)
    t_f = timed(lambda: ops.flattn_fwd(x, out, wqp, wop, gamma, bo, ekv, ctx, None, B * Fr, Fr, n, heads))
    print(f"{dt}: unfused chain {t_u:.0f} us | fused (3 kernels) {t_f:.0f} us ({rows * Cc * 2 * 3 / t_f / 1e3:.0f} GB/s of 2 x read + write)")
