"""Microbench of the non-GEMM kernels at the level-0 shapes of the b=8 training step (CUDA events, 10 reps after 3 warm-ups).
Prints achieved GB/s against the algorithmic bytes of each kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videometamaterials_b200 import ops

dt = torch.bfloat16
B, Fr, H, W, heads = 8, 11, 96, 96, 8
if len(sys.argv) > 1:
    H = W = int(sys.argv[1])
hd = heads * 32
HW = H * W
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def bench(name, fn, nbytes, reps=10):
    if only and not any(o in name for o in only):
        return
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    print(f"{name:28s} {ms*1e3:9.1f} us   {nbytes/ms/1e6:8.0f} GB/s  ({nbytes/1e6:.0f} MB)", flush=True)


torch.manual_seed(0)
qkv = torch.randn(B, Fr, H, W, 3 * hd, device="cuda").to(dt)
ekv = torch.randn(B, 11, 2 * hd, device="cuda")
dout = torch.randn(B, Fr, H, W, hd, device="cuda").to(dt)
out = torch.empty(B, Fr, H, W, hd, device="cuda", dtype=dt)
dqkv = torch.empty_like(qkv)
bias = torch.randn(heads, Fr, Fr, device="cuda")
freqs = 1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32)).cuda()
ang = torch.arange(Fr, device="cuda").float()[:, None] * freqs[None, :]
rot = torch.stack((ang.cos(), ang.sin()), -1).contiguous()
dekv = torch.zeros_like(ekv); dbias = torch.zeros_like(bias)
npos = B * Fr * HW
bench("tattn_fwd", lambda: ops.tattn_fwd(qkv, ekv, bias, rot, out, B, Fr, HW, heads), npos * (3 * hd + hd) * 2)
bench("tattn_bwd", lambda: ops.tattn_bwd(qkv, ekv, bias, rot, dout, dqkv, dekv, dbias, B, Fr, HW, heads), npos * (3 * hd + hd + 3 * hd) * 2)
bench("tattn_bwd_prerot", lambda: ops.tattn_bwd(qkv, ekv, bias, rot, dout, dqkv, dekv, dbias, B, Fr, HW, heads, pre_rotated=True), npos * (3 * hd + hd + 3 * hd) * 2)
ctx = torch.empty(B * Fr, heads, 32, 32, device="cuda"); kstat = torch.empty(B * Fr, heads, 32, 2, device="cuda"); dctx = torch.empty_like(ctx)
bench("lattn_fwd", lambda: ops.lattn_fwd(qkv, ekv, 11, out, ctx, kstat, B * Fr, Fr, HW, heads), npos * (2 * hd + 2 * hd + hd + hd) * 2)
bench("lattn_bwd", lambda: ops.lattn_bwd(qkv, ekv, 11, dout, ctx, kstat, dctx, dqkv, dekv, B * Fr, Fr, HW, heads), npos * (3 * hd + hd + 3 * hd + hd) * 2)
# GroupNorm / LayerNorm at C=64
C = 64
x = torch.randn(B, Fr * HW, C, device="cuda").to(dt); dy = torch.randn_like(x); y = torch.empty_like(x); dx = torch.empty_like(x)
stats = torch.stack(((x.float().reshape(B, -1, 8, 8).sum((1, 3))), (x.float().reshape(B, -1, 8, 8) ** 2).sum((1, 3))), -1).double().contiguous()
gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda"); ss = torch.randn(B, 2 * C, device="cuda") * 0.1
dgamma = torch.zeros(C, device="cuda"); dbeta = torch.zeros(C, device="cuda"); dss = torch.zeros_like(ss)
bench("gn_silu_fwd", lambda: ops.gn_silu_fwd(x, y, stats, gamma, beta, ss, B, Fr * HW, C, 8), npos * C * 4)
bench("gn_silu_fwd+res", lambda: ops.gn_silu_fwd(x, y, stats, gamma, beta, ss, B, Fr * HW, C, 8, res=dy), npos * C * 6)
bench("gn_silu_bwd", lambda: ops.gn_silu_bwd(x, dy, dx, stats, gamma, beta, ss, B, Fr * HW, C, 8, dgamma, dbeta, dss), npos * C * 10)
x2 = x.reshape(-1, C); dy2 = dy.reshape(-1, C); y2 = y.reshape(-1, C); dx2 = dx.reshape(-1, C)
bench("ln_fwd", lambda: ops.ln_fwd(x2, y2, gamma), npos * C * 4)
bench("ln_bwd", lambda: ops.ln_bwd(x2, dy2, y2, dx2, gamma, dgamma), npos * C * 8)
co = torch.zeros(C, device="cuda")
bench("colsum64", lambda: ops.colsum(x2, co), npos * C * 2)
q2 = qkv.reshape(-1, 3 * hd); co2 = torch.zeros(3 * hd, device="cuda")
bench("colsum768", lambda: ops.colsum(q2, co2), npos * 3 * hd * 2)
