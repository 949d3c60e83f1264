"""Level-0 timing of the fused to_qkv backward (vmm_qkv_bwd) against the two launches it replaces (b = 8, 96 x 96 x 11, 64 channels)."""
import torch
from videometamaterials_b200 import ops

rows = 8 * 11 * 96 * 96
dt = torch.bfloat16
torch.manual_seed(0)
dqkv = (torch.randn(rows, 768, device="cuda") * 0.5).to(dt)
xn = torch.randn(rows, 64, device="cuda").to(dt)
wd = ops.pack_linear((torch.randn(768, 64, device="cuda") * 0.1).t(), dt)
dxn = torch.empty(rows, 64, device="cuda", dtype=dt)
dw = torch.zeros(768, 64, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


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


def two():
    ops.linear_rows([dqkv], wd, 64, dxn)
    ops.wgrad_linear(dqkv, [xn], dw)


t2 = timed(two)
t1 = timed(lambda: ops.qkv_bwd(dqkv, xn, wd, dxn, dw))
nbytes = rows * (768 + 64 + 64) * 2
print(f"two launches {t2:.0f} us | fused {t1:.0f} us ({nbytes / t1 / 1e3:.0f} GB/s of dqkv + xn read + dxn written once)")
x = torch.randn(rows, 64, device="cuda").to(dt)
dres = torch.randn(rows, 64, device="cuda").to(dt)
gamma = torch.ones(64, device="cuda")
dx = torch.empty_like(x)
dg = torch.zeros(64, device="cuda")


def three():
    two()
    ops.ln_bwd(x, dxn, dres, dx, gamma, dg)


t3 = timed(three)
t4 = timed(lambda: ops.qkv_ln_bwd(dqkv, xn, wd, x, dres, gamma, dx, dw, dg))
print(f"with the LayerNorm backward: three launches {t3:.0f} us | fused {t4:.0f} us ({rows * (768 + 4 * 64) * 2 / t4 / 1e3:.0f} GB/s)")
