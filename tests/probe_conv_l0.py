"""Device timing of one 3x3 convolution shape under the cgemm experiment switches (run by hand on a GPU box):
   PYTHONPATH=. [VMM_...=1] python tests/probe_conv_l0.py [cin] [n] [hw] [gn] [bias]"""
import sys

import torch

from videometamaterials_b200 import ops

cin = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hw = int(sys.argv[3]) if len(sys.argv) > 3 else 96
gn = int(sys.argv[4]) if len(sys.argv) > 4 else 1
bias = int(sys.argv[5]) if len(sys.argv) > 5 else 1
bf, dt, dev = 88, torch.bfloat16, "cuda"
nrot = 4
xs = [torch.randn(bf, hw, hw, cin, device=dev).to(dt) for _ in range(nrot)]
outs = [torch.empty(bf, hw, hw, n, device=dev, dtype=dt) for _ in range(nrot)]
wt = (torch.randn(n, cin, 3, 3, device=dev) / (9 * cin) ** 0.5)
wp = ops.pack_conv_taps(wt, [cin], dt)
b = torch.randn(n, device=dev) if bias else None
stats = torch.zeros(8, 8, 2, device=dev, dtype=torch.float64) if gn else None


def run(k):
    ops.conv3x3([xs[k]], wp, n, outs[k], bias=b, gn_stats=stats, gn_group=n // 8, frames_per_sample=11)


for k in range(nrot):
    run(k)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(5):
        for k in range(nrot):
            run(k)
g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (5 * nrot)
fl = 2.0 * bf * hw * hw * n * 9 * cin
print(f"conv cin={cin} n={n} hw={hw} gn={gn} bias={bias}: {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s")
