"""Device microbenchmark of the GroupNorm / LayerNorm kernels at the four resolution levels of the b = 8 training step
(not a test: run by hand, `PYTHONPATH=. python tests/probe_norm.py`).  Buffers rotate over more than the L2 capacity."""
import json
import sys

import torch

from videometamaterials_b200 import ops


def timed(fn, n_rot, reps=3):
    """n_rot * reps launches captured in one CUDA graph (no host launch cost in the figure), us per launch."""
    for k in range(n_rot):
        fn(k)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for k in range(n_rot):
                fn(k)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * n_rot)


def main():
    dt = torch.bfloat16
    out = []
    for (hw, C) in ((9216, 64), (2304, 128), (576, 256), (144, 512)):
        B, pix = 8, 11 * hw
        nbytes = B * pix * C * 2
        n_rot = max(2, int(400e6 // (3 * nbytes)) + 1)
        xs = [torch.randn(B, pix, C, device="cuda", dtype=dt) for _ in range(n_rot)]
        dys = [torch.randn(B, pix, C, device="cuda", dtype=dt) for _ in range(n_rot)]
        ys = [torch.empty(B, pix, C, device="cuda", dtype=dt) for _ in range(n_rot)]
        stats = torch.zeros(B, 8, 2, device="cuda", dtype=torch.float64)
        xf = xs[0].double().view(B, pix, 8, C // 8)
        stats[:, :, 0] = xf.sum(dim=(1, 3))
        stats[:, :, 1] = (xf * xf).sum(dim=(1, 3))
        gamma = torch.ones(C, device="cuda")
        beta = torch.zeros(C, device="cuda")
        ss = torch.randn(B, 2 * C, device="cuda") * 0.1
        dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        dss = torch.zeros(B, 2 * C, device="cuda")
        cs = torch.zeros(C, device="cuda")
        t_f = timed(lambda k: ops.gn_silu_fwd(xs[k], ys[k], stats, gamma, beta, ss, B, pix, C, 8), n_rot)
        t_fr = timed(lambda k: ops.gn_silu_fwd(xs[k], ys[k], stats, gamma, beta, ss, B, pix, C, 8, res=dys[k]), n_rot)
        t_b = timed(lambda k: ops.gn_silu_bwd(xs[k], dys[k], ys[k], stats, gamma, beta, ss, B, pix, C, 8, dg, db, dss, dx_colsum=cs), n_rot)
        rows = B * pix
        t_lf = timed(lambda k: ops.ln_fwd(xs[k].view(rows, C), ys[k].view(rows, C), gamma), n_rot)
        t_lb = timed(lambda k: ops.ln_bwd(xs[k].view(rows, C), dys[k].view(rows, C), dys[k].view(rows, C), ys[k].view(rows, C), gamma, dg), n_rot)
        r = dict(hw=hw, C=C, mbytes=nbytes / 1e6, n_rot=n_rot,
                 gn_fwd_us=t_f, gn_fwd_gbs=2 * nbytes / t_f / 1e3, gn_fwd_res_us=t_fr, gn_fwd_res_gbs=3 * nbytes / t_fr / 1e3,
                 gn_bwd_us=t_b, gn_bwd_alg_gbs=3 * nbytes / t_b / 1e3, gn_bwd_traffic_gbs=5 * nbytes / t_b / 1e3,
                 ln_fwd_us=t_lf, ln_fwd_gbs=2 * nbytes / t_lf / 1e3, ln_bwd_us=t_lb, ln_bwd_gbs=4 * nbytes / t_lb / 1e3)
        print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)
        out.append(r)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
