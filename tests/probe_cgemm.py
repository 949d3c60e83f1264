"""GPU bring-up probe for vmm_cgemm: each case runs in its own process so that one trap / hang does not
take the others down.  Usage (on a GPU box):  python tests/probe_cgemm.py            (all cases)
                                               python tests/probe_cgemm.py CASE       (one case, in-process)
Not collected by pytest (no test_ prefix); the pytest parity tests live in tests/test_gpu_*.py.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ref_conv(x, w, b, pad):
    import torch.nn.functional as F
    # x (bf,h,w,c) 16-bit, w (n,c,kh,kw) 16-bit -> fp32 (bf,h,w,n)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=pad)
    return y.permute(0, 2, 3, 1).contiguous()


def run_case(name):
    import torch
    from videometamaterials_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"

    def report(tag, got, want, tol):
        err = float((got.float() - want).norm() / want.norm().clamp_min(1e-20))
        mx = float((got.float() - want).abs().max())
        ok = err < tol
        print(f"  {tag}: rel_l2={err:.3e} max_abs={mx:.3e} tol={tol:g} -> {'OK' if ok else 'FAIL'}")
        return ok

    def gemm_rows(M, K, N, dtype, bias=False, out_fp32=False):
        a = torch.randn(M, K, device=dev).to(dtype)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dtype)
        b = torch.randn(N, device=dev) if bias else None
        wp = ops.pack_linear(w.float(), dtype)
        out = torch.empty(M, N, device=dev, dtype=torch.float32 if out_fp32 else dtype)
        ops.cgemm([a.view(1, 1, M, K)], [[(0, 0, 0, 0, K)]], wp, N, out, (1, 1, M), bias=b)
        torch.cuda.synchronize()
        want = a.float() @ w.float().t() + (b if bias else 0)
        return report(f"gemm M={M} K={K} N={N} {dtype} fp32out={out_fp32}", out, want, 2e-3 if out_fp32 else 6e-3)

    def conv(bf, h, w_, cins, n, dtype, bias=True, res=False, gn=False, fps=11):
        xs = [torch.randn(bf, h, w_, c, device=dev).to(dtype) for c in cins]
        cin = sum(cins)
        wt = (torch.randn(n, cin, 3, 3, device=dev) / (9 * cin) ** 0.5).to(dtype)
        b = torch.randn(n, device=dev) if bias else None
        r = torch.randn(bf, h, w_, n, device=dev).to(dtype) if res else None
        wp = ops.pack_conv_taps(wt.float(), cins, dtype)
        taps, ktot = ops.taps_conv(3, 3, cins, 1)
        assert ktot == wp.shape[1]
        out = torch.empty(bf, h, w_, n, device=dev, dtype=dtype)
        nsamp = -(-bf // fps)
        groups = 8
        stats = torch.zeros(nsamp, groups, 2, device=dev, dtype=torch.float64) if gn else None
        ops.conv3x3(xs, wp, n, out, bias=b, res=r, gn_stats=stats, gn_group=n // groups, frames_per_sample=fps)
        torch.cuda.synchronize()
        want = ref_conv(torch.cat(xs, dim=-1), wt, b, 1)
        if res:
            want = want + r.float()
        ok = report(f"conv bf={bf} {h}x{w_} cin={cins} n={n} {dtype} res={res} gn={gn}", out, want, 6e-3)
        if gn:
            o = out.double()
            pad = nsamp * fps - bf
            if pad:
                o = torch.cat((o, o.new_zeros(pad, h, w_, n)), 0)
            o = o.view(nsamp, fps, h, w_, groups, n // groups)
            s1 = o.sum(dim=(1, 2, 3, 5))
            s2 = (o * o).sum(dim=(1, 2, 3, 5))
            want_s = torch.stack((s1, s2), dim=-1)
            ok &= report("   gn stats", stats, want_s.float().double(), 1e-4)
        return ok

    import torch
    bf16, f16 = torch.bfloat16, torch.float16
    cases = {
        "gemm_small": lambda: gemm_rows(256, 64, 64, bf16),
        "gemm_k256": lambda: gemm_rows(1000, 256, 128, bf16, bias=True),
        "gemm_n512": lambda: gemm_rows(4096, 1024, 512, bf16, bias=True),
        "gemm_n768_f16": lambda: gemm_rows(3000, 64, 768, f16),
        "gemm_fp32out_n3": lambda: gemm_rows(777, 64, 3, bf16, bias=True, out_fp32=True),
        "gemm_k16": lambda: gemm_rows(512, 16, 48, bf16),
        "conv_16": lambda: conv(2, 16, 16, [64], 64, bf16),
        "conv_96_gn": lambda: conv(11, 96, 96, [64], 64, bf16, gn=True),
        "conv_cat_res_gn": lambda: conv(22, 48, 48, [128, 64], 128, bf16, res=True, gn=True),
        "conv_24_f16": lambda: conv(11, 24, 24, [256], 256, f16, gn=True),
        "conv_12_512": lambda: conv(22, 12, 12, [512], 512, bf16, gn=True),
        "conv_smallc": lambda: conv(22, 16, 16, [16], 16, bf16, gn=True),
        "conv_smallc_cat": lambda: conv(11, 8, 8, [32, 16], 16, f16, res=True, gn=True),
        "conv_halo_cat": lambda: conv(13, 32, 24, [64, 64], 64, bf16, res=True, gn=True),
        "conv_halo_96_128": lambda: conv(5, 96, 96, [64], 128, f16, gn=True),
        "conv_halo_smallc": lambda: conv(11, 16, 16, [16, 8], 16, f16, res=True, gn=True),
    }
    if name.startswith("perfrows"):
        # perfrows:M:K:N  -- thin-K row GEMMs (QKV projection and friends)
        _, M, K, N = name.split(":")
        M, K, N = int(M), int(K), int(N)
        a = torch.randn(M, K, device=dev).to(bf16)
        wp = ops.pack_linear(torch.randn(N, K, device=dev) / K ** 0.5, bf16)
        out = torch.empty(M, N, device=dev, dtype=bf16)
        run = lambda: ops.cgemm([a.view(1, 1, M, K)], [[(0, 0, 0, 0, K)]], wp, N, out, (1, 1, M))
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        byts = 2.0 * M * (K + N)
        print(f"  perf rows M={M} K={K} N={N}: {ms * 1e3:.1f} us  {2.0 * M * K * N / ms / 1e9:.1f} TFLOP/s  {byts / ms / 1e6:.0f} GB/s")
        return True
    if name in ("perf", "perf1", "perf256"):
        shapes = [(88, 96, 64, 64), (88, 48, 128, 128), (88, 24, 256, 256), (88, 12, 512, 512), (88, 96, 128, 64), (88, 96, 64, 128),
                  (88, 48, 64, 128), (88, 48, 128, 64), (88, 48, 256, 128)]
        for (bf, h, cin, n) in (shapes[:1] if name == "perf1" else shapes[2:3] if name == "perf256" else shapes):
            x = torch.randn(bf, h, h, cin, device=dev).to(bf16)
            wt = (torch.randn(n, cin, 3, 3, device=dev) / (9 * cin) ** 0.5).to(bf16)
            wp = ops.pack_conv_taps(wt.float(), [cin], bf16)
            out = torch.empty(bf, h, h, n, device=dev, dtype=bf16)
            stats = torch.zeros(8, 8, 2, device=dev, dtype=torch.float64)
            bias = torch.zeros(n, device=dev)
            for _ in range(3):
                ops.conv3x3([x], wp, n, out, bias=bias, gn_stats=stats, gn_group=n // 8, frames_per_sample=11)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.conv3x3([x], wp, n, out, bias=bias, gn_stats=stats, gn_group=n // 8, frames_per_sample=11)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            fl = 2.0 * bf * h * h * n * 9 * cin
            print(f"  perf conv3x3 bf={bf} {h}x{h} {cin}->{n}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s")
            if name == "perf1":     # where does the tile time go: the same launch without the GroupNorm statistics / without bias and statistics
                for label, kw in (("no gn_stats", dict(bias=bias)), ("no bias, no gn_stats", dict())):
                    for _ in range(3):
                        ops.conv3x3([x], wp, n, out, **kw)
                    e0.record()
                    for _ in range(10):
                        ops.conv3x3([x], wp, n, out, **kw)
                    e1.record()
                    torch.cuda.synchronize()
                    ms2 = e0.elapsed_time(e1) / 10
                    print(f"    {label}: {ms2 * 1e3:.1f} us  {fl / ms2 / 1e9:.1f} TFLOP/s")
        return True
    return cases[name]()


ALL = ["gemm_small", "gemm_k256", "gemm_n512", "gemm_n768_f16", "gemm_fp32out_n3", "gemm_k16", "conv_16", "conv_96_gn",
       "conv_cat_res_gn", "conv_24_f16", "conv_12_512", "conv_smallc", "conv_smallc_cat", "conv_halo_cat", "conv_halo_96_128", "conv_halo_smallc", "perf"]

if __name__ == "__main__":
    if len(sys.argv) > 1:
        ok = run_case(sys.argv[1])
        sys.exit(0 if ok else 1)
    summary = []
    for c in ALL:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), c], capture_output=True, text=True, timeout=150)
            out = (r.stdout + r.stderr).strip().splitlines()
            status = "PASS" if r.returncode == 0 else f"FAIL(rc={r.returncode})"
        except subprocess.TimeoutExpired as e:
            out = ((e.stdout or b"").decode(errors="replace") + (e.stderr or b"").decode(errors="replace")).splitlines()
            status = "TIMEOUT"
        print(f"[{status}] {c} ({time.time() - t0:.1f}s)")
        for line in out[-12:]:
            print("   ", line)
        summary.append((c, status))
    print("SUMMARY", summary)
