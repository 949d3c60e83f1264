"""Host rows next to the hot path, timed on this machine's CPU beside the UNMODIFIED reference
(imported from /root/reference through oracle/shims: build container only).  Writes profiles/r1_host_rows.md.

    python tools/bench_host_rows.py
"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
os.chdir(REF)

from denoising_diffusion_pytorch.video_denoising_diffusion_pytorch import Dataset as RefDataset  # noqa: E402
from denoising_diffusion_pytorch.video_denoising_diffusion_pytorch import video_tensor_to_gif as ref_gif  # noqa: E402
from src.utils import clean_pred as ref_clean_pred  # noqa: E402

from videometamaterials_b200.dataset import Dataset, clean_pred, video_tensor_to_gif, write_synthetic_dataset  # noqa: E402


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts)


def main():
    from scipy import ndimage
    rows = []
    torch.set_num_threads(1)
    # clean_pred on 48 x 48 quarter frames (what save_preds produces at 96 px), smooth random fields thresholded at 0.5
    rng = np.random.default_rng(0)
    g = ndimage.uniform_filter(rng.random((16, 48, 48)), size=(1, 5, 5))
    g = (g - g.min()) / (g.max() - g.min())
    a = ref_clean_pred(g.copy(), 48)
    b = clean_pred(g.copy(), 48)
    assert np.array_equal(a, b)
    t_ref = best(lambda: ref_clean_pred(g.copy(), 48)) / 16
    t_own = best(lambda: clean_pred(g.copy(), 48)) / 16
    rows.append(("`clean_pred`, one 48x48 geometry", t_ref * 1e3, t_own * 1e3, "identical output"))
    # Dataset item (five 96x96x11 GIFs of incompressible noise: worst case for the decoder)
    d = tempfile.mkdtemp() + "/"
    write_synthetic_dataset(d, 16, 96, 11)
    kw = dict(selected_channels=[0, 1, 3], num_frames=11, per_frame_cond=True, reference_frame='lagrangian')
    r = RefDataset(d, 96, **kw)
    o = Dataset(d, 96, **kw)
    o.decode_cache_bytes = 0
    c = Dataset(d, 96, **kw)
    t_pre = best(lambda: c.preload(), reps=1)
    for i in range(16):
        assert torch.equal(r[i][0], o[i][0]) and torch.equal(r[i][0], c[i][0])
    every = lambda ds: (lambda: [ds[i] for i in range(16)])
    t_r, t_o, t_c = best(every(r)) / 16, best(every(o)) / 16, best(every(c)) / 16
    rows.append(("`Dataset.__getitem__`, decode every time", t_r * 1e3, t_o * 1e3, "bit-equal items"))
    rows.append(("`Dataset.__getitem__`, decoded-frame cache", t_r * 1e3, t_c * 1e3, f"bit-equal items; preload of 80 GIFs took {t_pre * 1e3:.0f} ms on {os.cpu_count()} threads"))
    # tiled prediction GIF of save_preds: 3 x 2 tiles of 100 x 100, 11 frames, one channel
    v = torch.rand(1, 11, 300, 200)
    p1, p2 = tempfile.mktemp(suffix=".gif"), tempfile.mktemp(suffix=".gif")
    ref_gif(v, p1)
    video_tensor_to_gif(v, p2)
    assert open(p1, "rb").read() == open(p2, "rb").read()
    rows.append(("`video_tensor_to_gif`, 11 frames of 300x200", best(lambda: ref_gif(v, p1)) * 1e3, best(lambda: video_tensor_to_gif(v, p2)) * 1e3,
                 "identical file bytes"))
    out = ["# Host rows (SURVEY.md §8f N1 / N2) beside the reference's own code", "",
           f"`tools/bench_host_rows.py`, build container ({os.cpu_count()} vCPU), one torch thread, best of 3.  Both columns are CPU code;",
           "the reference is the unmodified source under /root/reference.  These rows are not on the GPU path: what matters is that",
           "they do not starve it (one B200 consumes ~210 clips/s in training and produces 4 sampled clips every ~3 s).", "",
           "| Row | reference ms | this repo ms | speed-up | parity |", "|---|---:|---:|---:|---|"]
    for name, tr, to, par in rows:
        out.append(f"| {name} | {tr:.2f} | {to:.2f} | {tr / to:.1f}x | {par} |")
    out += ["", f"Dataset through the cache: {1 / t_c:.0f} items/s per process (reference: {1 / t_r:.0f} items/s), i.e. one loader process feeds "
                f"{1 / t_c / 210:.1f} B200s at the measured 210 clips/s instead of {1 / t_r / 210:.2f}."]
    path = os.path.join(ROOT, "profiles", "r1_host_rows.md")
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
