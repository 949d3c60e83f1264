"""Dataset row measured on the GPU box: the host path (PIL decode in `Dataset`, what the reference's loader does
per item, plus the H2D copy of the batch) beside the device path (`DeviceDataset`: one-off GIF decode on the GPU, then one gather +
normalise launch per batch from the HBM-resident 8-bit planes).  Writes gpurun_out/device_dataset.json.

    python tools/bench_device_dataset.py [n_samples]
"""
import json
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from videometamaterials_b200.dataset import Dataset, write_synthetic_dataset  # noqa: E402
from videometamaterials_b200.device_dataset import DeviceDataset  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    d = tempfile.mkdtemp() + "/"
    write_synthetic_dataset(d, n, 96, 11)                   # five 96x96x11 GIFs of incompressible noise per sample: the decoder's worst case
    kw = dict(selected_channels=[0, 1, 3], num_frames=11, per_frame_cond=True, reference_frame='lagrangian')
    ds = Dataset(d, 96, **kw)
    ds.decode_cache_bytes = 0
    res = {"samples": n, "gif_files": n * 4, "frames": n * 4 * 11}
    # host path: PIL decode + normalise per item (one process, as the reference's DataLoader with num_workers=0, VDDP:1471)
    t = time.perf_counter()
    k = min(n, 64)
    items = [ds[i] for i in range(k)]
    res["host_item_ms_decode_every_time"] = (time.perf_counter() - t) / k * 1e3
    # device path: one-off decode
    torch.cuda.synchronize()
    t = time.perf_counter()
    dds = DeviceDataset(ds)
    torch.cuda.synchronize()
    res["device_build_s_total"] = time.perf_counter() - t          # file reads + native scan + upload + decode kernels
    # the decode kernels alone: run them again on the already scanned tables
    from videometamaterials_b200 import device_dataset as dd
    names = [str(ds.paths[sub][i]) for i in range(n) for sub in dds.planes]
    blobs = [open(p, 'rb').read() for p in names]
    res["gif_bytes"] = sum(len(b) for b in blobs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter()
    tables = [dd.scan_gif(b)[1] for b in blobs]
    res["host_scan_ms_per_file"] = (time.perf_counter() - t) / len(blobs) * 1e3
    dd.decode_gifs(blobs[:8], 11, (96, 96), 'cuda')              # warm
    torch.cuda.synchronize()
    t = time.perf_counter()
    e0.record()
    kt = []
    out, _ = dd.decode_gifs(blobs, 11, (96, 96), 'cuda', timing=kt)
    e1.record()
    torch.cuda.synchronize()
    res["decode_call_ms"] = (time.perf_counter() - t) * 1e3        # includes host scan + upload
    res["decode_frames_per_s"] = res["frames"] / (res["decode_call_ms"] * 1e-3)
    res["decode_kernels_ms"] = sum(kt)                              # LZW + compositing kernels alone (CUDA events)
    res["decode_kernels_frames_per_s"] = res["frames"] / (sum(kt) * 1e-3)
    res["decode_kernels_GBps_in_plus_out"] = (res["gif_bytes"] + res["frames"] * 96 * 96 * 2) / (sum(kt) * 1e-3) / 1e9
    assert torch.equal(out.view(dds.u8.shape), dds.u8)
    # batches
    idx = torch.randint(0, n, (8,), device='cuda')
    for _ in range(3):
        dds.batch(idx)
    torch.cuda.synchronize()
    e0.record()
    reps = 200
    for _ in range(reps):
        x, lab = dds.batch(idx)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res["device_batch8_us"] = ms * 1e3
    res["device_items_per_s"] = 8 / (ms * 1e-3)
    out_bytes = x.numel() * 4
    in_bytes = 8 * 4 * 11 * 96 * 96       # 3 channel planes + the topology plane
    res["device_batch_GBps"] = (out_bytes + in_bytes) / (ms * 1e-3) / 1e9
    for i in range(k):
        assert torch.equal(dds[i][0].cpu().view(torch.int32), items[i][0].view(torch.int32))
    res["bit_equal_items_checked"] = k
    res["hbm_bytes_per_sample"] = int(dds.u8[0].numel())
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "device_dataset.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
