"""The reference's GIF dataset held DECODED in HBM ("GPU-side decode / normalise").

`DeviceDataset(ds)` takes a host `Dataset` (dataset.py: paths, per-sample ranges, global extrema, labels: all bit-equal to the
reference's `Dataset`, VDDP:1126-1397) and

  1. reads every GIF file an item needs into one host buffer, scans the containers natively (`vmm_gif_scan`, host C: extension /
     image-descriptor / sub-block walk) and uploads bytes + frame tables once;
  2. decodes on the device (`vmm_gif_decode`: LZW with one warp per frame, then per-file compositing + palette luma = what PIL's
     seek + convert('L') returns) into u8[sample][plane][frame][h][w], which stays resident: 96 x 96 x 11 x 5 planes = 507 KB per
     sample, so 180 GB of HBM hold ~350 000 samples;
  3. serves batches with ONE launch (`vmm_dataset_items`): gather by sample index + ToTensor + per-sample range + void pixels + global
     range, fp32 (b, c, f, h, w) straight into the layout `GaussianDiffusion.forward` takes, bit-identical to `Dataset.__getitem__`.

There is no host fallback: a GIF feature the device compositor does not reproduce (disposal method 3, a transparent first frame)
raises, naming the file; resize / crop / flip transforms are not applied on the device, so files whose logical screen differs from
`image_size` raise as well (the reference's data is written at the training resolution).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import ops
from .dataset import _LAYOUT, Dataset


def scan_gif(buf: bytes, max_frames: int = 64, pil_compat: bool = True) -> Tuple[Tuple[int, int], np.ndarray]:
    """Frame table of one GIF file (host; `vmm_gif_scan`): ((width, height), structured array of vmm_gif_frame rows).
    pil_compat: reproduce Pillow's decode of a palette that arrives while the image is still in mode 'L' (include/vmm.h), so that the
    planes equal what the reference's `gif_to_tensor` reads; False maps every frame through its own palette."""
    info = _lib.GifInfo()
    frames = (_lib.GifFrame * max_frames)()
    n = _lib.lib.vmm_gif_scan(buf, len(buf), 1 if pil_compat else 0, C.byref(info), frames, max_frames)
    if n < 0:
        raise _lib.VmmError(_lib.lib.vmm_last_error().decode())
    if n > max_frames:
        return scan_gif(buf, n, pil_compat)
    arr = np.frombuffer(frames, dtype=GIF_FRAME_DTYPE, count=n).copy()
    return (int(info.width), int(info.height)), arr


GIF_FRAME_DTYPE = np.dtype([("data_ofs", "<u4"), ("pal_ofs", "<u4"), ("px_ofs", "<u4"), ("x", "<u2"), ("y", "<u2"), ("w", "<u2"), ("h", "<u2"),
                            ("pal_size", "<u2"), ("min_code", "u1"), ("interlace", "u1"), ("disposal", "u1"), ("has_transp", "u1"),
                            ("transp", "u1"), ("background", "u1"), ("reserved", "<u4")])
assert GIF_FRAME_DTYPE.itemsize == C.sizeof(_lib.GifFrame) == 32


def decode_gifs(blobs: Sequence[bytes], frames_per_file: Optional[int], size_hw: Tuple[int, int], device, names: Optional[Sequence[str]] = None,
                chunk_bytes: int = 1 << 30, pil_compat: bool = True, timing: Optional[list] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Decode GIF files on the device: (u8 (n_files, frames_per_file, H, W) luminance, int32 (n_files,) frames each file holds, capped
    at frames_per_file; None = the largest frame count among the files).  Launched in chunks of about `chunk_bytes` of workspace."""
    H, W = size_hw
    label = (lambda i: names[i]) if names else (lambda i: f"file {i}")
    tables: List[np.ndarray] = []
    for i, blob in enumerate(blobs):
        try:
            (w, h), fr = scan_gif(blob, pil_compat=pil_compat)
        except _lib.VmmError as e:
            raise _lib.VmmError(f"{label(i)}: {e}") from None
        if (h, w) != (H, W):
            raise _lib.VmmError(f"{label(i)}: logical screen {w}x{h}, expected {W}x{H} (the device path applies no resize / crop)")
        tables.append(fr)
    if frames_per_file is None:
        frames_per_file = max(len(t) for t in tables)
    tables = [t[:frames_per_file] for t in tables]
    counts = torch.tensor([len(t) for t in tables], dtype=torch.int32)
    out = torch.empty((len(blobs), frames_per_file, H, W), dtype=torch.uint8, device=device)

    def launch(i0, i1):
        fr = np.concatenate(tables[i0:i1])
        begin = np.zeros(i1 - i0 + 1, dtype=np.int32)
        begin[1:] = np.cumsum([len(t) for t in tables[i0:i1]])
        px = fr["w"].astype(np.int64) * fr["h"].astype(np.int64)
        ofs = np.zeros(len(fr), dtype=np.int64)
        ofs[1:] = np.cumsum((px + 15) // 16 * 16)[:-1]          # 16-byte aligned index streams (vector stores in the decode kernel)
        ws_total = int(ofs[-1] + px[-1])
        if ws_total >= 1 << 32:
            raise _lib.VmmError("decode_gifs: more than 4 GiB of decode workspace in one chunk; lower chunk_bytes")
        fr["px_ofs"] = ofs.astype(np.uint32)
        fofs = np.zeros(i1 - i0 + 1, dtype=np.int64)
        fofs[1:] = np.cumsum([len(b) for b in blobs[i0:i1]])
        d_bytes = torch.from_numpy(np.concatenate([np.frombuffer(b, dtype=np.uint8) for b in blobs[i0:i1]])).to(device)      # one host copy
        d_fofs = torch.from_numpy(fofs).to(device)
        d_begin = torch.from_numpy(begin).to(device)
        d_fr = torch.from_numpy(fr.view(np.uint8).reshape(-1).copy()).to(device)
        d_ws = torch.empty(ws_total, dtype=torch.uint8, device=device)
        d_err = torch.zeros(1, dtype=torch.int32, device=device)
        if timing is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.check(_lib.lib.vmm_gif_decode(d_bytes.data_ptr(), d_fofs.data_ptr(), d_begin.data_ptr(), d_fr.data_ptr(), i1 - i0, len(fr),
                                           frames_per_file, H, W, int(px.max()), d_ws.data_ptr(), out[i0:i1].data_ptr(), d_err.data_ptr(), ops.stream_ptr()),
                   "vmm_gif_decode")
        if timing is not None:
            e1.record()
            torch.cuda.synchronize()
            timing.append(e0.elapsed_time(e1))          # ms of the two decode kernels of this chunk
        bad = int(d_err.item())         # the read-back also keeps this chunk's buffers alive until its kernels are done
        if bad:
            raise _lib.VmmError(f"vmm_gif_decode: {bad} frame(s) with a short or invalid LZW stream in {label(i0)} .. {label(i1 - 1)}")

    i0, ws_bytes, nbytes = 0, 0, 0
    for i, (blob, fr) in enumerate(zip(blobs, tables)):
        need = int((fr["w"].astype(np.int64) * fr["h"].astype(np.int64)).sum())
        if i > i0 and (ws_bytes + need > chunk_bytes or nbytes + len(blob) > chunk_bytes):
            launch(i0, i)
            i0, ws_bytes, nbytes = i, 0, 0
        ws_bytes += need
        nbytes += len(blob)
    launch(i0, len(blobs))
    return out, counts


def _read_file(name: str) -> bytes:
    with open(name, 'rb') as f:
        return f.read()


def decode_files(names: Sequence[str], frames_per_file: Optional[int], size_hw: Tuple[int, int], device, pil_compat: bool = True,
                 files_per_pass: int = 8192, decode=None, read=_read_file) -> Tuple[torch.Tensor, torch.Tensor]:
    """`decode_gifs` over files on disk.  Up to `files_per_pass` files are held in host memory at a time (a 96 x 96 x 11 noise GIF is
    ~115 KB, so the default keeps under 1 GB of file bytes resident; the real data set has tens of thousands of samples x 4-5 files):
    larger sets are read, decoded and released pass by pass into one output tensor.  Without a frame count (`frames_per_file=None`: the
    largest count among ALL files decides the layout) everything is read at once."""
    decode = decode if decode is not None else decode_gifs
    if frames_per_file is None or len(names) <= files_per_pass:
        return decode([read(nm) for nm in names], frames_per_file, size_hw, device, names, pil_compat=pil_compat)
    out, counts = None, torch.zeros(len(names), dtype=torch.int32)
    for i0 in range(0, len(names), files_per_pass):
        part = names[i0:i0 + files_per_pass]
        o, c = decode([read(nm) for nm in part], frames_per_file, size_hw, device, part, pil_compat=pil_compat)
        if out is None:
            out = torch.empty((len(names),) + tuple(o.shape[1:]), dtype=o.dtype, device=o.device)
        out[i0:i0 + len(part)].copy_(o)
        counts[i0:i0 + len(part)] = c
    return out, counts


def item_tables(ds: Dataset):
    """The tables `vmm_dataset_items` needs, from a host Dataset: (planes, ch_plane, ch_has_range, sample_rng (n, c, 2) fp32,
    global_rng (c, 2) fp32).  Host arithmetic only.  The reference's fp32 image tensor meets float64 0-dim range tensors (VDDP:1340-1358):
    torch computes such an operation in fp32 with the scalar converted to fp32, and the spans (max - min) are float64 operations whose
    RESULT is converted; that is what the fp32 table entries hold."""
    key = ds.reference_frame
    if key == 'lagrangian' and ds.num_frames == 1:
        key = 'lagrangian_1'
        ds.selected_channels = [0, 1]               # as Dataset.__getitem__ does (VDDP:1390)
    channels = _LAYOUT[key]['channels']
    selected = list(ds.selected_channels)
    planes = ['topo'] + sorted({channels[c][0] for c in selected} - {'topo'})
    r = ds.frame_ranges                              # float64 (n, 8 | 4)
    n, n_ch = r.shape[0], len(selected)
    srng = torch.zeros(n, n_ch, 2, dtype=torch.float32)
    grng = torch.zeros(n_ch, 2, dtype=torch.float32)
    ch_plane, ch_has = [], []
    for j, c in enumerate(selected):
        sub, lo, hi, glo, ghi = channels[c]
        ch_plane.append(planes.index(sub))
        ch_has.append(0 if hi is None else 1)
        if hi is None:
            continue
        smin = r[:, lo] if lo is not None else torch.zeros(n, dtype=torch.float64)
        srng[:, j, 0] = smin.to(torch.float32)
        srng[:, j, 1] = (r[:, hi] - smin).to(torch.float32)
        gmin = getattr(ds, glo) if glo is not None else torch.zeros((), dtype=torch.float64)
        grng[j, 0] = gmin.to(torch.float32)
        grng[j, 1] = (getattr(ds, ghi) - gmin).to(torch.float32)
    return planes, ch_plane, ch_has, srng, grng


class DeviceDataset:
    """Decoded dataset resident in HBM; `batch(indices)` is one kernel launch.  See the module docstring."""

    def __init__(self, ds: Dataset, device="cuda", max_frames: Optional[int] = None, pil_compat: bool = True, files_per_pass: int = 8192):
        if not isinstance(ds, Dataset):
            raise TypeError("DeviceDataset wraps a videometamaterials_b200.dataset.Dataset")
        if ds.horizontal_flip:
            raise NotImplementedError("DeviceDataset: horizontal_flip is a host-side PIL transform")
        self.ds = ds
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceDataset needs a CUDA device (there is no host implementation of the decode / item kernels)")
        planes, ch_plane, ch_has, srng, grng = item_tables(ds)
        self.selected = list(ds.selected_channels)
        self.planes = planes
        n = len(ds)
        S = ds.image_size
        # 1 + 2: file bytes -> device -> decoded planes
        names = [str(ds.paths[sub][i]) for i in range(n) for sub in planes]
        # an item uses at most num_frames frames of a file (cast_num_frames); without force_num_frames, all the frames the files hold
        fpf = int(max_frames) if max_frames is not None else (max(int(ds.num_frames), 1) if ds.force_num_frames else None)
        u8, counts = decode_files(names, fpf, (S, S), self.device, pil_compat=pil_compat, files_per_pass=files_per_pass)
        fpf = u8.shape[1]
        self.u8 = u8.view(n, len(planes), fpf, S, S)
        cnt = counts.view(n, len(planes))
        if not bool((cnt == cnt[:, :1]).all()):
            bad = int(torch.nonzero((cnt != cnt[:, :1]).any(dim=1))[0])
            raise _lib.VmmError(f"sample {bad}: its GIF files hold different numbers of frames {cnt[bad].tolist()} (the reference's torch.cat fails there)")
        self.sample_frames = cnt[:, 0].contiguous().to(self.device)
        self.frames = fpf
        n_ch = len(self.selected)
        self.sample_rng = srng.to(self.device)
        self.global_rng = grng.to(self.device)
        self.ch_plane = torch.tensor(ch_plane, dtype=torch.int32, device=self.device)
        self.ch_has_range = torch.tensor(ch_has, dtype=torch.int32, device=self.device)
        self.labels = ds.labels.to(self.device)
        self.frames_out = int(ds.num_frames) if ds.force_num_frames else None
        self.labels_scaling = ds.labels_scaling
        self.zero_u_2 = ds.zero_u_2

    def __len__(self) -> int:
        return self.u8.shape[0]

    def batch(self, index: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(x (b, c, f, h, w) fp32, labels (b, ...)) for the sample indices `index` (int64, host or device)."""
        index = torch.as_tensor(index, dtype=torch.int64).to(self.device).contiguous()
        b = int(index.numel())
        S = self.ds.image_size
        if self.frames_out is not None:
            f_out = self.frames_out
        else:                                    # force_num_frames=False: the frames the files hold (one count per batch, as torch.stack requires)
            cnt = self.sample_frames[index]
            f_out = int(cnt[0])
            if not bool((cnt == f_out).all()):
                raise RuntimeError("batch of samples with different frame counts and force_num_frames=False")
        n_ch = len(self.selected)
        out = torch.empty((b, n_ch, f_out, S, S), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.vmm_dataset_items(self.u8.data_ptr(), index.data_ptr(), b, len(self.planes), 0, n_ch, self.ch_plane.data_ptr(),
                                              self.ch_has_range.data_ptr(), self.sample_rng.data_ptr(), self.global_rng.data_ptr(),
                                              self.sample_frames.data_ptr(), self.frames, f_out, S * S, out.data_ptr(), ops.stream_ptr()),
                   "vmm_dataset_items")
        return out, self.labels[index]

    def __getitem__(self, i: int) -> Tuple[torch.Tensor, torch.Tensor]:
        x, lab = self.batch(torch.tensor([int(i)]))
        return x[0], lab[0]

    def loader(self, batch_size: int, shuffle: bool = True, drop_last: bool = False, generator: Optional[torch.Generator] = None,
               rank: int = 0, world_size: int = 1) -> "DeviceLoader":
        return DeviceLoader(self, batch_size, shuffle, drop_last, generator, rank, world_size)


class DeviceLoader:
    """Iterates batches like `DataLoader(ds, batch_size, shuffle)`: a fresh permutation per pass (`torch.randperm` from `generator`, as
    RandomSampler draws it), the last batch ragged unless drop_last.  With world_size > 1 every rank walks its own strided share of ONE
    permutation drawn from a common seed + pass number (what DistributedSampler does), padded by wrapping to equal length."""

    def __init__(self, dds: DeviceDataset, batch_size: int, shuffle: bool, drop_last: bool, generator, rank: int, world_size: int):
        self.dds, self.batch_size, self.shuffle, self.drop_last = dds, int(batch_size), shuffle, drop_last
        self.generator, self.rank, self.world_size = generator, int(rank), int(world_size)
        self.epoch = 0
        self.seed = 0

    def _order(self) -> torch.Tensor:
        n = len(self.dds)
        if self.world_size > 1:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            order = torch.randperm(n, generator=g) if self.shuffle else torch.arange(n)
            total = -(-n // self.world_size) * self.world_size
            if total > n:
                order = torch.cat((order, order[: total - n]))
            order = order[self.rank:total:self.world_size]
        elif self.shuffle:
            order = torch.randperm(n, generator=self.generator)
        else:
            order = torch.arange(n)
        self.epoch += 1
        return order

    def __len__(self) -> int:
        n = -(-len(self.dds) // self.world_size)
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        order = self._order()
        for i in range(len(self)):
            yield self.dds.batch(order[i * self.batch_size:(i + 1) * self.batch_size])
