"""Host-side data plumbing around the hot path.  CPU code, no kernels.

  * `Dataset`: the reference's on-disk layout (VDDP:1126-1397), Lagrangian reference frame, per-frame labels:
      <folder>/gifs/{topo,u_1,u_2,s_mises,s_22}/<i>.gif   8-bit 'L', num_frames frames
      <folder>/frame_range_data.csv   N x 8: min_u1,max_u1,min_u2,max_u2,max_mises,min_s22,max_s22,max_energy
      <folder>/stress_strain_data.csv N x K stress curve, interpolated to num_frames points
  * `write_synthetic_dataset`: writes random data in exactly that layout (the real dataset is an external download).
  * `SyntheticLagrangianDataset`: the same tensors generated in memory, for benches.
  * `Normalization`: the 'global-min-max-2' label scaling of src/normalization.py:35-37,79-80.
  * `clean_pred`, `video_tensor_to_gif`: post-processing used by Trainer.save_preds (src/utils.py:32-82, VDDP:1091-1098).
"""
from __future__ import annotations

import csv
import os
import threading
from pathlib import Path
from typing import Optional, Sequence

import numpy as np
import torch
from torch.utils import data


class Normalization:
    """Scale every label column to [-1, 1] with the global min / max of the training labels."""

    def __init__(self, labels: torch.Tensor, data_type=None, strategy: str = 'global-min-max-2'):
        if strategy != 'global-min-max-2':
            raise NotImplementedError(strategy)
        self.strategy = strategy
        self.globalmin = torch.min(labels)
        self.globalmax = torch.max(labels)

    def normalize(self, x: torch.Tensor) -> torch.Tensor:
        gmin, gmax = self.globalmin.to(x.device), self.globalmax.to(x.device)
        return 2. * (x - gmin) / (gmax - gmin) - 1.

    def unnormalize(self, x: torch.Tensor) -> torch.Tensor:
        gmin, gmax = self.globalmin.to(x.device), self.globalmax.to(x.device)
        return (0.5 * x + 0.5) * (gmax - gmin) + gmin


def fixed_normalization(gmin: float, gmax: float) -> Normalization:
    n = Normalization(torch.tensor([gmin, gmax]))
    return n


def _resize_short_edge(fr, size: int):
    """T.Resize(int) on a PIL image: short edge -> `size`, aspect ratio kept, bilinear."""
    from PIL import Image
    w, h = fr.size
    short, long = (w, h) if w <= h else (h, w)
    if short == size:
        return fr
    new_long = int(size * long / short)
    return fr.resize((size, new_long) if w <= h else (new_long, size), Image.BILINEAR)


def _center_crop(fr, size: int):
    """T.CenterCrop(int) on a PIL image (zero padding when the image is smaller, torchvision's rounding of the offsets)."""
    from PIL import Image
    w, h = fr.size
    if w < size or h < size:
        pl, pt = max((size - w) // 2, 0), max((size - h) // 2, 0)
        canvas = Image.new(fr.mode, (max(w, size), max(h, size)), 0)
        canvas.paste(fr, (pl, pt))
        fr, (w, h) = canvas, canvas.size
    if (w, h) == (size, size):
        return fr
    top, left = int(round((h - size) / 2.0)), int(round((w - size) / 2.0))
    return fr.crop((left, top, left + size, top + size))


def _gif_frames_u8(path, image_size: int, horizontal_flip: bool = False) -> torch.Tensor:
    """(frames, h, w) uint8 from an 8-bit GIF: every frame converted to 'L', then Resize -> (flip) -> CenterCrop as the
    reference's transform does (VDDP:1076-1106, 1252-1257); the final ToTensor (/255) is left to the caller."""
    from PIL import Image
    img = Image.open(path)
    frames, i = [], 0
    while True:
        try:
            img.seek(i)
        except EOFError:
            break
        fr = _resize_short_edge(img.convert('L'), image_size)
        if horizontal_flip and float(torch.rand(1)) < 0.5:
            fr = fr.transpose(Image.FLIP_LEFT_RIGHT)
        fr = _center_crop(fr, image_size)
        frames.append(torch.from_numpy(np.asarray(fr, dtype=np.uint8).copy()))
        i += 1
    return torch.stack(frames, dim=0)


def _gif_frames(path, image_size: int, horizontal_flip: bool = False) -> torch.Tensor:
    """(1, frames, h, w) float32 in [0, 1] (what the reference's gif_to_tensor returns for one channel)."""
    return _gif_frames_u8(path, image_size, horizontal_flip).to(torch.float32).div(255)[None]


def _cast_frames(t: torch.Tensor, frames: int) -> torch.Tensor:
    f = t.shape[1]
    if f == frames:
        return t
    if f > frames:
        return t[:, :frames]
    return torch.nn.functional.pad(t, (0, 0, 0, 0, 0, frames - f))


def _interp_curves(curves: np.ndarray, num_frames: int, strain: float = 0.2) -> np.ndarray:
    """Stress curves sampled at `num_frames` strains in [0, strain]; the first frame sits at 1 % of the final strain
    (VDDP:1260-1269, 1789-1799)."""
    given = np.linspace(0., strain, num=curves.shape[1])
    ev = np.linspace(0., strain, num=num_frames)
    ev[0] = 0.01 * strain
    return np.array([np.interp(ev, given, curves[i, :]) for i in range(curves.shape[0])])


# Per reference frame: the GIF sub-folders that must exist, and for every tensor channel
#   (sub-folder, column of frame_range_data.csv holding the per-sample min or None for 0, column of the per-sample max,
#    attribute with the global min or None for 0, attribute with the global max); a channel with no columns is the
#    topology mask itself and is passed through.
_LAYOUT = {
    'eulerian': dict(
        folders=('topo', 's_mises', 's_22', 'ener'),
        globals_=(('max_s_mises', 0, 'max'), ('min_s_22', 1, 'min'), ('max_s_22', 2, 'max'), ('max_strain_energy', 3, 'max')),
        channels=(('topo', None, None, None, None), ('s_mises', None, 0, None, 'max_s_mises'), ('s_22', 1, 2, 'min_s_22', 'max_s_22'),
                  ('ener', None, 3, None, 'max_strain_energy'))),
    'lagrangian': dict(
        folders=('topo', 'u_1', 'u_2', 's_mises', 's_22', 'ener'),
        globals_=(('min_u_1', 0, 'min'), ('max_u_1', 1, 'max'), ('min_u_2', 2, 'min'), ('max_u_2', 3, 'max'), ('max_s_mises', 4, 'max'),
                  ('min_s_22', 5, 'min'), ('max_s_22', 6, 'max'), ('max_strain_energy', 7, 'max')),
        channels=(('u_1', 0, 1, 'min_u_1', 'max_u_1'), ('u_2', 2, 3, 'min_u_2', 'max_u_2'), ('s_mises', None, 4, None, 'max_s_mises'),
                  ('s_22', 5, 6, 'min_s_22', 'max_s_22'))),
    # single-frame ablation of the Lagrangian data: topology + sigma_22 (VDDP:1365-1390)
    'lagrangian_1': dict(channels=(('topo', None, None, None, None), ('s_22', 5, 6, 'min_s_22', 'max_s_22'))),
}


class Dataset(data.Dataset):
    """The reference's GIF dataset (VDDP:1126-1397), same constructor arguments and the same tensors: fields are stored per
    sample as 8-bit GIFs scaled to that sample's own range (`frame_range_data.csv`); an item maps them back to physical
    values, zeroes the void pixels of the topology mask and rescales with the global extrema of the folder.  The arithmetic
    keeps the reference's types (float32 pixels, float64 0-dim range scalars) so that items are bit-identical."""

    # Decoded 8-bit frames are kept in host memory up to this many bytes (0 disables).  PIL needs ~3 ms per 96x96x11 GIF and an
    # item reads five of them: ~60 items/s per process, where one B200 trains > 200 clips/s.  From the cache an item costs the
    # float32 normalisation only.  What is cached is the integer image before `/255`, so items stay bit-identical.
    decode_cache_bytes = 32 << 30

    def __init__(self, folder, image_size, labels_scaling=None, selected_channels=[0, 1, 2, 3], num_frames=16, horizontal_flip=False,
                 force_num_frames=True, exts=['gif'], per_frame_cond=False, reference_frame='eulerian', decode_cache_bytes=None):
        super().__init__()
        self._cache, self._cache_used = {}, 0
        self._cache_lock = threading.Lock()          # preload() fills the cache from a thread pool
        if decode_cache_bytes is not None:
            self.decode_cache_bytes = int(decode_cache_bytes)
        if reference_frame not in ('eulerian', 'lagrangian'):
            raise ValueError(f'unknown reference_frame {reference_frame!r}')
        folder = str(folder)
        self.image_size, self.num_frames, self.selected_channels = image_size, num_frames, selected_channels
        self.horizontal_flip, self.force_num_frames = horizontal_flip, force_num_frames
        self.reference_frame = reference_frame
        lay = _LAYOUT[reference_frame]

        def listing(sub):
            paths = sorted((p for ext in exts for p in Path(folder + 'gifs/' + sub + '/').glob(f'**/*.{ext}')),
                           key=lambda p: int(p.name.split('.')[0]))
            assert all(int(p.stem) == i for i, p in enumerate(paths)), 'file position is not equal to index'
            return paths

        self.paths = {}
        for sub in lay['folders']:
            self.paths[sub] = listing(sub)
            assert len(self.paths[sub]) == len(self.paths['topo']), 'number of files in fields and top folders are not equal.'
        self.frame_ranges = torch.tensor(np.genfromtxt(folder + 'frame_range_data.csv', delimiter=','))
        rows = []
        for name, col, kind in lay['globals_']:
            v = torch.min(self.frame_ranges[:, col]) if kind == 'min' else torch.max(self.frame_ranges[:, col])
            setattr(self, name, v)
            rows.append([name, v.item()])
        self.zero_u_2 = self.normalize(torch.zeros(1), self.min_u_2, self.max_u_2) if reference_frame == 'lagrangian' else None
        with open(folder + 'min_max_values.csv', 'w', newline='') as f:
            csv.writer(f).writerows(rows)
        labels = np.genfromtxt(folder + 'stress_strain_data.csv', delimiter=',')
        if per_frame_cond:
            self.labels = torch.tensor(_interp_curves(labels, num_frames)).float()
        else:
            self.labels = torch.tensor(labels[:, 1:]).float()            # the first point of every curve is zero
        self.detached_labels = self.labels.clone().detach().numpy()
        self.labels_scaling = labels_scaling if labels_scaling is not None else \
            Normalization(self.labels, ['continuous'] * self.labels.shape[1], 'global-min-max-2')
        self.labels = self.labels_scaling.normalize(self.labels)

    def interpolate(self, tensor, num_frames):
        f = tensor.shape[1]
        if f == num_frames:
            return tensor
        if f > num_frames:
            return tensor[:, :num_frames]
        return torch.nn.functional.interpolate(tensor.unsqueeze(0), num_frames).squeeze(0)

    def normalize(self, arr, min_val, max_val):
        return (arr - min_val) / (max_val - min_val)

    def unnorm(self, arr, min_val, max_val):
        return arr * (max_val - min_val) + min_val

    def __len__(self):
        return len(self.paths['topo'])

    def __getstate__(self):          # DataLoader workers started with 'spawn' pickle the dataset: locks do not pickle
        st = dict(self.__dict__)
        st.pop('_cache_lock', None)
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._cache_lock = threading.Lock()

    def _frames_u8(self, sub, index) -> torch.Tensor:
        key = (sub, int(index))
        u8 = self._cache.get(key)
        if u8 is None:
            u8 = _gif_frames_u8(self.paths[sub][index], self.image_size, self.horizontal_flip)
            if not self.horizontal_flip:
                with self._cache_lock:
                    if key not in self._cache and self._cache_used + u8.numel() <= self.decode_cache_bytes:
                        self._cache[key] = u8
                        self._cache_used += u8.numel()
        return u8

    def _frames(self, sub, index) -> torch.Tensor:
        """(frames, h, w) float32 in [0, 1]: ToTensor of the decoded frames."""
        return self._frames_u8(sub, index).to(torch.float32).div(255)

    def preload(self, num_threads: int = 0) -> int:
        """Decode every GIF an item reads into the cache with a thread pool (PIL decodes outside the GIL); returns the number of
        cached bytes.  Optional: without it the cache fills during the first epoch."""
        from concurrent.futures import ThreadPoolExecutor
        key = 'lagrangian_1' if (self.reference_frame == 'lagrangian' and self.num_frames == 1) else self.reference_frame
        subs = sorted({'topo'} | {c[0] for c in _LAYOUT[key]['channels']})
        jobs = [(sub, i) for i in range(len(self)) for sub in subs]
        with ThreadPoolExecutor(max_workers=num_threads or min(32, os.cpu_count() or 1)) as pool:
            list(pool.map(lambda j: self._frames_u8(*j), jobs))
        return self._cache_used

    def __getitem__(self, index):
        key = self.reference_frame
        if key == 'lagrangian' and self.num_frames == 1:
            key = 'lagrangian_1'
            self.selected_channels = [0, 1]
        # all-ones where the topology has material, 0 in the void (0 / 255 is the only 8-bit value that maps to 0.0): AND-ing the
        # float bits with it writes an exact +0.0 without the data-dependent branch of a masked store (4x faster on noisy masks)
        keep = -(self._frames_u8('topo', index) != 0).to(torch.int32)
        r = self.frame_ranges[index]
        channels = _LAYOUT[key]['channels']
        planes = []
        for ch in self.selected_channels:                            # only the planes that are returned are computed
            sub, lo, hi, glo, ghi = channels[ch]
            t = self._frames(sub, index)
            if hi is not None:
                t = self.unnorm(t, r[lo] if lo is not None else 0., r[hi])     # physical value of this sample
                t.view(torch.int32).bitwise_and_(keep)                            # true zero of the field outside the material
                t = self.normalize(t, getattr(self, glo) if glo is not None else 0., getattr(self, ghi))
            planes.append(t)
        tensor = torch.stack(planes, dim=0)
        if self.force_num_frames:
            tensor = _cast_frames(tensor, self.num_frames)
        return tensor, self.labels[index, :]


class SyntheticLagrangianDataset(data.Dataset):
    """In-memory stand-in with the Dataset's value ranges: fields in [0, 1], labels in [-1, 1]."""

    def __init__(self, n: int, image_size: int, channels: int, num_frames: int, seed: int = 0):
        self.n, self.shape = n, (channels, num_frames, image_size, image_size)
        self.seed = seed
        self.labels_scaling = fixed_normalization(0.0, 100.0)
        self.zero_u_2 = torch.tensor([0.5])
        self.num_frames = num_frames

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        return torch.rand(self.shape, generator=g), torch.rand(self.num_frames, generator=g) * 2 - 1


_FOLDERS = ('topo', 'u_1', 'u_2', 's_mises', 's_22', 'ener')


def write_dataset(folder, fields: dict, frame_ranges: np.ndarray, curves: np.ndarray) -> None:
    """Write samples in the reference's on-disk layout (module docstring).  `fields[sub]` is a uint8 array (n, frames, h, w)
    for every sub-folder to create; `frame_ranges` is n x 8 (Lagrangian) or n x 4 (Eulerian), `curves` n x K."""
    from PIL import Image
    folder = str(folder)
    for sub, arr in fields.items():
        os.makedirs(os.path.join(folder, 'gifs', sub), exist_ok=True)
        for i in range(arr.shape[0]):
            frames = [Image.fromarray(np.ascontiguousarray(fr), 'L') for fr in arr[i]]
            frames[0].save(os.path.join(folder, 'gifs', sub, f'{i}.gif'), save_all=True, append_images=frames[1:], duration=200, loop=0)
    np.savetxt(os.path.join(folder, 'frame_range_data.csv'), np.asarray(frame_ranges), delimiter=',')
    np.savetxt(os.path.join(folder, 'stress_strain_data.csv'), np.asarray(curves), delimiter=',')


def synthetic_dataset_arrays(n: int, image_size: int = 96, num_frames: int = 11, curve_points: int = 51, seed: int = 0,
                             reference_frame: str = 'lagrangian', size_hw=None):
    """Random content for `write_dataset`: binary topology masks, uniformly random 8-bit fields (every frame differs: the GIF
    encoder would merge identical frames), plausible per-sample ranges and monotone stress curves."""
    rng = np.random.default_rng(seed)
    h, w = size_hw if size_hw is not None else (image_size, image_size)
    fields = {}
    for sub in (_FOLDERS if reference_frame == 'lagrangian' else ('topo', 's_mises', 's_22', 'ener')):
        if sub == 'topo':
            fields[sub] = (rng.random((n, num_frames, h, w)) > 0.3).astype(np.uint8) * 255
        else:
            fields[sub] = rng.integers(0, 256, (n, num_frames, h, w), dtype=np.uint8)
    u = rng.random((n, 8))
    ranges = np.stack([-u[:, 0], u[:, 1], -u[:, 2], u[:, 3], 50 + 50 * u[:, 4], -30 * u[:, 5], 30 * u[:, 6], u[:, 7]], axis=1)
    if reference_frame != 'lagrangian':
        ranges = ranges[:, 4:]
    curves = np.cumsum(rng.random((n, curve_points)), axis=1) * 2.0
    curves[:, 0] = 0.
    return fields, ranges, curves


def write_synthetic_dataset(folder, n: int, image_size: int = 96, num_frames: int = 11, curve_points: int = 51, seed: int = 0,
                            reference_frame: str = 'lagrangian') -> None:
    """Write `n` random samples in the reference's on-disk layout (the real dataset is an external download)."""
    write_dataset(folder, *synthetic_dataset_arrays(n, image_size, num_frames, curve_points, seed, reference_frame))


def video_tensor_to_gif(tensor: torch.Tensor, path: str, duration: int = 200, loop: int = 0, optimize: bool = False):
    """(channels, frames, h, w) in [0, 1] -> animated GIF (VDDP:1091-1098).  Pixels are `floor(255 x)` like the reference's
    `T.ToPILImage` (`mul(255).byte()`); values outside [0, 1] are clamped here (the reference lets them wrap modulo 256)."""
    from PIL import Image
    frames = []
    for fr in tensor.unbind(dim=1):
        arr = fr.detach().float().clamp(0, 1).mul(255).to(torch.uint8).cpu().numpy()
        img = Image.fromarray(arr[0]) if arr.shape[0] == 1 else Image.fromarray(np.ascontiguousarray(np.moveaxis(arr, 0, -1)))
        frames.append(img.convert('L').convert('P') if not optimize else img)
    frames[0].save(path, save_all=True, append_images=frames[1:], duration=duration, loop=loop, optimize=optimize)
    return frames


def clean_pred(geom: np.ndarray, pixels: int) -> np.ndarray:
    """Post-processing of predicted geometries (src/utils.py:32-82): binarise at 0.5, drop pixels without a 4-neighbour, keep
    the largest 4-connected component; returns (n, pixels*pixels) ints.

    The reference does this with a Python double loop and a networkx graph per sample (O(pixels^2) interpreter work, the
    dominant cost of `eval_target` once sampling is fast); here it is one `scipy.ndimage.label` pass.  Components of equal
    size are resolved as networkx enumerates them there: components containing an axis-0 edge first, by the raster position
    of their first such edge, then the single-row components by their first axis-1 edge.  A geometry without any pair of
    adjacent pixels comes back empty (the reference raises IndexError on it)."""
    from scipy import ndimage
    g = (np.asarray(geom).reshape(-1, pixels, pixels) > 0.5).astype(int)
    out = np.zeros_like(g)
    big = 2 * pixels * pixels
    raster = np.arange(pixels * pixels).reshape(pixels, pixels)
    for i in range(g.shape[0]):
        cur = g[i]
        pad = np.pad(cur, 1)
        neigh = pad[:-2, 1:-1] + pad[2:, 1:-1] + pad[1:-1, :-2] + pad[1:-1, 2:]
        cur = cur * (neigh > 0)
        lab, k = ndimage.label(cur)
        if k == 0:
            continue
        sizes = np.bincount(lab.ravel(), minlength=k + 1)[1:]
        cand = np.flatnonzero(sizes == sizes.max()) + 1
        if len(cand) > 1:
            key0, key1 = np.full(k + 1, big), np.full(k + 1, big)
            e0 = (cur[1:] & cur[:-1]).astype(bool)
            np.minimum.at(key0, lab[:-1][e0], raster[:-1][e0])
            e1 = (cur[:, 1:] & cur[:, :-1]).astype(bool)
            np.minimum.at(key1, lab[:, :-1][e1], raster[:, :-1][e1])
            key = np.where(key0 < big, key0, pixels * pixels + key1)
            cand = cand[np.argsort(key[cand], kind='stable')]
        out[i] = (lab == cand[0]).astype(int)
    return out.reshape(g.shape[0], -1)
