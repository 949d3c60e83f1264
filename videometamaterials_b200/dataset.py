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


def _gif_frames(path, image_size: int) -> torch.Tensor:
    """(frames, h, w) float in [0, 1] from an 8-bit GIF (T.Resize + CenterCrop + ToTensor of the reference)."""
    from PIL import Image
    img = Image.open(path)
    frames, i = [], 0
    while True:
        try:
            img.seek(i)
        except EOFError:
            break
        fr = img.convert('L')
        if fr.size != (image_size, image_size):
            fr = fr.resize((image_size, image_size), Image.BILINEAR)
        frames.append(torch.from_numpy(np.asarray(fr, dtype=np.uint8).copy()).float() / 255.)
        i += 1
    return torch.stack(frames, dim=0)


def _cast_frames(t: torch.Tensor, frames: int) -> torch.Tensor:
    f = t.shape[1]
    if f == frames:
        return t
    if f > frames:
        return t[:, :frames]
    return torch.nn.functional.pad(t, (0, 0, 0, 0, 0, frames - f))


class Dataset(data.Dataset):
    def __init__(self, folder, image_size, labels_scaling=None, selected_channels=(0, 1, 2, 3), num_frames=11, per_frame_cond=True,
                 reference_frame='lagrangian', **unused):
        super().__init__()
        if reference_frame != 'lagrangian' or not per_frame_cond or num_frames == 1:
            raise NotImplementedError("only the shipped configuration (lagrangian frame, per-frame labels) is implemented")
        folder = str(folder)
        if not folder.endswith('/'):
            folder += '/'
        self.image_size, self.num_frames, self.selected_channels = image_size, num_frames, list(selected_channels)
        self.reference_frame = reference_frame

        def listing(sub):
            paths = sorted(Path(folder + 'gifs/' + sub + '/').glob('*.gif'), key=lambda p: int(p.stem))
            assert all(int(p.stem) == i for i, p in enumerate(paths)), 'file position is not equal to index'
            return paths

        self.paths = {k: listing(k) for k in ('topo', 'u_1', 'u_2', 's_mises', 's_22')}
        n = len(self.paths['topo'])
        assert all(len(v) == n for v in self.paths.values()), 'number of files in fields and top folders are not equal.'
        self.frame_ranges = torch.tensor(np.genfromtxt(folder + 'frame_range_data.csv', delimiter=',')).reshape(n, -1)
        fr = self.frame_ranges
        self.min_u_1, self.max_u_1 = fr[:, 0].min(), fr[:, 1].max()
        self.min_u_2, self.max_u_2 = fr[:, 2].min(), fr[:, 3].max()
        self.max_s_mises = fr[:, 4].max()
        self.min_s_22, self.max_s_22 = fr[:, 5].min(), fr[:, 6].max()
        self.max_strain_energy = fr[:, 7].max()
        self.zero_u_2 = self._norm(torch.zeros(1), self.min_u_2, self.max_u_2)
        with open(folder + 'min_max_values.csv', 'w', newline='') as f:
            csv.writer(f).writerows([['min_u_1', self.min_u_1.item()], ['max_u_1', self.max_u_1.item()], ['min_u_2', self.min_u_2.item()],
                                     ['max_u_2', self.max_u_2.item()], ['max_s_mises', self.max_s_mises.item()],
                                     ['min_s_22', self.min_s_22.item()], ['max_s_22', self.max_s_22.item()],
                                     ['max_strain_energy', self.max_strain_energy.item()]])
        labels = np.genfromtxt(folder + 'stress_strain_data.csv', delimiter=',').reshape(n, -1)
        strain = 0.2
        given = np.linspace(0., strain, num=labels.shape[1])
        ev = np.linspace(0., strain, num=num_frames)
        ev[0] = 0.01 * strain
        labels = np.array([np.interp(ev, given, labels[i, :]) for i in range(n)])
        self.labels = torch.tensor(labels).float()
        self.labels_scaling = labels_scaling if labels_scaling is not None else Normalization(self.labels)
        self.labels = self.labels_scaling.normalize(self.labels)

    @staticmethod
    def _norm(a, lo, hi):
        return (a - lo) / (hi - lo)

    @staticmethod
    def _unnorm(a, lo, hi):
        return a * (hi - lo) + lo

    def __len__(self):
        return len(self.paths['topo'])

    def __getitem__(self, i):
        g = lambda k: _gif_frames(self.paths[k][i], self.image_size)
        topo = g('topo')
        t = torch.stack((g('u_1'), g('u_2'), g('s_mises'), g('s_22')), dim=0).double()
        r = self.frame_ranges[i]
        t[0] = self._unnorm(t[0], r[0], r[1])
        t[1] = self._unnorm(t[1], r[2], r[3])
        t[2] = self._unnorm(t[2], 0., r[4])
        t[3] = self._unnorm(t[3], r[5], r[6])
        t[:, topo == 0.] = 0.                                            # void pixels carry the true zero of each field
        t[0] = self._norm(t[0], self.min_u_1, self.max_u_1)
        t[1] = self._norm(t[1], self.min_u_2, self.max_u_2)
        t[2] = self._norm(t[2], 0., self.max_s_mises)
        t[3] = self._norm(t[3], self.min_s_22, self.max_s_22)
        return _cast_frames(t[self.selected_channels].float(), self.num_frames), self.labels[i, :]


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


def write_synthetic_dataset(folder, n: int, image_size: int = 96, num_frames: int = 11, curve_points: int = 51, seed: int = 0) -> None:
    """Write `n` random samples in the reference's on-disk layout (see module docstring)."""
    from PIL import Image
    folder = str(folder)
    rng = np.random.default_rng(seed)
    for sub in ('topo', 'u_1', 'u_2', 's_mises', 's_22', 'ener'):
        os.makedirs(os.path.join(folder, 'gifs', sub), exist_ok=True)
    ranges, curves = [], []
    for i in range(n):
        for sub in ('topo', 'u_1', 'u_2', 's_mises', 's_22', 'ener'):
            # every frame differs (identical GIF frames would be merged by the encoder); topology frames are binary masks
            frames = [Image.fromarray(((rng.random((image_size, image_size)) > 0.3).astype(np.uint8) * 255) if sub == 'topo'
                                      else rng.integers(0, 256, (image_size, image_size), dtype=np.uint8), 'L')
                      for _ in range(num_frames)]
            frames[0].save(os.path.join(folder, 'gifs', sub, f'{i}.gif'), save_all=True, append_images=frames[1:], duration=200, loop=0)
        ranges.append([-rng.random(), rng.random(), -rng.random(), rng.random(), 50 + 50 * rng.random(), -30 * rng.random(), 30 * rng.random(), rng.random()])
        curves.append(np.cumsum(rng.random(curve_points)) * 2.0)
    np.savetxt(os.path.join(folder, 'frame_range_data.csv'), np.array(ranges), delimiter=',')
    np.savetxt(os.path.join(folder, 'stress_strain_data.csv'), np.array(curves), delimiter=',')


def video_tensor_to_gif(tensor: torch.Tensor, path: str, duration: int = 200, loop: int = 0):
    """(channels, frames, h, w) in [0, 1] -> animated GIF (VDDP:1091-1098)."""
    from PIL import Image
    frames = []
    for fr in tensor.unbind(dim=1):
        arr = (fr.clamp(0, 1) * 255).round().to(torch.uint8).cpu().numpy()
        arr = arr[0] if arr.shape[0] == 1 else np.moveaxis(arr, 0, -1)
        frames.append(Image.fromarray(arr).convert('L').convert('P'))
    frames[0].save(path, save_all=True, append_images=frames[1:], duration=duration, loop=loop, optimize=False)


def clean_pred(geom: np.ndarray, pixels: int) -> np.ndarray:
    """Binarise, drop pixels without any 4-neighbour, keep the largest 4-connected component (src/utils.py:32-82);
    returns (n, pixels*pixels) ints."""
    from scipy import ndimage
    g = (np.asarray(geom).reshape(-1, pixels, pixels) > 0.5).astype(int)
    out = np.zeros_like(g)
    for i in range(g.shape[0]):
        cur = g[i].copy()
        pad = np.pad(cur, 1)
        neigh = pad[:-2, 1:-1] + pad[2:, 1:-1] + pad[1:-1, :-2] + pad[1:-1, 2:]
        cur[(cur == 1) & (neigh == 0)] = 0
        lab, k = ndimage.label(cur)
        if k > 0:
            sizes = ndimage.sum(cur, lab, index=np.arange(1, k + 1))
            cur = (lab == (1 + int(np.argmax(sizes)))).astype(int)
        out[i] = cur
    return out.reshape(g.shape[0], -1)
