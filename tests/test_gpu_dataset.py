"""The device dataset on a B200, through the C ABI: `vmm_gif_decode` against PIL on the corpus of
tests/emu_gif.py, `DeviceDataset.batch` (= `vmm_dataset_items`) bit for bit against `Dataset.__getitem__`, the loader, and a Trainer
that trains from the HBM-resident set."""
import numpy as np
import pytest
import torch

import emu_gif

pytestmark = pytest.mark.gpu


def test_gif_decode_matches_pil_on_the_corpus():
    from videometamaterials_b200 import _lib
    from videometamaterials_b200.device_dataset import decode_gifs
    n0 = _lib.launch_count()
    checked = 0
    for name, blob in emu_gif.corpus():
        if name == 'P_colour_transparency':
            with pytest.raises(_lib.VmmError, match='transparent first frame'):
                decode_gifs([blob], None, (40, 56), 'cuda')
            continue
        ref = emu_gif.pil_frames(blob)
        out, counts = decode_gifs([blob], None, ref.shape[1:], 'cuda', [name])
        assert int(counts[0]) == len(ref), name
        got = out[0].cpu().numpy()
        assert got.shape == ref.shape and np.array_equal(got, ref), (name, int((got != ref).sum()))
        checked += 1
    assert checked >= 30 and _lib.launch_count() >= n0 + 2 * checked


def test_gif_decode_many_files_in_chunks_and_padded_frames():
    """Files of one size in ONE call, split into several launches by a small chunk budget; frames_per_file above / below the files'
    frame counts (zero frames appended / extra frames dropped); the decode is deterministic."""
    from videometamaterials_b200.device_dataset import decode_gifs
    from PIL import Image
    import io
    rng = np.random.default_rng(5)
    blobs, refs = [], []
    for i in range(23):
        n = int(rng.integers(1, 7))
        frames = [Image.fromarray(rng.integers(0, 256, (48, 64), dtype=np.uint8), 'L') for _ in range(n)]
        bio = io.BytesIO()
        frames[0].save(bio, format='GIF', save_all=True, append_images=frames[1:], duration=200, loop=0)
        blobs.append(bio.getvalue())
        refs.append(emu_gif.pil_frames(blobs[-1]))
    out, counts = decode_gifs(blobs, 4, (48, 64), 'cuda', chunk_bytes=40_000)
    out2, _ = decode_gifs(blobs, 4, (48, 64), 'cuda')
    assert torch.equal(out, out2)
    got = out.cpu().numpy()
    for i, ref in enumerate(refs):
        k = min(len(ref), 4)
        assert int(counts[i]) == k
        assert np.array_equal(got[i, :k], ref[:k]), i
        assert not got[i, k:].any()


def test_pil_compat_off_decodes_the_encoded_image():
    """The file Pillow does not round-trip (tests/test_cpu_gif.py): default = what Pillow (the reference) reads, pil_compat=False = the frames
    that were written."""
    from videometamaterials_b200.device_dataset import decode_gifs
    from PIL import Image
    import io
    rng = np.random.default_rng(5)
    differs = 0
    for _ in range(8):
        arrs = np.stack(emu_gif.quirk_frames(rng))
        frames = [Image.fromarray(a, 'L') for a in arrs]
        bio = io.BytesIO()
        frames[0].save(bio, format='GIF', save_all=True, append_images=frames[1:], duration=200, loop=0)
        blob = bio.getvalue()
        ref = emu_gif.pil_frames(blob)
        a = decode_gifs([blob], None, (48, 64), 'cuda')[0][0].cpu().numpy()
        b = decode_gifs([blob], None, (48, 64), 'cuda', pil_compat=False)[0][0].cpu().numpy()
        assert np.array_equal(a, ref) and np.array_equal(b, arrs)
        differs += int(not np.array_equal(a, b))
    assert differs > 0


def test_short_lzw_stream_is_reported():
    from videometamaterials_b200 import _lib
    from videometamaterials_b200.device_dataset import decode_gifs
    rng = np.random.default_rng(0)
    gp = rng.integers(0, 256, (4, 3), dtype=np.uint8)
    idx = rng.integers(0, 4, (8, 8), dtype=np.uint8)
    good = emu_gif.write_gif(8, 8, gp, [dict(x=0, y=0, idx=idx)] * 2)
    with pytest.raises(_lib.VmmError, match='short or invalid LZW'):
        decode_gifs([good[:-12]], None, (8, 8), 'cuda')


@pytest.mark.parametrize("frame,num_frames,channels,size", [('lagrangian', 11, [0, 1, 2, 3], 96), ('lagrangian', 11, [0, 1, 3], 32), ('eulerian', 11, [0, 1, 2, 3], 32),
                                                            ('lagrangian', 1, [0, 1], 32), ('lagrangian', 14, [1, 3], 16), ('lagrangian', 5, [0, 2], 16)])
def test_device_dataset_items_bit_equal_dataset(tmp_path, frame, num_frames, channels, size):
    from videometamaterials_b200.dataset import Dataset, write_synthetic_dataset
    from videometamaterials_b200.device_dataset import DeviceDataset
    folder = str(tmp_path / "d") + "/"
    n = 9
    write_synthetic_dataset(folder, n, image_size=size, num_frames=11, seed=3, reference_frame=frame)
    ds = Dataset(folder, size, selected_channels=list(channels), num_frames=num_frames, per_frame_cond=True, reference_frame=frame)
    dds = DeviceDataset(ds)
    assert len(dds) == n and dds.u8.dtype == torch.uint8 and dds.u8.is_cuda
    order = [8, 0, 3, 3, 5]
    x, lab = dds.batch(torch.tensor(order))
    want = torch.stack([ds[i][0] for i in order])
    want_lab = torch.stack([ds[i][1] for i in order])
    assert x.shape == want.shape and x.dtype == torch.float32 and x.is_cuda
    assert torch.equal(x.cpu().view(torch.int32), want.view(torch.int32))          # bit for bit, including the sign of zeros
    assert torch.equal(lab.cpu(), want_lab)
    xi, li = dds[4]
    assert torch.equal(xi.cpu(), ds[4][0]) and torch.equal(li.cpu(), ds[4][1])


def test_device_loader_order_and_sharding(tmp_path):
    from videometamaterials_b200.dataset import Dataset, write_synthetic_dataset
    from videometamaterials_b200.device_dataset import DeviceDataset
    folder = str(tmp_path / "d") + "/"
    write_synthetic_dataset(folder, 7, image_size=16, num_frames=11, seed=1)
    ds = Dataset(folder, 16, selected_channels=[0, 1, 3], num_frames=11, per_frame_cond=True, reference_frame='lagrangian')
    dds = DeviceDataset(ds)
    items = torch.stack([ds[i][0] for i in range(7)])

    def which(x):           # sample index of every clip of a batch
        return [int(torch.nonzero((items.view(7, -1) == c.cpu().view(1, -1)).all(dim=1))[0]) for c in x]

    g = torch.Generator().manual_seed(11)
    dl = dds.loader(3, shuffle=True, generator=g)
    assert len(dl) == 3
    seen = [which(x) for x, _ in dl]
    assert [len(s) for s in seen] == [3, 3, 1] and sorted(sum(seen, [])) == list(range(7))
    assert sum(seen, []) == torch.randperm(7, generator=torch.Generator().manual_seed(11)).tolist()      # RandomSampler's draw
    again = [which(x) for x, _ in dl]
    assert sorted(sum(again, [])) == list(range(7)) and again != seen                                       # a new permutation per pass
    assert [which(x) for x, _ in dds.loader(4, shuffle=False)] == [[0, 1, 2, 3], [4, 5, 6]]
    assert len(dds.loader(3, drop_last=True)) == 2
    # two ranks: disjoint strided shares of one permutation, padded to equal length by wrapping
    r0 = sum((which(x) for x, _ in dds.loader(2, shuffle=True, rank=0, world_size=2)), [])
    r1 = sum((which(x) for x, _ in dds.loader(2, shuffle=True, rank=1, world_size=2)), [])
    assert len(r0) == len(r1) == 4 and sorted(set(r0 + r1)) == list(range(7))


def test_trainer_trains_from_the_device_dataset(tmp_path, monkeypatch):
    from videometamaterials_b200 import Accelerator, GaussianDiffusion, Trainer, Unet3D, _lib
    from videometamaterials_b200.dataset import write_synthetic_dataset
    monkeypatch.chdir(tmp_path)
    data = str(tmp_path / "data") + "/"
    write_synthetic_dataset(data + "training/", 6, image_size=16, num_frames=11, seed=0)
    write_synthetic_dataset(data + "validation/", 2, image_size=16, num_frames=11, seed=1)

    def make(device_dataset):
        torch.manual_seed(0)
        model = Unet3D(dim=16, dim_mults=(1, 2), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
                       cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True)
        gd = GaussianDiffusion(model, image_size=16, channels=3, num_frames=11, timesteps=8, loss_type='l1', use_dynamic_thres=True,
                               sampling_timesteps=8)
        return Trainer(gd, folder=data + "training/", validation_folder=data + "validation/", results_folder='run', selected_channels=[0, 1, 3],
                       train_batch_size=2, test_batch_size=4, train_lr=1e-3, save_and_sample_every=100, train_num_steps=4, ema_decay=0.9,
                       step_start_ema=2, update_ema_every=2, log=True, null_cond_prob=0., per_frame_cond=True, reference_frame='lagrangian',
                       run_name='dd' if device_dataset else 'host', accelerator=Accelerator(mixed_precision='fp16'), log_every=1,
                       device_dataset=device_dataset)

    tr = make(True)
    assert tr.dds is not None and tr.dds.u8.shape == (6, 4, 11, 16, 16)      # topo + the three selected fields
    x, cond = next(tr.dl)
    assert x.is_cuda and x.shape == (2, 3, 11, 16, 16) and cond.shape == (2, 11) and cond.is_cuda
    # the batch is made of the host Dataset's items
    items = torch.stack([tr.ds[i][0] for i in range(6)]).view(6, -1)
    for c in x:
        assert bool((items == c.cpu().view(1, -1)).all(dim=1).any())
    n0 = _lib.launch_count()
    tr.train()
    torch.cuda.synchronize()
    losses = [l["training loss"] for l in tr.accelerator.logs if "training loss" in l]
    assert len(losses) == 5 and all(np.isfinite(losses)) and _lib.launch_count() > n0 + 500
    # without a CUDA device the option refuses instead of decoding on the host
    with pytest.raises(RuntimeError, match='CUDA'):
        from videometamaterials_b200.device_dataset import DeviceDataset
        DeviceDataset(tr.ds, device='cpu')
