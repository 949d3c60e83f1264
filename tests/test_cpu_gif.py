"""Host side of the device dataset, without a GPU: the native container scanner `vmm_gif_scan` and the
contract statements of `vmm_gif_decode` / `vmm_dataset_items` (tests/emu_gif.py) against PIL, which is what the reference decodes
with (VDDP:1076-1106), and against `Dataset.__getitem__` (pinned bit-equal to the reference's Dataset by tests/golden/host_dataset.pt).
The `-m gpu` twin (tests/test_gpu_dataset.py) holds the kernels to the same checkers on the same corpus."""
import numpy as np
import pytest
import torch

import emu_gif


def _scan(blob):
    from videometamaterials_b200.device_dataset import scan_gif
    return scan_gif(blob)


def test_scan_and_decode_statement_match_pil_on_the_corpus():
    seen = dict(interlace=0, transp=0, disposal2=0, local_palette=0, partial=0, small_code=0)
    for name, blob in emu_gif.corpus():
        ref = emu_gif.pil_frames(blob)
        if name == 'P_colour_transparency':
            continue                                   # rejected, see test_scan_rejects_what_the_compositor_does_not_reproduce
        (w, h), tab = _scan(blob)
        assert (h, w) == ref.shape[1:] and len(tab) == len(ref), name
        got = emu_gif.decode(blob, tab, (w, h))
        assert np.array_equal(got, ref), name
        seen['interlace'] += int(tab['interlace'].any())
        seen['transp'] += int(tab['has_transp'].any())
        seen['disposal2'] += int((tab['disposal'] == 2).any())
        seen['local_palette'] += int((tab['pal_ofs'] != 13).any())
        seen['partial'] += int(((tab['w'] != w) | (tab['h'] != h)).any())
        seen['small_code'] += int((tab['min_code'] < 8).any())
    assert all(v > 0 for v in seen.values()), seen     # the corpus really exercises every branch of the compositor


def test_pillow_mode_l_palette_quirk_is_reproduced_and_can_be_turned_off():
    """A file whose first frame carries the plain grey ramp and whose second frame brings a compacted local palette: Pillow decodes the
    second frame as raw indices (it does not round-trip the file it wrote); the scanner reproduces that by default, because the
    reference reads through Pillow, and maps every frame through its own palette with pil_compat=False (= the encoded image)."""
    from PIL import Image
    import io
    from videometamaterials_b200.device_dataset import scan_gif
    rng = np.random.default_rng(5)
    hit = 0
    for _ in range(12):
        arrs = emu_gif.quirk_frames(rng)
        bio = io.BytesIO()
        frames = [Image.fromarray(a, 'L') for a in arrs]
        frames[0].save(bio, format='GIF', save_all=True, append_images=frames[1:], duration=200, loop=0)
        blob = bio.getvalue()
        ref = emu_gif.pil_frames(blob)
        (w, h), tab = scan_gif(blob)
        assert np.array_equal(emu_gif.decode(blob, tab, (w, h)), ref)
        (w, h), tab_true = scan_gif(blob, pil_compat=False)
        assert np.array_equal(emu_gif.decode(blob, tab_true, (w, h)), np.stack(arrs))
        hit += int(not np.array_equal(ref, np.stack(arrs)))
    assert hit > 0          # Pillow really mis-decodes some of these files: the quirk is exercised, not hypothetical


def test_frames_beyond_a_files_count_are_zero_and_extra_frames_are_dropped():
    name, blob = [c for c in emu_gif.corpus() if c[0] == 'L_boxes'][0]
    (w, h), tab = _scan(blob)
    ref = emu_gif.pil_frames(blob)
    more = emu_gif.decode(blob, tab, (w, h), frames_per_file=len(tab) + 2)
    assert np.array_equal(more[:len(tab)], ref) and not more[len(tab):].any()
    fewer = emu_gif.decode(blob, tab[:3], (w, h), frames_per_file=3)
    assert np.array_equal(fewer, ref[:3])


def test_scan_rejects_what_the_compositor_does_not_reproduce():
    from videometamaterials_b200 import _lib
    rng = np.random.default_rng(0)
    gp = rng.integers(0, 256, (4, 3), dtype=np.uint8)
    idx = rng.integers(0, 4, (8, 8), dtype=np.uint8)
    cases = {
        'transparent first frame': emu_gif.write_gif(8, 8, gp, [dict(x=0, y=0, idx=idx, transp=1)]),
        'disposal method 3': emu_gif.write_gif(8, 8, gp, [dict(x=0, y=0, idx=idx), dict(x=0, y=0, idx=idx, disposal=3)]),
        'outside the logical screen': emu_gif.write_gif(8, 8, gp, [dict(x=4, y=0, idx=idx)]),
        'not a GIF': b'PNG' + bytes(40),
        'no image': emu_gif.write_gif(8, 8, gp, []),
    }
    for what, blob in cases.items():
        with pytest.raises(_lib.VmmError, match=what):
            _scan(blob)
    # a file cut inside its image data still yields its frame table (PIL opens such files too); the DEVICE decode reports the short stream
    good = emu_gif.write_gif(8, 8, gp, [dict(x=0, y=0, idx=idx)] * 2)
    (w, h), tab = _scan(good[:-12])
    assert len(tab) == 2
    _, complete = emu_gif.lzw(good[:-12], tab[1])
    assert not complete


@pytest.mark.parametrize("frame,num_frames,channels", [('lagrangian', 11, [0, 1, 2, 3]), ('lagrangian', 11, [0, 1, 3]), ('eulerian', 11, [0, 1, 2, 3]),
                                                       ('lagrangian', 1, [0, 1]), ('lagrangian', 14, [1, 3]), ('lagrangian', 5, [0, 2])])
def test_item_tables_and_statement_equal_dataset_items(tmp_path, frame, num_frames, channels):
    """`item_tables` (product, host) + the statement of `vmm_dataset_items` reproduce Dataset.__getitem__ BIT FOR BIT: this pins the
    fp32 table entries (float64 spans converted after the subtraction) and the order of the separately rounded operations."""
    from videometamaterials_b200.dataset import Dataset, write_synthetic_dataset
    from videometamaterials_b200.device_dataset import item_tables
    folder = str(tmp_path / "d") + "/"
    write_synthetic_dataset(folder, 5, image_size=16, num_frames=11, seed=3, reference_frame=frame)
    ds = Dataset(folder, 16, selected_channels=list(channels), num_frames=num_frames, per_frame_cond=True, reference_frame=frame)
    planes, ch_plane, ch_has, srng, grng = item_tables(ds)
    assert planes[0] == 'topo'
    u8 = np.stack([np.stack([ds._frames_u8(sub, i).numpy() for sub in planes]) for i in range(len(ds))])
    frames = u8.shape[2]
    got = emu_gif.dataset_items(u8, [4, 0, 2], 0, ch_plane, ch_has, srng.numpy(), grng.numpy(), np.full(len(ds), frames), num_frames)
    want = torch.stack([ds[i][0] for i in (4, 0, 2)]).numpy()
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.int32), want.view(np.int32))


def test_device_dataset_has_no_host_path():
    from videometamaterials_b200.device_dataset import DeviceDataset
    with pytest.raises(TypeError):
        DeviceDataset(object())


def test_device_loader_order_logic_without_a_device():
    """DeviceLoader's index bookkeeping (host logic only; a stand-in dataset records the index batches): RandomSampler's draw for one
    rank, disjoint strided shares of one permutation padded by wrapping for two, a new permutation per pass, drop_last."""
    from videometamaterials_b200.device_dataset import DeviceLoader

    class Stub:
        def __len__(self):
            return 7

        def batch(self, idx):
            return idx.tolist(), None

    g = torch.Generator().manual_seed(11)
    dl = DeviceLoader(Stub(), 3, True, False, g, 0, 1)
    first = [b for b, _ in dl]
    assert [len(b) for b in first] == [3, 3, 1] and sum(first, []) == torch.randperm(7, generator=torch.Generator().manual_seed(11)).tolist()
    assert sorted(sum([b for b, _ in dl], [])) == list(range(7))
    assert [b for b, _ in DeviceLoader(Stub(), 4, False, False, None, 0, 1)] == [[0, 1, 2, 3], [4, 5, 6]]
    assert len(DeviceLoader(Stub(), 3, True, True, None, 0, 1)) == 2
    r0 = sum([b for b, _ in DeviceLoader(Stub(), 2, True, False, None, 0, 2)], [])
    r1 = sum([b for b, _ in DeviceLoader(Stub(), 2, True, False, None, 1, 2)], [])
    assert len(r0) == len(r1) == 4 and sorted(set(r0 + r1)) == list(range(7)) and len(set(r0) & set(r1)) <= 1      # 8 slots for 7 samples: one wraps
    e0 = DeviceLoader(Stub(), 2, True, False, None, 0, 2)
    assert sum([b for b, _ in e0], []) != sum([b for b, _ in e0], [])                                               # reshuffled every pass


def test_decode_files_pass_by_pass_assembles_one_tensor():
    """decode_files reads / decodes / releases a large file list in passes (host logic; the decode and the file reads are stand-ins here):
    same result as one pass, every file read exactly once, the per-file frame counts in file order."""
    from videometamaterials_b200.device_dataset import decode_files
    names = [f"f{i}" for i in range(23)]
    reads = []

    def read(nm):
        reads.append(nm)
        return bytes([int(nm[1:])])

    def decode(blobs, fpf, size_hw, device, nms, pil_compat=True):
        assert [bytes([int(n[1:])]) for n in nms] == list(blobs)
        out = torch.stack([torch.full((fpf,) + tuple(size_hw), b[0], dtype=torch.uint8) for b in blobs])
        return out, torch.tensor([b[0] % 5 for b in blobs], dtype=torch.int32)

    whole, c_whole = decode_files(names, 3, (2, 2), "cpu", files_per_pass=100, decode=decode, read=read)
    assert reads == names
    reads.clear()
    parts, c_parts = decode_files(names, 3, (2, 2), "cpu", files_per_pass=5, decode=decode, read=read)
    assert reads == names and torch.equal(parts, whole) and torch.equal(c_parts, c_whole) and parts.shape == (23, 3, 2, 2)
    assert [int(parts[i, 0, 0, 0]) for i in range(23)] == list(range(23))
    reads.clear()
    every, _ = decode_files(names, None, (2, 2), "cpu", files_per_pass=5, decode=lambda b, f, s, d, n, pil_compat=True: decode(b, 3, s, d, n), read=read)
    assert reads == names and torch.equal(every, whole)          # no frame count given: one pass over everything
