"""Host-side logic without a GPU: the C-ABI library loads and exports every declared symbol, the module tree has
the reference's state_dict, packing / tap tables are consistent, and the product refuses to run on CPU."""
import ctypes
import json
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from videometamaterials_b200 import _lib
    header = open(os.path.join(ROOT, "include", "vmm.h")).read()
    declared = set(re.findall(r"\b(vmm_[a-z0-9_]+)\s*\(", header))
    declared -= {"vmm_view4", "vmm_tap"}
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/vmm.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert _lib.lib.vmm_abi_version() == 1


def test_compute_call_without_gpu_fails_loudly():
    from videometamaterials_b200 import Unet3D, ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        ops.cgemm([torch.zeros(1, 1, 128, 64, dtype=torch.bfloat16)], [[(0, 0, 0, 0, 64)]], torch.zeros(64, 64, dtype=torch.bfloat16), 64,
                  torch.zeros(128, 64, dtype=torch.bfloat16), (1, 1, 128))
    m = Unet3D(dim=16, dim_mults=(1, 2), per_frame_cond=True, use_temporal_attention_cond=True, cond_attention='self-stacked')
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 11, 16, 16), torch.zeros(1, dtype=torch.long), cond=torch.zeros(1, 11))


def test_state_dict_matches_reference_layout(golden_dir):
    from videometamaterials_b200 import GaussianDiffusion, Unet3D
    kat = json.load(open(os.path.join(golden_dir, "kat.json")))
    m = Unet3D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
               cond_bias=True, cond_attention='self-stacked', cond_attention_tokens=16, use_temporal_attention_cond=True, cond_to_time='add',
               per_frame_cond=True, padding_mode='zeros')
    sd = m.state_dict()
    assert list(sd.keys()) == kat["full_state_dict_keys"]
    assert [list(v.shape) for v in sd.values()] == kat["full_state_dict_shapes"]
    assert sum(p.numel() for p in m.parameters()) == kat["full_param_count"]
    gd = GaussianDiffusion(m, image_size=96, channels=3, num_frames=11, timesteps=256, use_dynamic_thres=True, sampling_timesteps=256)
    assert len(gd.state_dict()) == 377 + 12
    for k, vals in kat["buffers_0_128_255"].items():
        assert [float(getattr(gd, k)[i]) for i in (0, 128, 255)] == vals
    assert not gd.is_ddim_sampling
    gd2 = GaussianDiffusion(m, image_size=96, channels=3, num_frames=11, timesteps=256, sampling_timesteps=250)
    assert gd2.is_ddim_sampling


def test_unsupported_configurations_raise():
    from videometamaterials_b200 import Unet3D
    with pytest.raises(NotImplementedError):
        Unet3D(dim=16, per_frame_cond=False)
    with pytest.raises(NotImplementedError):
        Unet3D(dim=16, per_frame_cond=True, use_temporal_attention_cond=True, padding_mode='circular')


def test_packing_and_tap_tables():
    from videometamaterials_b200 import ops
    w = torch.arange(2 * 5 * 3 * 3, dtype=torch.float32).reshape(2, 5, 3, 3)
    p = ops.pack_conv_taps(w, [3, 2], torch.float32)
    taps, ktot = ops.taps_conv(3, 3, [3, 2], 1)
    assert p.shape == (16, ktot) and ktot == 9 * 128
    for (src, dy, dx, kofs, c) in taps:
        c0 = 0 if src == 0 else 3
        assert torch.equal(p[:2, kofs:kofs + c], w[:, c0:c0 + c, dy + 1, dx + 1])
        assert float(p[:2, kofs + c:kofs + 64].abs().sum()) == 0
    assert len(ops.down_taps(64)) == 16 and {t[0] for t in ops.down_taps(64)} == {0, 1, 2, 3}
    phases, offs = ops.up_taps(64)
    assert len(phases) == 4 and all(len(ph) == 4 for ph in phases) and offs == [(0, 0), (0, 1), (1, 0), (1, 1)]
    # every (ky, kx) of the 4x4 kernel is used exactly once across the 4 output phases
    assert sorted(t[3] // 64 for ph in phases for t in ph) == list(range(16))
    assert ops.choose_tile(88, 96, 96)[0] * ops.choose_tile(88, 96, 96)[1] * ops.choose_tile(88, 96, 96)[2] == 128
    assert ops.choose_tile(1, 1, 5000) == (1, 1, 128)


def test_dataset_roundtrip_and_postprocessing(tmp_path):
    from videometamaterials_b200.dataset import Dataset, clean_pred, write_synthetic_dataset
    import numpy as np
    folder = str(tmp_path / "training") + "/"
    write_synthetic_dataset(folder, 3, image_size=16, num_frames=11)
    ds = Dataset(folder, 16, selected_channels=[0, 1, 3], num_frames=11, per_frame_cond=True, reference_frame='lagrangian')
    x, lab = ds[1]
    assert x.shape == (3, 11, 16, 16) and lab.shape == (11,)
    assert float(x.min()) >= 0.0 and float(x.max()) <= 1.0 and float(lab.abs().max()) <= 1.0 + 1e-6
    assert os.path.exists(folder + "min_max_values.csv")
    g = np.zeros((1, 6, 6))
    g[0, 0, 0] = 1          # isolated pixel
    g[0, 2:5, 2:5] = 1      # big component
    g[0, 0, 4:6] = 1        # small component
    out = clean_pred(g, 6).reshape(6, 6)
    assert out.sum() == 9 and out[0, 0] == 0 and out[0, 4] == 0


def test_cond_to_gpu_split():
    from videometamaterials_b200.trainer import num_to_groups
    assert num_to_groups(5, 2) == [2, 2, 1] and num_to_groups(4, 2) == [2, 2] and num_to_groups(0, 2) == []
