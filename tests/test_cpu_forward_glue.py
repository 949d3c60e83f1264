"""The product's forward GLUE on the CPU: `blocks.unet_forward` (packing, channels-last layouts, tap tables, the implicit concat
of the skip connections, the prepared-input layout, the rotary epilogue of the to_qkv projection, block wiring) run with every
kernel wrapper replaced by a torch statement of its contract (tests/emu_ops.py, tests/emu_cgemm.py) must reproduce the oracle's
forward.  This checks everything between the public API and the C ABI without a GPU; the kernels themselves are the `-m gpu` tests'
business.  The one product line that refuses CPU tensors is passed by a tensor subclass that claims to be on a CUDA device."""
import pytest
import torch

import emu_ops


class _ClaimsCuda(torch.Tensor):
    is_cuda = property(lambda self: True)


@pytest.mark.parametrize("size,b", [(16, 2), (12, 1)])
def test_unet_forward_glue_matches_oracle(monkeypatch, size, b):
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import Unet3D, blocks, ops
    emu_ops.install(monkeypatch, ops)
    cfg = O.UnetCfg(dim=16, dim_mults=(1, 2))
    sd = O.synthetic_state_dict(cfg, seed=13)
    model = Unet3D(dim=16, dim_mults=(1, 2), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
                   cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True)
    model.load_state_dict(sd)
    model.compute_dtype, model._packed = torch.float32, None          # fp32 "16-bit" operands: the glue is exact, tolerances tight
    g = torch.Generator().manual_seed(size)
    x = torch.randn(b, 3, 11, size, size, generator=g)
    cond = torch.rand(b, 11, generator=g) * 2 - 1
    t = torch.randint(0, 8, (b,), generator=g)
    mask = torch.tensor([False, True][:b])
    rel = lambda a, r: float((a - r).norm() / r.norm())
    for pre_rot in (True, False):                                      # rotary in the to_qkv epilogue, or inside the attention core
        monkeypatch.setattr(blocks, "ROTARY_IN_EPILOGUE", pre_rot)
        eps = blocks.unet_forward(model, x.as_subclass(_ClaimsCuda), None, None, t, cond, mask)      # (b, f, h, w, c)
        want = O.unet_forward(sd, cfg, x, t, cond, mask)
        assert tuple(eps.shape) == (b, 11, size, size, 3)
        assert rel(torch.Tensor(eps).permute(0, 4, 1, 2, 3), want) < 2e-5, pre_rot
    # q_sample fused into the input preparation (training form): x_t = a x0 + s noise
    noise = torch.randn(b, 3, 11, size, size, generator=g)
    a, s = torch.rand(b, generator=g), torch.rand(b, generator=g)
    eps = blocks.unet_forward(model, x.as_subclass(_ClaimsCuda), noise, (a, None, s), t, cond, mask)
    want = O.unet_forward(sd, cfg, a.view(-1, 1, 1, 1, 1) * x + s.view(-1, 1, 1, 1, 1) * noise, t, cond, mask)
    assert rel(torch.Tensor(eps).permute(0, 4, 1, 2, 3), want) < 2e-5
