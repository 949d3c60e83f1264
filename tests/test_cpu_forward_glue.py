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


@pytest.mark.parametrize("size,b,mode,tcond,c2t", [(16, 2, "zeros", True, "add"), (12, 1, "zeros", True, "add"), (16, 2, "circular", True, "add"),
                                                   (12, 2, "circular_1d", True, "add"), (12, 2, "zeros", False, "add"), (12, 2, "zeros", True, "concat"),
                                                   (12, 2, "circular", False, "concat")])
def test_unet_forward_glue_matches_oracle(monkeypatch, size, b, mode, tcond, c2t):
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import Unet3D, blocks, ops
    emu_ops.install(monkeypatch, ops)
    # the oracle's wrap modes and its use_temporal_attention_cond / cond_to_time branches are pinned to the reference (test_cpu_oracle.py)
    cfg = O.UnetCfg(dim=16, dim_mults=(1, 2), padding_mode=mode, temporal_cond=tcond, cond_to_time=c2t)
    sd = O.synthetic_state_dict(cfg, seed=13)
    model = Unet3D(dim=16, dim_mults=(1, 2), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
                   cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=tcond, cond_to_time=c2t, per_frame_cond=True,
                   padding_mode=mode)
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


@pytest.mark.parametrize("channels,mults,l2,size,mode,tcond,c2t", [(3, (1, 2), False, 12, "zeros", True, "add"), (4, (1, 2, 4), True, 16, "zeros", True, "add"),
                                                                  (1, (1,), False, 8, "zeros", True, "add"), (3, (1, 2), False, 12, "circular", True, "add"),
                                                                  (3, (1, 2, 4), True, 16, "circular_1d", True, "add"),
                                                                  (3, (1, 2), False, 12, "zeros", False, "add"), (3, (1, 2), True, 12, "zeros", True, "concat"),
                                                                  (3, (1, 2), False, 12, "zeros", False, "concat")])
def test_training_glue_matches_oracle_gradients(monkeypatch, channels, mults, l2, size, mode, tcond, c2t):
    """The same for the training form: `blocks_bwd.training_loss` + backward (block-level autograd Functions, data gradients
    through the transformed weight packs with the fused concat split, weight gradients scattered into the master layout, bias
    gradients folded into the GroupNorm backward, the batched conditioning path) must give the oracle's loss and every parameter
    gradient.  fp32 throughout, so the bound is tight; the L1 loss back-propagates signs, which fp32 reproduces.  Cases: the shipped
    shape family, four channels on three levels with the L2 loss, and a single-channel single-level network."""
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import Unet3D, blocks, blocks_bwd, ops
    emu_ops.install_training(monkeypatch, ops)
    cfg = O.UnetCfg(dim=16, dim_mults=mults, channels=channels, padding_mode=mode, temporal_cond=tcond, cond_to_time=c2t)
    sd = O.synthetic_state_dict(cfg, seed=14)
    model = Unet3D(dim=16, dim_mults=mults, channels=channels, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
                   cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=tcond, cond_to_time=c2t, per_frame_cond=True,
                   padding_mode=mode)
    model.load_state_dict(sd)
    model.compute_dtype, model._packed = torch.float32, None
    g = torch.Generator().manual_seed(3)
    b = 2
    x01 = torch.rand(b, channels, 11, size, size, generator=g)
    cond = torch.rand(b, 11, generator=g) * 2 - 1
    t = torch.tensor([2, 6])
    noise = torch.randn(b, channels, 11, size, size, generator=g)
    mask = torch.tensor([False, True])
    S = O.schedule(8)
    P = {k: v.clone().requires_grad_(v.is_floating_point() and "freqs" not in k) for k, v in sd.items()}
    want = O.p_losses(P, cfg, S, x01, t, cond, noise, mask, loss_type="l2" if l2 else "l1")
    want.backward()
    for pre_rot in (True, False):
        monkeypatch.setattr(blocks, "ROTARY_IN_EPILOGUE", pre_rot)
        arena = blocks_bwd.get_arena(model)
        arena.zero_grad()
        a, s = S["sqrt_alphas_cumprod"][t].contiguous(), S["sqrt_one_minus_alphas_cumprod"][t].contiguous()
        loss = blocks_bwd.training_loss(model, (x01 * 2 - 1).as_subclass(_ClaimsCuda), noise, (a, None, s), t, cond, mask, l2=l2)
        loss.backward()
        assert abs(float(loss.detach()) - float(want.detach())) < 1e-5 * float(want.detach())
        worst = (0.0, None)
        for k, p in model.named_parameters():
            ref = P[k].grad
            if ref is None or float(ref.norm()) == 0.0:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, k       # the never-used parameters keep a zero gradient
                continue
            e = float((p.grad - ref).norm() / ref.norm())
            worst = max(worst, (e, k))
        assert worst[0] < 1e-4, (pre_rot, worst)            # measured 3e-6 .. 4e-6 over 126 / 196 / 266 parameter tensors


@pytest.mark.parametrize("ln_form", [True, False])
def test_training_glue_through_the_fused_qkv_backward(monkeypatch, ln_form):
    """The 64-channel branch of AttnBlockFn.backward (one fused launch for the to_qkv data gradient + weight gradient, with or without the
    PreNorm backward in its epilogue, csrc/qkvbwd.cu) wired with the right operands: a dim-64 network on the CPU contract statements, the
    row-count threshold of the branch lowered, against the oracle's loss and every parameter gradient."""
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import Unet3D, blocks, blocks_bwd, ops
    emu_ops.install_training(monkeypatch, ops)
    monkeypatch.setattr(ops, "qkv_bwd_eligible", lambda dq, xn: dq.shape[1] == 768 and xn.shape[1] == 64)
    monkeypatch.setattr(ops, "FUSED_QKV_LN_BWD", ln_form)
    used = []
    for name in ("qkv_bwd", "qkv_ln_bwd"):
        fn = getattr(ops, name)
        monkeypatch.setattr(ops, name, (lambda f, n: (lambda *a, **k: (used.append(n), f(*a, **k))[1]))(fn, name))
    cfg = O.UnetCfg(dim=64, dim_mults=(1,))
    sd = O.synthetic_state_dict(cfg, seed=15)
    model = Unet3D(dim=64, dim_mults=(1,), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
                   cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True)
    model.load_state_dict(sd)
    model.compute_dtype, model._packed = torch.float32, None
    g = torch.Generator().manual_seed(4)
    b, size = 2, 8
    x01 = torch.rand(b, 3, 11, size, size, generator=g)
    cond = torch.rand(b, 11, generator=g) * 2 - 1
    t = torch.tensor([2, 6])
    noise = torch.randn(b, 3, 11, size, size, generator=g)
    mask = torch.tensor([False, True])
    S = O.schedule(8)
    P = {k: v.clone().requires_grad_(v.is_floating_point() and "freqs" not in k) for k, v in sd.items()}
    want = O.p_losses(P, cfg, S, x01, t, cond, noise, mask)
    want.backward()
    blocks_bwd.get_arena(model).zero_grad()
    a, s = S["sqrt_alphas_cumprod"][t].contiguous(), S["sqrt_one_minus_alphas_cumprod"][t].contiguous()
    loss = blocks_bwd.training_loss(model, (x01 * 2 - 1).as_subclass(_ClaimsCuda), noise, (a, None, s), t, cond, mask)
    loss.backward()
    assert used and set(used) == {"qkv_ln_bwd" if ln_form else "qkv_bwd"} and len(used) == 7      # init, downs.0.{2,3}, mid spatial + temporal, ups.0.{2,3}: every attention block of a one-level dim-64 network
    assert abs(float(loss.detach()) - float(want.detach())) < 1e-5 * float(want.detach())
    worst = (0.0, None)
    for k, p in model.named_parameters():
        ref = P[k].grad
        if ref is None or float(ref.norm()) == 0.0:
            continue
        worst = max(worst, (float((p.grad - ref).norm() / ref.norm()), k))
    assert worst[0] < 1e-4, worst


class _Replay:
    """torch.randn / randn_like return a recorded list of tensors, in call order (as the GPU sampling tests do)."""

    def __init__(self, tensors):
        self.tensors, self.i = list(tensors), 0

    def __enter__(self):
        self._a, self._b = torch.randn, torch.randn_like

        def nxt(*a, **k):
            t = self.tensors[self.i]
            self.i += 1
            return t.clone().as_subclass(_ClaimsCuda)

        torch.randn, torch.randn_like = nxt, nxt
        return self

    def __exit__(self, *e):
        torch.randn, torch.randn_like = self._a, self._b


def test_sampler_glue_matches_oracle(monkeypatch, golden_dir):
    """GaussianDiffusion.p_sample / sample (ancestral with dynamic thresholding, guided and unguided) and ddim_sample, eager form
    and the device-coefficient DDIM step that the CUDA graph replays (`_ddim_step_core`), against the oracle on recorded noise."""
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import GaussianDiffusion, Unet3D, ops
    emu_ops.install_sampler(monkeypatch, ops)
    cfg = O.UnetCfg(dim=16, dim_mults=(1, 2))
    sd = O.synthetic_state_dict(cfg, seed=15)
    rel = lambda a, r: float((torch.Tensor(a) - r).norm() / r.norm())

    def make(sampling_T):
        model = Unet3D(dim=16, dim_mults=(1, 2), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
                       cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True)
        model.load_state_dict(sd)
        model.compute_dtype, model._packed = torch.float32, None
        return GaussianDiffusion(model, image_size=12, channels=3, num_frames=11, timesteps=8, loss_type='l1', use_dynamic_thres=True,
                                 sampling_timesteps=sampling_T)

    g = torch.Generator().manual_seed(4)
    b = 2
    cond = torch.rand(b, 11, generator=g) * 2 - 1
    noises = [torch.randn(b, 3, 11, 12, 12, generator=g) for _ in range(9)]
    S = O.schedule(8)
    gd = make(8)
    for w in (5.0, 1.0):
        with _Replay(noises):
            got = gd.sample(cond=cond, guidance_scale=w)
        want = O.p_sample_loop(sd, cfg, S, noises[0], cond, w, noises[1:])
        assert rel(got, want) < 1e-4, w
    gd4 = make(4)
    with _Replay(noises):
        got = gd4.sample(cond=cond, guidance_scale=5.0)
    want = O.ddim_sample(sd, cfg, S, noises[0], cond, 5.0, 4)
    assert rel(got, want) < 1e-4
    # the graph-replayed DDIM step: per-step scalars gathered from the shifted alphas_cumprod table on the device
    img = noises[0].clone().as_subclass(_ClaimsCuda)
    st = dict(x=img, t=torch.zeros(b, dtype=torch.long), tn=torch.zeros(b, dtype=torch.long), cond=cond, x0=torch.empty(b, 3, 11, 12, 12),
              eps=torch.empty(b, 3, 11, 12, 12), zeros=torch.zeros(b), acp_next=torch.cat((torch.ones(1), gd4.alphas_cumprod)).contiguous())
    for time, time_next in gd4._ddim_pairs():
        st["t"].fill_(time)
        st["tn"].fill_(time_next)
        st["x"] = torch.Tensor(gd4._ddim_step_core(st, 5.0)).as_subclass(_ClaimsCuda)
    assert rel((st["x"] + 1) * 0.5, want) < 1e-4
    # eta > 0 (ddim_sampling_eta, VDDP:1006-1016): the stochastic DDIM update through vmm_posterior_step, eager and graph-step form
    gde = make(4)
    gde.ddim_sampling_eta = 0.5
    with _Replay(noises):
        got = gde.sample(cond=cond, guidance_scale=5.0)
    want = O.ddim_sample(sd, cfg, S, noises[0], cond, 5.0, 4, eta=0.5, noises=noises[1:])
    assert rel(got, want) < 1e-4
    st.update(x=noises[0].clone().as_subclass(_ClaimsCuda), noise=torch.zeros(b, 3, 11, 12, 12))
    k = 1
    for time, time_next in gde._ddim_pairs():
        st["t"].fill_(time)
        st["tn"].fill_(time_next)
        if time_next >= 0:
            st["noise"].copy_(noises[k])
            k += 1
        st["x"] = torch.Tensor(gde._ddim_step_core(st, 5.0)).as_subclass(_ClaimsCuda)
    assert rel((st["x"] + 1) * 0.5, want) < 1e-4
    assert float((torch.Tensor(got) - O.ddim_sample(sd, cfg, S, noises[0], cond, 5.0, 4)).abs().max()) > 1e-3      # eta changes the sample


def test_trainer_steps_reduce_the_loss(monkeypatch, tmp_path):
    """The whole optimisation step on the CPU: Trainer.train_step -> p_losses -> block Functions -> gradients in the arena ->
    fused Adam (+EMA) -> weight repack -> next forward, with the kernels replaced by their contract statements and the optimiser
    kernel by the oracle's.  On a fixed batch with a fixed (t, noise, mask) draw the loss must go down step after step."""
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import Accelerator, GaussianDiffusion, Trainer, Unet3D, ops
    emu_ops.install_training(monkeypatch, ops)
    monkeypatch.setattr(ops, "adam_ema_step", lambda p, g, m, v, ema, lr, b1, b2, eps, step, gs, mode, beta:
                        O.adam_ema_step(p, g, m, v, ema, step, lr=lr, beta1=b1, beta2=b2, eps=eps, grad_scale=gs, ema_mode=mode, ema_beta=beta))
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(0)
    m = Unet3D(dim=16, dim_mults=(1, 2), channels=3, cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, per_frame_cond=True)
    gd = GaussianDiffusion(m, image_size=8, channels=3, num_frames=11, timesteps=8, use_dynamic_thres=True, sampling_timesteps=8)
    t = Trainer(gd, None, None, [0, 1, 3], train_batch_size=2, test_batch_size=4, train_lr=3e-3, results_folder='run', log=True,
                null_cond_prob=0.1, per_frame_cond=True, reference_frame='lagrangian', accelerator=Accelerator(cpu=True),
                update_ema_every=2, step_start_ema=4)
    t.use_cuda_graph = False                                   # the input below only claims to be a CUDA tensor
    m.compute_dtype, m._packed = torch.float32, None
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 3, 11, 8, 8, generator=g).as_subclass(_ClaimsCuda)
    cond = torch.rand(2, 11, generator=g) * 2 - 1
    w0 = m.init_conv.weight.detach().clone()
    losses = []
    for i in range(8):
        torch.manual_seed(100)                                 # the same t / noise / conditioning-drop draw every step
        t.step = i
        losses.append(float(t.train_step(x, cond)))
    assert all(b < a for a, b in zip(losses, losses[1:])), losses
    assert losses[-1] < 0.8 * losses[0]
    assert not torch.equal(m.init_conv.weight.detach(), w0)
    ema_w = t.ema_model.denoise_fn.init_conv.weight.detach()
    assert not torch.equal(ema_w, w0) and not torch.equal(ema_w, m.init_conv.weight.detach())      # copied at steps 0, 2; averaged at 4, 6


def test_interpolate_glue(monkeypatch):
    """GaussianDiffusion.interpolate (VDDP:1020-1034) with the conditioning the reference forgets to pass: two q_sample draws at step
    t, a blend, then p_sample from t - 1 to 0, against the same composition of the oracle's pieces on the recorded noise."""
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import GaussianDiffusion, Unet3D, ops
    emu_ops.install_sampler(monkeypatch, ops)
    cfg = O.UnetCfg(dim=16, dim_mults=(1, 2))
    sd = O.synthetic_state_dict(cfg, seed=16)
    model = Unet3D(dim=16, dim_mults=(1, 2), channels=3, cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True,
                   per_frame_cond=True)
    model.load_state_dict(sd)
    model.compute_dtype, model._packed = torch.float32, None
    gd = GaussianDiffusion(model, image_size=12, channels=3, num_frames=11, timesteps=8, use_dynamic_thres=True, sampling_timesteps=8)
    g = torch.Generator().manual_seed(6)
    b, t, lam = 2, 3, 0.3
    x1, x2 = (torch.rand(b, 3, 11, 12, 12, generator=g) * 2 - 1 for _ in range(2))
    cond = torch.rand(b, 11, generator=g) * 2 - 1
    noises = [torch.randn(b, 3, 11, 12, 12, generator=g) for _ in range(2 + t)]
    with _Replay(noises):
        got = gd.interpolate(x1.as_subclass(_ClaimsCuda), x2.as_subclass(_ClaimsCuda), t=t, lam=lam, cond=cond, guidance_scale=5.0)
    S = O.schedule(8)
    tb = torch.full((b,), t)
    img = (1 - lam) * O.q_sample(S, x1, tb, noises[0]) + lam * O.q_sample(S, x2, tb, noises[1])
    for k, i in enumerate(reversed(range(t))):
        img = O.p_sample(sd, cfg, S, img, torch.full((b,), i), cond, 5.0, noises[2 + k])
    assert float((torch.Tensor(got) - img).norm() / img.norm()) < 1e-4
    with pytest.raises(ValueError):
        gd.interpolate(x1, x2, t=t)
