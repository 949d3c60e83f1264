"""Whole-network parity on the GPU against the committed golden fixtures (reference outputs, fp32 CPU).

Tolerances (relative L2, stated per SURVEY.md section 0 D5): the network stores activations in 16 bit, so one
forward is expected at ~1e-3 (fp16) / ~6e-3 (bf16) of the fp32 reference; guided outputs amplify the
difference of two forwards by w=5.  The post-network sampler arithmetic is fp32 and adds nothing measurable.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

FWD_TOL = {torch.float16: 3e-3, torch.bfloat16: 2e-2}


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp_min(1e-30))


def build(cfg_dim, mults, T, size, sampling_T, dtype, seed):
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import GaussianDiffusion, Unet3D
    model = Unet3D(dim=cfg_dim, dim_mults=mults, channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True,
                   resnet_groups=8, cond_bias=True, cond_attention='self-stacked', cond_attention_tokens=16,
                   use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True, padding_mode='zeros')
    sd = O.synthetic_state_dict(O.UnetCfg(dim=cfg_dim, dim_mults=mults), seed=seed)
    model.load_state_dict(sd, strict=True)
    model.set_compute_dtype(dtype)
    gd = GaussianDiffusion(model, image_size=size, channels=3, num_frames=11, timesteps=T, loss_type='l1', use_dynamic_thres=True,
                           sampling_timesteps=sampling_T).cuda()
    return model, gd, sd


class Replay:
    def __init__(self, tensors):
        self.tensors, self.i = list(tensors), 0

    def __enter__(self):
        self._a, self._b = torch.randn, torch.randn_like

        def nxt(*a, **k):
            t = self.tensors[self.i]
            self.i += 1
            return t.clone().cuda()

        torch.randn, torch.randn_like = nxt, nxt
        return self

    def __exit__(self, *e):
        torch.randn, torch.randn_like = self._a, self._b


@pytest.fixture(scope="module")
def gold_small(golden_dir):
    return torch.load(os.path.join(golden_dir, "small_unet.pt"))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_small_unet_forward(gold_small, dtype):
    g = gold_small
    model, gd, _ = build(16, (1, 2), g["T"], g["size"], g["T"], dtype, g["seed"])
    x, t, cond = g["x"].cuda(), g["t"].cuda(), g["cond"].cuda()
    with torch.no_grad():
        y_cond = model(x, t, cond=cond, null_cond_prob=0.0)
        y_null = model(x, t, cond=cond, null_cond_prob=1.0)
        y_g = model.forward_with_guidance_scale(x, t, cond=cond, guidance_scale=5.0)
    e = [rel(y_cond, g["y_cond"]), rel(y_null, g["y_null"]), rel(y_g, g["y_guided"])]
    print("small forward rel-L2 (cond, null, guided):", dtype, e)
    assert e[0] < FWD_TOL[dtype] and e[1] < FWD_TOL[dtype] and e[2] < 5 * FWD_TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float16])
def test_small_sampling(gold_small, dtype):
    g = gold_small
    model, gd, _ = build(16, (1, 2), g["T"], g["size"], g["T"], dtype, g["seed"])
    x, t, cond = g["x"].cuda(), g["t"].cuda(), g["cond"].cuda()
    noises = list(g["step_noise"])
    with Replay([noises[0]]):
        p1 = gd.p_sample(x, t, cond=cond, guidance_scale=5.0)
    with Replay([noises[0]]):
        p0 = gd.p_sample(x, torch.zeros_like(t), cond=cond, guidance_scale=5.0)
    with Replay(noises):
        loop = gd.sample(cond=cond, guidance_scale=5.0)
    _, gd4, _ = build(16, (1, 2), g["T"], g["size"], 4, dtype, g["seed"])
    with Replay(noises):
        ddim = gd4.sample(cond=cond, guidance_scale=5.0)
    e = dict(p_sample_t=rel(p1, g["p_sample_t"]), p_sample_t0=rel(p0, g["p_sample_t0"]), loop=rel(loop, g["loop"]),
             ddim4=rel(ddim, g["ddim4"]))
    print("small sampling rel-L2:", e)
    assert e["p_sample_t"] < 1e-2 and e["p_sample_t0"] < 1e-2
    assert e["loop"] < 5e-2 and e["ddim4"] < 5e-2


@pytest.mark.parametrize("sampling_T", [8, 4])
def test_graph_replayed_sampling_matches_eager(gold_small, sampling_T):
    """GaussianDiffusion.use_cuda_graph replays one captured step (ancestral: `_p_sample_core`; DDIM: `_ddim_step_core`, whose
    per-step scalars are read on the device) instead of launching ~600 kernels per step from Python.  From the same torch seed it
    must reproduce the eager loop, which the test above pins to the reference: same Philox consumption (randn for x_T, one
    normal_() per step), same kernels; differences come only from the order of floating-point atomics (bound 1e-2 after 8 steps; the eager loop itself is 2e-3 from the reference)."""
    g = gold_small
    _, gd, _ = build(16, (1, 2), g["T"], g["size"], sampling_T, torch.float16, g["seed"])
    assert gd.is_ddim_sampling == (sampling_T < g["T"])
    cond = g["cond"].cuda()
    for w in (5.0, 1.0):
        gd.use_cuda_graph = False
        torch.manual_seed(21)
        eager = gd.sample(cond=cond, guidance_scale=w)
        gd.use_cuda_graph = True
        torch.manual_seed(21)
        first = gd.sample(cond=cond, guidance_scale=w)          # captures, then replays
        torch.manual_seed(21)
        again = gd.sample(cond=cond, guidance_scale=w)          # replays only
        torch.cuda.synchronize()
        e = (rel(first, eager), rel(again, eager))
        print("graph vs eager sampling rel-L2:", sampling_T, w, e)
        assert e[0] < 1e-2 and e[1] < 1e-2
        assert float(eager.min()) >= 0.0 or gd.is_ddim_sampling    # ancestral x0 is thresholded into [-1, 1] -> [0, 1] after unnormalize
    # other conditionings through the same graphs (static buffers are refreshed)
    n_graphs = len(gd._graphs)
    torch.manual_seed(22)
    other = gd.sample(cond=-cond, guidance_scale=5.0)
    assert len(gd._graphs) == n_graphs and rel(other, again) > 1e-3
    # a repack (one gather launch over the parameter arena) rewrites the SAME packed-operand buffer: the captured graph stays valid ...
    gd.denoise_fn._packed = None
    torch.manual_seed(21)
    gd.use_cuda_graph = True
    again2 = gd.sample(cond=cond, guidance_scale=1.0)
    assert len(gd._graphs) == n_graphs and rel(again2, again) < 1e-2
    # ... and replays with the NEW weights after a parameter update (no stale operands baked into the graph)
    with torch.no_grad():
        gd.denoise_fn.final_conv[1].weight.mul_(2.0)
    torch.manual_seed(21)
    again3 = gd.sample(cond=cond, guidance_scale=1.0)
    assert len(gd._graphs) == n_graphs and rel(again3, again) > 1e-2


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_full_unet_forward_slices(golden_dir, dtype):
    """Shipped 96x96x11 configuration, b=1, t=128: compare against the strided slices of the reference output."""
    g = torch.load(os.path.join(golden_dir, "full_unet_slices.pt"))
    model, gd, _ = build(64, (1, 2, 4, 8), 256, 96, 256, dtype, g["seed"])
    gen = torch.Generator().manual_seed(g["data_seed"])
    x = torch.randn(1, 3, 11, 96, 96, generator=gen)
    cond = torch.rand(1, 11, generator=gen) * 2 - 1
    noise = torch.randn(1, 3, 11, 96, 96, generator=gen)
    t = torch.tensor([g["t"]])
    sl = (slice(None), slice(None), slice(None, None, 2), slice(None, None, 8), slice(None, None, 8))
    with torch.no_grad():
        yc = model(x.cuda(), t.cuda(), cond=cond.cuda(), null_cond_prob=0.0)
        yn = model(x.cuda(), t.cuda(), cond=cond.cuda(), null_cond_prob=1.0)
    with Replay([noise]):
        ps = gd.p_sample(x.cuda(), t.cuda(), cond=cond.cuda(), guidance_scale=5.0)
    e = dict(cond=rel(yc[sl], g["y_cond_slice"]), null=rel(yn[sl], g["y_null_slice"]), p_sample=rel(ps[sl], g["p_sample_slice"]),
             norm=abs(float(yc.norm()) - g["y_cond_norm"]) / g["y_cond_norm"])
    print("full forward rel-L2:", dtype, e)
    assert e["cond"] < FWD_TOL[dtype] and e["null"] < FWD_TOL[dtype] and e["norm"] < FWD_TOL[dtype]
    assert e["p_sample"] < 10 * FWD_TOL[dtype]


@pytest.mark.parametrize("size,b,check_grads", [(20, 3, True), (24, 1, True), (36, 2, False)])
def test_ragged_sizes_against_the_oracle(size, b, check_grads):
    """Resolutions that are not multiples of the 8 x 16 conv tile (20 -> 10, 36 -> 18 after the downsample) and odd batch sizes:
    forward, guided forward, loss and gradient norms against the CPU oracle computed here on the same seeded inputs.
    Measured on B200: forward 1.6e-3, guided 2.6e-3 at all three sizes; worst gradient-norm deviation 1.5 % (20, b=3) and 3.1 %
    (24, b=1).  At 36 x 36 the gradient norms are not compared in fp16: the linear-attention parameters there have true norms of
    5e-8 .. 5e-7 (v / (h w), VDDP:371), which a loss scale of 4096 leaves in fp16's subnormal range (observed: 36 % off on
    ups.1.2 to_qkv, everything else inside the bound).  Training runs in bf16, whose exponent range has no such floor."""
    from oracle import vdm_oracle as O
    cfg = O.UnetCfg(dim=16, dim_mults=(1, 2))
    model, gd, sd = build(16, (1, 2), 8, size, 8, torch.float16, seed=11)
    g = torch.Generator().manual_seed(size * 10 + b)
    x = torch.randn(b, 3, 11, size, size, generator=g)
    cond = torch.rand(b, 11, generator=g) * 2 - 1
    t = torch.randint(0, 8, (b,), generator=g)
    noise = torch.randn(b, 3, 11, size, size, generator=g)
    with torch.no_grad():
        y = model(x.cuda(), t.cuda(), cond=cond.cuda(), null_cond_prob=0.0)
        yg = model.forward_with_guidance_scale(x.cuda(), t.cuda(), cond=cond.cuda(), guidance_scale=3.0)
        y_ref = O.unet_forward(sd, cfg, x, t, cond, torch.zeros(b, dtype=torch.bool))
        yg_ref = O.unet_forward_guided(sd, cfg, x, t, cond, 3.0)
    e = (rel(y, y_ref), rel(yg, yg_ref))
    print("ragged forward rel-L2:", size, b, e)
    assert e[0] < FWD_TOL[torch.float16] and e[1] < 5 * FWD_TOL[torch.float16]
    # loss + gradients (fp16 needs the loss scale)
    x01 = torch.rand(b, 3, 11, size, size, generator=g)
    P = {k: v.clone().requires_grad_(v.is_floating_point() and "freqs" not in k) for k, v in sd.items()}
    loss_ref = O.p_losses(P, cfg, O.schedule(8), x01, t, cond, noise, torch.zeros(b, dtype=torch.bool))
    loss_ref.backward()
    loss = gd.p_losses((x01 * 2 - 1).cuda(), t.cuda(), cond=cond.cuda(), noise=noise.cuda(), null_cond_prob=0.0)
    (loss * 4096.0).backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(loss_ref)) / float(loss_ref) < 2e-3
    if not check_grads:
        return
    params = dict(model.named_parameters())
    # relative deviation of every gradient norm, with an absolute floor of 1e-7: the linear-attention PreNorm gains have true
    # gradient norms of ~5e-8 at 36 x 36 (v / (h w), VDDP:371), four orders below the other tensors (~5e-4) and inside the 16-bit
    # noise floor, so their relative error is meaningless
    worst = max((abs(float(params[k].grad.norm()) / 4096.0 - float(p.grad.norm())) / (float(p.grad.norm()) + 1e-7 / 0.08), k)
                for k, p in P.items() if p.requires_grad and p.grad is not None and float(p.grad.norm()) > 0)
    print("ragged worst grad-norm deviation:", size, b, worst)
    assert worst[0] < 0.08, worst            # same bound family as test_small_training_loss_and_gradients (sign flips of the L1 loss)


@pytest.mark.parametrize("frames,b", [(22, 2), (7, 1)])
def test_other_frame_counts_against_the_oracle(frames, b):
    """BASELINE configs[4] names 22 frames.  The reference's shipped configuration cannot run there (11 hard-coded tokens, VDDP:603 /
    VDDP:777), so this package's `num_frames` extension sets tokens == frames and the temporal attention runs on the
    generic kernels (csrc/tattn_generic.cu).  No reference parity exists for these rows: the check is against this repo's oracle,
    which is pinned to the reference at 11 frames and generic in the frame count.  Forward, guided forward, loss, gradient norms."""
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import GaussianDiffusion, Unet3D
    size = 16
    cfg = O.UnetCfg(dim=16, dim_mults=(1, 2), frames=frames)
    model = Unet3D(dim=16, dim_mults=(1, 2), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
                   cond_bias=True, cond_attention='self-stacked', cond_attention_tokens=16, use_temporal_attention_cond=True,
                   cond_to_time='add', per_frame_cond=True, padding_mode='zeros', num_frames=frames)
    sd = O.synthetic_state_dict(cfg, seed=5)
    model.load_state_dict(sd, strict=True)
    model.set_compute_dtype(torch.float16)
    gd = GaussianDiffusion(model, image_size=size, channels=3, num_frames=frames, timesteps=8, loss_type='l1', use_dynamic_thres=True,
                           sampling_timesteps=8).cuda()
    g = torch.Generator().manual_seed(frames * 10 + b)
    x = torch.randn(b, 3, frames, size, size, generator=g)
    cond = torch.rand(b, frames, generator=g) * 2 - 1
    t = torch.randint(0, 8, (b,), generator=g)
    noise = torch.randn(b, 3, frames, size, size, generator=g)
    with torch.no_grad():
        y = model(x.cuda(), t.cuda(), cond=cond.cuda(), null_cond_prob=0.0)
        yg = model.forward_with_guidance_scale(x.cuda(), t.cuda(), cond=cond.cuda(), guidance_scale=3.0)
        y_ref = O.unet_forward(sd, cfg, x, t, cond, torch.zeros(b, dtype=torch.bool))
        yg_ref = O.unet_forward_guided(sd, cfg, x, t, cond, 3.0)
    e = (rel(y, y_ref), rel(yg, yg_ref))
    print("frames forward rel-L2:", frames, b, e)
    assert e[0] < FWD_TOL[torch.float16] and e[1] < 5 * FWD_TOL[torch.float16]
    x01 = torch.rand(b, 3, frames, size, size, generator=g)
    P = {k: v.clone().requires_grad_(v.is_floating_point() and "freqs" not in k) for k, v in sd.items()}
    loss_ref = O.p_losses(P, cfg, O.schedule(8), x01, t, cond, noise, torch.zeros(b, dtype=torch.bool))
    loss_ref.backward()
    loss = gd.p_losses((x01 * 2 - 1).cuda(), t.cuda(), cond=cond.cuda(), noise=noise.cuda(), null_cond_prob=0.0)
    (loss * 4096.0).backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(loss_ref)) / float(loss_ref) < 2e-3
    params = dict(model.named_parameters())
    worst = max((abs(float(params[k].grad.norm()) / 4096.0 - float(p.grad.norm())) / (float(p.grad.norm()) + 1e-7 / 0.08), k)
                for k, p in P.items() if p.requires_grad and p.grad is not None and float(p.grad.norm()) > 0)
    print("frames worst grad-norm deviation:", frames, b, worst)
    assert worst[0] < 0.08, worst
    # one guided sampling step through the public sampler
    with torch.no_grad():
        img = gd.p_sample(x.cuda(), t.cuda(), cond=cond.cuda(), guidance_scale=3.0)
    assert tuple(img.shape) == (b, 3, frames, size, size) and bool(torch.isfinite(img).all())


def test_focus_present_mask_like_the_reference(gold_small):
    """N4: an all-False focus_present_mask is inert (VDDP:514); any True entry raises, as the unmodified reference does in the
    shipped configuration (its (n, n) self-mask meets the (n, 2n) similarity of stacked cond + frame keys, VDDP:514-524)."""
    g = gold_small
    model, gd, _ = build(16, (1, 2), g["T"], g["size"], g["T"], torch.float16, g["seed"])
    x, t, cond = g["x"].cuda(), g["t"].cuda(), g["cond"].cuda()
    with torch.no_grad():
        y0 = model(x, t, cond=cond, null_cond_prob=0.0)
        y1 = model(x, t, cond=cond, null_cond_prob=0.0, focus_present_mask=torch.zeros(x.shape[0], dtype=torch.bool, device="cuda"))
    assert rel(y1, y0) < 1e-3          # same path twice (fp32 atomics in the linear-attention context sums are not order-stable)
    with pytest.raises(RuntimeError, match="focus_present_mask"):
        model(x, t, cond=cond, null_cond_prob=0.0, focus_present_mask=torch.ones(x.shape[0], dtype=torch.bool, device="cuda"))
    with pytest.raises(RuntimeError, match="focus_present_mask"):
        model(x, t, cond=cond, null_cond_prob=0.0, prob_focus_present=0.5)


def test_forward_needs_cuda():
    from videometamaterials_b200 import Unet3D
    m = Unet3D(dim=16, dim_mults=(1, 2), per_frame_cond=True, use_temporal_attention_cond=True, cond_attention='self-stacked')
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 11, 16, 16), torch.zeros(1, dtype=torch.long), cond=torch.zeros(1, 11))


@pytest.mark.parametrize("dtype,scale", [(torch.bfloat16, 1.0), (torch.float16, 4096.0)])
def test_small_training_loss_and_gradients(gold_small, dtype, scale):
    """loss + every parameter gradient of p_losses().backward() against the reference's (fp32 CPU autograd).
    fp16 needs a loss scale (the reference trains under a GradScaler); bf16 does not."""
    g = gold_small
    model, gd, _ = build(16, (1, 2), g["T"], g["size"], g["T"], dtype, g["seed"])
    x01, t, cond, noise = g["x01"].cuda(), g["t"].cuda(), g["cond"].cuda(), g["noise"].cuda()
    loss = gd.p_losses(x01 * 2 - 1, t, cond=cond, noise=noise, null_cond_prob=0.0)
    (loss * scale).backward()
    torch.cuda.synchronize()
    tol = 2e-3 if dtype == torch.float16 else 1e-2
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < tol, (float(loss), float(g["loss"]))
    params = dict(model.named_parameters())
    worst = []
    for k, want in g["grad_norms"].items():
        got = float(params[k].grad.norm()) / scale
        worst.append((abs(got - want) / max(want, 1e-12), k, got, want))
    worst.sort(reverse=True)
    print("worst grad-norm deviations:", dtype, [(round(w, 4), k) for w, k, _, _ in worst[:6]])
    gtol = 0.05 if dtype == torch.float16 else 0.15
    assert worst[0][0] < gtol, worst[:5]
    for k in g["grad_none"]:
        assert float(params[k].grad.abs().max()) == 0.0, k       # never-used parameters keep a zero gradient
    errs = {}
    for k, want in g["grad_samples"].items():
        got = params[k].grad.flatten()[:64].float().cpu() / scale
        errs[k] = float((got - want).norm() / want.norm().clamp_min(1e-30))
    print("grad sample rel-L2:", dtype, {k: round(v, 4) for k, v in errs.items()})
    # 64-element slices: bf16 (8-bit mantissa activations and gradients) is ~8x noisier than fp16 on such a small sample.
    # The L1 loss back-propagates sign(pred - noise): a handful of sign flips from 16-bit rounding perturbs EVERY gradient, so
    # the whole-tensor error against the fp32 oracle moves between 0.9 % and 2.5 % (median over parameters) from one input /
    # rounding pattern to the next (tests/probe_grad_accuracy.py, three seeds, both rotary placements); slices reach 6 %.
    assert max(errs.values()) < (0.08 if dtype == torch.float16 else 0.5), errs


def test_gather_repack_equals_slicing_pack(gold_small):
    """After a training step the packed GEMM operands come from the one-launch gather; they must equal the reference
    slicing pack of the updated parameters bit for bit."""
    from videometamaterials_b200 import blocks
    from videometamaterials_b200.blocks_bwd import get_arena
    g = gold_small
    model, gd, _ = build(16, (1, 2), g["T"], g["size"], g["T"], torch.bfloat16, g["seed"])
    get_arena(model)
    model.repack()
    assert getattr(model, "_pack_plan", None) is not None
    want = blocks.pack_all(model, torch.bfloat16)
    got = model.packed()
    assert set(got) == set(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k


def test_graph_replay_matches_eager_step():
    """Trainer._fwd_bwd replays forward + backward from a CUDA graph.  With the same Philox seed the replayed step must give
    the loss and the whole gradient arena of the eager step.  Two EAGER runs from one seed already differ by ~2e-3 relative
    L2 in bf16 (atomic summation order flips 16-bit roundings), so the bound is 1e-2 on the arena and 5e-2 per parameter."""
    from videometamaterials_b200 import Accelerator, Trainer
    from videometamaterials_b200.blocks_bwd import get_arena
    model, gd, _ = build(16, (1, 2), 8, 16, 8, torch.bfloat16, seed=3)
    tr = Trainer(gd, folder=None, validation_folder=None, selected_channels=[0, 1, 3], train_batch_size=2, results_folder="/tmp/vmm_graph_test",
                 log=False, null_cond_prob=0.1, per_frame_cond=True, reference_frame='lagrangian', accelerator=Accelerator("bf16"))
    g = torch.Generator().manual_seed(11)
    x = torch.rand(2, 3, 11, 16, 16, generator=g).cuda()
    cond = (torch.rand(2, 11, generator=g) * 2 - 1).cuda()
    arena = get_arena(tr.model.denoise_fn)
    tr.use_cuda_graph = False
    torch.cuda.manual_seed(123)
    loss_e = float(tr._fwd_bwd(x, cond).detach())
    grad_e = arena.flat_grad.clone()
    assert float(grad_e.norm()) > 0
    tr.use_cuda_graph, tr.graph_warmup = True, 0
    tr._fwd_bwd(x, cond)                        # captures, then replays once
    assert tr._graph_state["graph"] is not None and tr._graph_state["launches"] > 100
    for _ in range(2):                          # replays proper, each from the same seed / offset
        torch.cuda.manual_seed(123)
        loss_g = float(tr._fwd_bwd(x, cond))
        torch.cuda.synchronize()
        assert abs(loss_g - loss_e) < 1e-5 * max(1.0, abs(loss_e))
        assert rel(arena.flat_grad, grad_e) < 1e-2
        o = 0
        for i, p in enumerate(arena.params):
            k = p.numel()
            if float(grad_e[o:o + k].norm()) > 0:
                assert rel(arena.flat_grad[o:o + k], grad_e[o:o + k]) < 5e-2, (i, tuple(p.shape))
            o += k
    # a different batch through the same graph changes the result (the static input buffers are refreshed)
    torch.cuda.manual_seed(123)
    loss_2 = float(tr._fwd_bwd(x.flip(0) * 0.5, cond))
    assert abs(loss_2 - loss_e) > 1e-6


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_full_size_batch_properties(dtype):
    """BASELINE configs[1] size (b=8, 96x96x11): properties that need no oracle.  Tolerance = FWD_TOL, the bound of one
    forward against the fp32 reference: two runs of the same clip differ by the order of the fp32 / fp64 atomics (linear
    attention context, GroupNorm statistics), which flips 16-bit roundings that then propagate through ~60 layers.
    (1) batch invariance: every normalisation is per sample, so sample 5 of the batch equals the same clip run alone;
    (2) the conditioning drop: null_cond_prob = 1 output is independent of cond;
    (3) the fused guidance step at w = 1 short-circuits to the conditional forward (VDDP:719-721)."""
    model, gd, _ = build(64, (1, 2, 4, 8), 256, 96, 256, dtype, 0)
    tol = FWD_TOL[dtype]
    gen = torch.Generator().manual_seed(77)
    x = torch.randn(8, 3, 11, 96, 96, generator=gen).cuda()
    cond = (torch.rand(8, 11, generator=gen) * 2 - 1).cuda()
    t = torch.randint(0, 256, (8,), generator=gen).cuda()
    with torch.no_grad():
        y8 = model(x, t, cond=cond, null_cond_prob=0.0)
        y1 = model(x[5:6], t[5:6], cond=cond[5:6], null_cond_prob=0.0)
        assert torch.isfinite(y8).all()
        assert rel(y8[5:6], y1) < tol
        yn_a = model(x[:2], t[:2], cond=cond[:2], null_cond_prob=1.0)
        yn_b = model(x[:2], t[:2], cond=-cond[:2], null_cond_prob=1.0)
        assert rel(yn_a, yn_b) < tol
        yg = model.forward_with_guidance_scale(x[:2], t[:2], cond=cond[:2], guidance_scale=1.0)
        assert rel(yg, y8[:2]) < tol


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_full_size_conv_linearity(dt):
    """Level-0 3x3 conv at the full b=8 size (88 frame-images of 96x96, 64 -> 64, halo mode with resident weights):
    conv(x1 + x2) = conv(x1) + conv(x2) and conv(2 x) = 2 conv(x) without bias, to 16-bit rounding; a one-hot input
    reproduces the (flipped) kernel taps exactly."""
    from videometamaterials_b200 import ops
    torch.manual_seed(5)
    bf, H, W, C = 88, 96, 96, 64
    w = (torch.randn(C, C, 3, 3, device="cuda") / (9 * C) ** 0.5).to(dt)
    wp = ops.pack_conv_taps(w.float(), [C], dt)
    x1 = torch.randn(bf, H, W, C, device="cuda").to(dt)
    x2 = torch.randn(bf, H, W, C, device="cuda").to(dt)
    outs = []
    for xin in (x1, x2, (x1.float() + x2.float()).to(dt), (2 * x1.float()).to(dt)):
        o = torch.empty(bf, H, W, C, device="cuda", dtype=dt)
        ops.conv3x3([xin], wp, C, o)
        outs.append(o.float())
    tol = 1.2e-2 if dt == torch.bfloat16 else 1.5e-3          # three roundings: the summed input and two outputs
    assert rel(outs[2], outs[0] + outs[1]) < tol
    assert rel(outs[3], 2 * outs[0]) < 1e-6                   # scaling by two is exact in binary floating point
    # impulse response: x = delta at (frame 3, y 40, x 50, channel 7) -> out[3, 40 + dy, 50 + dx, n] = w[n, 7, 1 - dy, 1 - dx]
    imp = torch.zeros(bf, H, W, C, device="cuda", dtype=dt)
    imp[3, 40, 50, 7] = 1.0
    o = torch.empty(bf, H, W, C, device="cuda", dtype=dt)
    ops.conv3x3([imp], wp, C, o)
    want = w[:, 7].float().flip(1, 2).permute(1, 2, 0)       # (dy, dx, n)
    assert torch.equal(o[3, 39:42, 49:52].float(), want)
    o[3, 39:42, 49:52] = 0
    assert float(o.float().abs().max()) == 0.0


def test_ddim_eta_graph_matches_eager_and_golden(golden_dir):
    """`ddim_sampling_eta` > 0 (VDDP:1006-1016): the stochastic DDIM update runs through vmm_posterior_step (c1 x0 + c2 eps +
    sigma noise, no clamp).  Eager loop on the recorded noise against the reference's sample (tests/golden/ddim_eta.pt), and the
    graph-replayed loop against the eager one from the same torch seed.  (Added after the round's GPU budget was spent: the CPU
    glue test covers the arithmetic; this is its first run on a device, which is why it is the last test of the last file.)"""
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import GaussianDiffusion, Unet3D
    gold = torch.load(os.path.join(golden_dir, "ddim_eta.pt"))[0.5]
    cfg = O.UnetCfg(dim=16, dim_mults=(1, 2))
    model = Unet3D(dim=16, dim_mults=(1, 2), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
                   cond_bias=True, cond_attention='self-stacked', cond_attention_tokens=16, use_temporal_attention_cond=True,
                   cond_to_time='add', per_frame_cond=True, padding_mode='zeros')
    model.load_state_dict(O.synthetic_state_dict(cfg, seed=gold["seed"]), strict=True)
    model.set_compute_dtype(torch.float16)
    gd = GaussianDiffusion(model, image_size=12, channels=3, num_frames=11, timesteps=8, loss_type='l1', use_dynamic_thres=True,
                           sampling_timesteps=4, ddim_sampling_eta=0.5).cuda()
    g = torch.Generator().manual_seed(gold["data_seed"])
    cond = (torch.rand(2, 11, generator=g) * 2 - 1).cuda()
    noises = [torch.randn(2, 3, 11, 12, 12, generator=g) for _ in range(4)]
    with Replay(noises):
        got = gd.sample(cond=cond, guidance_scale=5.0)
    e = rel(got, gold["sample"])
    print("ddim eta=0.5 rel-L2 vs reference:", e)
    assert e < 5e-2                                             # same bound as the eta = 0 loop in test_small_sampling
    torch.manual_seed(5)
    eager = gd.sample(cond=cond, guidance_scale=5.0)
    gd.use_cuda_graph = True
    torch.manual_seed(5)
    graph = gd.sample(cond=cond, guidance_scale=5.0)
    assert rel(graph, eager) < 1e-2


# ------------------------------------------------------------------------------------------------
# padding_mode = 'circular' / 'circular_1d' (model.yaml:13).  Implemented after the round's GPU budget was spent: the host side
# is verified on the CPU against the oracle (tests/test_cpu_forward_glue.py, tests/test_cpu_gemm_formulation.py); what these
# tests add is the kernels reading wrap-padded views that are larger than the output grid.  They passed on their first device run (round 1 driver run, GPUTEST_r01.json).
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["circular", "circular_1d"])
def test_wrap_mode_convolutions_on_device(mode):
    import torch.nn.functional as F
    from videometamaterials_b200 import ops
    dt = torch.bfloat16
    g = torch.Generator().manual_seed(31)
    cl = lambda t: t.permute(0, 2, 3, 1).contiguous()

    def ref_pad(x, p):
        if mode == "circular":
            return F.pad(x, (p, p, p, p), mode="circular")
        return F.pad(F.pad(x, (p, p, 0, 0), mode="circular"), (0, 0, p, p))

    bf, h, w, cin, cout = 6, 24, 16, 64, 128
    x = torch.randn(bf, cin, h, w, generator=g).to(dt).cuda()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).to(dt).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    out = torch.empty(bf, h, w, cout, device="cuda", dtype=dt)
    stats = torch.zeros(bf, 8, 2, dtype=torch.float64, device="cuda")
    ops.conv3x3([cl(x)], ops.pack_conv_taps(wt.float(), [cin], dt), cout, out, mode=mode, bias=bias, gn_stats=stats, gn_group=cout // 8,
                frames_per_sample=1)
    want = cl(F.conv2d(ref_pad(x.float(), 1), wt.float(), bias))
    assert rel(out, want) < 6e-3
    wg = want.double().reshape(bf, h * w, 8, cout // 8)
    assert rel(stats[..., 0], wg.sum(dim=(1, 3))) < 1e-3 and rel(stats[..., 1], (wg * wg).sum(dim=(1, 3))) < 1e-3
    dy = torch.randn(bf, cout, h, w, generator=g).to(dt).cuda()
    xf = x.float().requires_grad_(True)
    wf = wt.float().requires_grad_(True)
    F.conv2d(ref_pad(xf, 1), wf, None).backward(dy.float())
    dx = torch.empty(bf, h, w, cin, device="cuda", dtype=dt)
    ops.conv3x3([cl(dy)], ops.pack_conv_taps(wt.float().flip(2, 3).permute(1, 0, 2, 3), [cout], dt), cin, dx, mode=mode)
    assert rel(dx, cl(xf.grad)) < 6e-3
    dw = torch.zeros(cout, cin, 1, 3, 3, device="cuda")
    ops.wgrad_conv3x3(cl(dy), [cl(x)], dw, mode=mode)
    assert rel(dw[:, :, 0], wf.grad) < 6e-3
    # strided / transposed 4x4 and the 7x7 input conv
    c = 64
    wd = (torch.randn(c, c, 4, 4, generator=g) / (16 * c) ** 0.5).to(dt).cuda()
    xd = x[:, :c]
    out = torch.empty(bf, h // 2, w // 2, c, device="cuda", dtype=dt)
    ops.conv_down(cl(xd), ops.pack_conv_taps(wd.float(), [c], dt), c, out, mode=mode)
    assert rel(out, cl(F.conv2d(ref_pad(xd.float(), 1), wd.float(), None, stride=2))) < 6e-3
    wu = (torch.randn(c, c, 1, 4, 4, generator=g) / (4 * c) ** 0.5).to(dt).cuda()
    out = torch.empty(bf, 2 * h, 2 * w, c, device="cuda", dtype=dt)
    ops.conv_up(cl(xd), ops.pack_conv_up(wu.float(), dt), c, out, mode=mode)
    assert rel(out, cl(F.conv_transpose2d(ref_pad(xd.float(), 2), wu[:, :, 0].float(), None, stride=2, padding=5))) < 6e-3


@pytest.mark.parametrize("mode", ["circular", "circular_1d"])
def test_wrap_mode_network_against_the_oracle(mode):
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import GaussianDiffusion, Unet3D
    cfg = O.UnetCfg(dim=16, dim_mults=(1, 2), padding_mode=mode)
    sd = O.synthetic_state_dict(cfg, seed=41)
    model = Unet3D(dim=16, dim_mults=(1, 2), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
                   cond_bias=True, cond_attention='self-stacked', cond_attention_tokens=16, use_temporal_attention_cond=True,
                   cond_to_time='add', per_frame_cond=True, padding_mode=mode)
    model.load_state_dict(sd, strict=True)
    model.set_compute_dtype(torch.float16)
    gd = GaussianDiffusion(model, image_size=16, channels=3, num_frames=11, timesteps=8, loss_type='l1', use_dynamic_thres=True,
                           sampling_timesteps=8).cuda()
    g = torch.Generator().manual_seed(42)
    b = 2
    x = torch.randn(b, 3, 11, 16, 16, generator=g)
    cond = torch.rand(b, 11, generator=g) * 2 - 1
    t = torch.tensor([1, 6])
    noise = torch.randn(b, 3, 11, 16, 16, generator=g)
    x01 = torch.rand(b, 3, 11, 16, 16, generator=g)
    with torch.no_grad():
        y = model(x.cuda(), t.cuda(), cond=cond.cuda(), null_cond_prob=0.0)
        y_ref = O.unet_forward(sd, cfg, x, t, cond, torch.zeros(b, dtype=torch.bool))
    e = rel(y, y_ref)
    print("wrap-mode forward rel-L2:", mode, e)
    assert e < FWD_TOL[torch.float16]
    P = {k: v.clone().requires_grad_(v.is_floating_point() and "freqs" not in k) for k, v in sd.items()}
    loss_ref = O.p_losses(P, cfg, O.schedule(8), x01, t, cond, noise, torch.zeros(b, dtype=torch.bool))
    loss_ref.backward()
    loss = gd.p_losses((x01 * 2 - 1).cuda(), t.cuda(), cond=cond.cuda(), noise=noise.cuda(), null_cond_prob=0.0)
    (loss * 4096.0).backward()
    torch.cuda.synchronize()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) / float(loss_ref.detach()) < 2e-3
    params = dict(model.named_parameters())
    worst = max((abs(float(params[k].grad.norm()) / 4096.0 - float(p.grad.norm())) / (float(p.grad.norm()) + 1e-7 / 0.08), k)
                for k, p in P.items() if p.requires_grad and p.grad is not None and float(p.grad.norm()) > 0)
    print("wrap-mode worst grad-norm deviation:", mode, worst)
    assert worst[0] < 0.08, worst


def test_module_forward_is_differentiable(gold_small):
    """calling the module under autograd must give gradients (the reference's nn.Module does, VDDP:730-821), not a
    silently detached tensor.  d/dparams of sum(y * w) against the CPU oracle's autograd, fp16 activations with a scale of 256."""
    from oracle import vdm_oracle as O
    g = gold_small
    model, gd, sd = build(16, (1, 2), g["T"], g["size"], g["T"], torch.float16, g["seed"])
    cfg = O.UnetCfg(dim=16, dim_mults=(1, 2))
    x, t, cond = g["x"], g["t"], g["cond"]
    w = torch.randn(x.shape, generator=torch.Generator().manual_seed(5))
    y = model(x.cuda(), t.cuda(), cond=cond.cuda(), null_cond_prob=0.0)
    assert y.requires_grad and y.shape == x.shape
    ((y * w.cuda()).sum() * 256.0).backward()
    torch.cuda.synchronize()
    with torch.no_grad():
        y2 = model(x.cuda(), t.cuda(), cond=cond.cuda(), null_cond_prob=0.0)
    assert not y2.requires_grad and rel(y2, y.detach()) < 1e-3
    P = {k: v.clone().requires_grad_(v.is_floating_point() and "freqs" not in k) for k, v in sd.items()}
    y_ref = O.unet_forward(P, cfg, x, t, cond, torch.zeros(x.shape[0], dtype=torch.bool))
    (y_ref * w).sum().backward()
    assert rel(y.detach(), y_ref.detach()) < FWD_TOL[torch.float16]
    params = dict(model.named_parameters())
    worst = max((abs(float(params[k].grad.norm()) / 256.0 - float(p.grad.norm())) / (float(p.grad.norm()) + 1e-6), k)
                for k, p in P.items() if p.requires_grad and p.grad is not None and float(p.grad.norm()) > 0)
    print("differentiable forward: worst grad-norm deviation", worst)
    assert worst[0] < 0.05, worst


@pytest.mark.parametrize("tcond,c2t,dim", [(False, "add", 16), (True, "concat", 16), (False, "concat", 64)])
def test_config_flags_network_against_the_oracle(tcond, c2t, dim):
    """`use_temporal_attention_cond=False` / `cond_to_time='concat'` (model.yaml:21-22; the oracle's branches are pinned to the unmodified
    reference by tests/golden/config_flags.pt): forward, guided forward, loss and every gradient norm on the kernels.  dim 64 puts the
    fused temporal-attention kernel on the label-free path at level 0; 'concat' takes the torch statement of the conditioning path."""
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import GaussianDiffusion, Unet3D
    cfg = O.UnetCfg(dim=dim, dim_mults=(1, 2), temporal_cond=tcond, cond_to_time=c2t)
    sd = O.synthetic_state_dict(cfg, seed=43)
    model = Unet3D(dim=dim, dim_mults=(1, 2), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True, resnet_groups=8,
                   cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=tcond, cond_to_time=c2t, per_frame_cond=True)
    model.load_state_dict(sd, strict=True)
    model.set_compute_dtype(torch.float16)
    gd = GaussianDiffusion(model, image_size=16, channels=3, num_frames=11, timesteps=8, loss_type='l1', use_dynamic_thres=True,
                           sampling_timesteps=8).cuda()
    g = torch.Generator().manual_seed(44)
    b = 2
    x = torch.randn(b, 3, 11, 16, 16, generator=g)
    cond = torch.rand(b, 11, generator=g) * 2 - 1
    t = torch.tensor([1, 6])
    noise = torch.randn(b, 3, 11, 16, 16, generator=g)
    x01 = torch.rand(b, 3, 11, 16, 16, generator=g)
    with torch.no_grad():
        y = model(x.cuda(), t.cuda(), cond=cond.cuda(), null_cond_prob=0.0)
        y_ref = O.unet_forward(sd, cfg, x, t, cond, torch.zeros(b, dtype=torch.bool))
        yg = model.forward_with_guidance_scale(x.cuda(), t.cuda(), cond=cond.cuda(), guidance_scale=3.0)
        yg_ref = O.unet_forward_guided(sd, cfg, x, t, cond, 3.0)
    print("config flags forward rel-L2:", tcond, c2t, rel(y, y_ref), rel(yg, yg_ref))
    assert rel(y, y_ref) < FWD_TOL[torch.float16] and rel(yg, yg_ref) < 2 * FWD_TOL[torch.float16]
    P = {k: v.clone().requires_grad_(v.is_floating_point() and "freqs" not in k) for k, v in sd.items()}
    loss_ref = O.p_losses(P, cfg, O.schedule(8), x01, t, cond, noise, torch.zeros(b, dtype=torch.bool))
    loss_ref.backward()
    loss = gd.p_losses((x01 * 2 - 1).cuda(), t.cuda(), cond=cond.cuda(), noise=noise.cuda(), null_cond_prob=0.0)
    (loss * 4096.0).backward()
    torch.cuda.synchronize()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) / float(loss_ref.detach()) < 2e-3
    params = dict(model.named_parameters())
    worst = max((abs(float(params[k].grad.norm()) / 4096.0 - float(p.grad.norm())) / (float(p.grad.norm()) + 1e-7 / 0.08), k)
                for k, p in P.items() if p.requires_grad and p.grad is not None and float(p.grad.norm()) > 0)
    print("config flags worst grad-norm deviation:", tcond, c2t, worst)
    assert worst[0] < 0.08, worst
    if not tcond:       # the temporal blocks' label projections are dead under the flag: no gradient here either
        gk = params["downs.0.3.fn.fn.fn.to_k.weight"].grad
        assert gk is None or float(gk.abs().max()) == 0.0
