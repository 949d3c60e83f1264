"""Parity at the BENCHMARKED configuration (dim 64, 4 levels, 96x96x11, 256 timesteps) against outputs of the unmodified
reference (fp32, CPU) committed under tests/golden/ by oracle/make_golden_full.py:

  * full_grads.pt            loss and every parameter gradient of p_losses().backward() for the 'l1' and the 'l2' loss
  * full_sample_ancestral.pt one 256-step ancestral sample (guidance 5, dynamic thresholding), noise regenerated from seeds
  * full_sample_ddim.pt      one 250-step DDIM sample

The sampling tests REPORT the end-to-end relative L2 against the north star's 1e-3 target (SURVEY.md section 0 D5 explains why
16-bit activations cannot be expected to meet it over 256 stochastic steps) and assert looser, measured bounds; the numbers
are copied into DESIGN.md section 4.
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

NOISE_SEED = 1000
SHAPE = (1, 3, 11, 96, 96)
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "full_parity_report.jsonl")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def report(**kw):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(json.dumps(kw) + "\n")


def build(dtype, sampling_T=256, loss_type="l1"):
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import GaussianDiffusion, Unet3D
    model = Unet3D(dim=64, dim_mults=(1, 2, 4, 8), channels=3, attn_heads=8, attn_dim_head=32, use_sparse_linear_attn=True,
                   resnet_groups=8, cond_bias=True, cond_attention='self-stacked', cond_attention_tokens=16,
                   use_temporal_attention_cond=True, cond_to_time='add', per_frame_cond=True, padding_mode='zeros')
    model.load_state_dict(O.synthetic_state_dict(O.UnetCfg(), seed=0), strict=True)
    model.set_compute_dtype(dtype)
    gd = GaussianDiffusion(model, image_size=96, channels=3, num_frames=11, timesteps=256, loss_type=loss_type, use_dynamic_thres=True,
                           sampling_timesteps=sampling_T).cuda()
    return model, gd


@pytest.mark.parametrize("loss_type", ["l2", "l1"])
@pytest.mark.parametrize("dtype,scale", [(torch.bfloat16, 1.0), (torch.float16, 4096.0)])
def test_full_config_gradients(golden_dir, dtype, scale, loss_type):
    """Every parameter gradient of the benchmarked model (b = 1, t = 100) against the reference's autograd.  The smooth 'l2' loss
    gives the tight bound; 'l1' back-propagates sign(pred - noise), where a few sign flips from 16-bit rounding move every
    gradient (same loss the bench trains with)."""
    g = torch.load(os.path.join(golden_dir, "full_grads.pt"))
    gl = g[loss_type]
    model, gd = build(dtype, loss_type=loss_type)
    gen = torch.Generator().manual_seed(g["data_seed"])
    x01 = torch.rand(*SHAPE, generator=gen)
    cond = torch.rand(1, 11, generator=gen) * 2 - 1
    noise = torch.randn(*SHAPE, generator=gen)
    t = torch.tensor([g["t"]])
    loss = gd.p_losses((x01 * 2 - 1).cuda(), t.cuda(), cond=cond.cuda(), noise=noise.cuda(), null_cond_prob=0.0)
    (loss * scale).backward()
    torch.cuda.synchronize()
    loss_err = abs(float(loss) - gl["loss"]) / gl["loss"]
    params = dict(model.named_parameters())
    rows = []
    for k, want in gl["grad_slices"].items():
        gr = params[k].grad.flatten().float().cpu() / scale
        got = gr[::max(1, gr.numel() // 256)][:256]
        wn = gl["grad_norms"][k]
        rows.append(dict(k=k, slice_err=float((got - want).norm() / want.norm().clamp_min(1e-30)),
                         norm_err=abs(float(gr.norm()) - wn) / max(wn, 1e-30), norm=wn, numel=gr.numel()))
    for k in gl["grad_none"]:
        assert float(params[k].grad.abs().max()) == 0.0, k
    rows.sort(key=lambda r: -r["slice_err"])
    sl = sorted(r["slice_err"] for r in rows)
    ne = sorted(r["norm_err"] for r in rows)
    # fp16 floor: gradients of the level-0/1 linear-attention parameters sit below fp16's normal range even
    # with the 4096 loss scale; they are excluded from the fp16 maximum (bf16 training has no such floor) and reported
    floor = [r for r in rows if dtype == torch.float16 and r["norm"] < 1e-5]
    kept = [r for r in rows if r not in floor]
    summary = dict(test="full_grads", loss_type=loss_type, dtype=str(dtype), loss_err=loss_err, n_params=len(rows),
                   slice_err_median=sl[len(sl) // 2], slice_err_p90=sl[int(0.9 * len(sl))], slice_err_max=max(r["slice_err"] for r in kept),
                   norm_err_median=ne[len(ne) // 2], norm_err_max=max(r["norm_err"] for r in kept), fp16_floor_excluded=len(floor),
                   worst=[(r["k"], round(r["slice_err"], 4), r["norm"]) for r in kept[:5]])
    print(summary)
    report(**summary)
    assert loss_err < (2e-3 if dtype == torch.float16 else 1e-2), summary
    # Measured on B200 (round 2, gpurun_out/full_parity_report.jsonl -> DESIGN.md section 4), relative L2 of 256-entry strided slices of
    # every parameter gradient / relative error of every gradient norm:
    #   l2 bf16: median 0.75 %, p90 2.4 %, max 3.8 % | norms: median 0.09 %, max 1.2 %     l2 fp16: median 0.13 %, p90 0.39 %, max 0.6 %
    #   l1 bf16: median 1.9 %,  p90 6.7 %, max 13.5 % | norms: median 0.13 %, max 2.1 %     l1 fp16: median 1.0 %,  max 3.1 %
    # bf16 keeps 8 mantissa bits per stored activation / gradient (2^-9 per rounding, ~200 roundings along the deepest path),
    # so percent-level element errors are its floor; the gradient NORMS, which average the rounding noise, agree to ~1e-3.
    if loss_type == "l2":
        assert summary["slice_err_median"] < (1.5e-2 if dtype == torch.bfloat16 else 4e-3), summary
        assert summary["slice_err_p90"] < (4e-2 if dtype == torch.bfloat16 else 1e-2), summary
        assert summary["slice_err_max"] < (6e-2 if dtype == torch.bfloat16 else 1.5e-2), summary
        assert summary["norm_err_median"] < (3e-3 if dtype == torch.bfloat16 else 1e-3), summary
        assert summary["norm_err_max"] < (2.5e-2 if dtype == torch.bfloat16 else 1e-2), summary
    else:
        assert summary["slice_err_median"] < (4e-2 if dtype == torch.bfloat16 else 2e-2), summary
        assert summary["slice_err_max"] < (0.25 if dtype == torch.bfloat16 else 0.08), summary
        assert summary["norm_err_max"] < (5e-2 if dtype == torch.bfloat16 else 3e-2), summary


class SeededNoise:
    """torch.randn / randn_like / Tensor.normal_ replay the CPU draws `randn(SHAPE, Generator(NOISE_SEED + k))` of the golden run."""

    def __enter__(self):
        self.k = 0
        self._a, self._b, self._c = torch.randn, torch.randn_like, torch.Tensor.normal_
        randn = torch.randn

        def draw():
            t = randn(SHAPE, generator=torch.Generator().manual_seed(NOISE_SEED + self.k)).cuda()
            self.k += 1
            return t

        def normal_(t, *a, **kw):
            return t.copy_(draw())

        torch.randn = lambda *a, **kw: draw()
        torch.randn_like = lambda *a, **kw: draw()
        torch.Tensor.normal_ = normal_
        return self

    def __exit__(self, *e):
        torch.randn, torch.randn_like, torch.Tensor.normal_ = self._a, self._b, self._c


def _sl(x):
    return x[:, :, :, ::4, ::4]


def test_full_ancestral_sample_end_to_end(golden_dir):
    """BASELINE configs[2] shape, shipped sampler: 256 ancestral steps, guidance 5, dynamic thresholding, fp16 activations, on
    the reference's noise.  Reports rel-L2 after selected steps and at the end against the 1e-3 target."""
    path = os.path.join(golden_dir, "full_sample_ancestral.pt")
    if not os.path.exists(path):
        pytest.skip("tests/golden/full_sample_ancestral.pt not generated yet (oracle/make_golden_full.py ancestral)")
    g = torch.load(path)
    model, gd = build(torch.float16)
    cond = g["cond"].cuda()
    errs = {}
    with SeededNoise():
        img = torch.randn(SHAPE)
        for i in reversed(range(256)):
            img = gd.p_sample(img, torch.full((1,), i, device="cuda", dtype=torch.long), cond=cond, guidance_scale=5.0)
            if i in g["states"]:
                errs[i] = rel(_sl(img), g["states"][i]["slice"])
    final = (img + 1) * 0.5
    e_final = rel(final, g["final"])
    # the same through the public sample() with the CUDA-graph loop (normal_() draws replayed)
    gd.use_cuda_graph = True
    with SeededNoise():
        final_graph = gd.sample(cond=cond, guidance_scale=5.0)
    e_graph = rel(final_graph, g["final"])
    summary = dict(test="full_ancestral_256", dtype="float16", rel_l2_final=e_final, rel_l2_final_graph_loop=e_graph,
                   rel_l2_after_step_t={str(k): v for k, v in sorted(errs.items(), reverse=True)}, target=1e-3,
                   meets_target=bool(e_final < 1e-3), max_abs_err=float((final.cpu() - g["final"]).abs().max()))
    print(summary)
    report(**summary)
    assert e_final < 5e-2 and e_graph < 5e-2, summary


def test_full_ddim_sample_end_to_end(golden_dir):
    """BASELINE configs[2] as named: 250-step DDIM (eta 0, no clamp), guidance 5.  With RANDOM-INIT weights the un-clamped DDIM
    recursion diverges in the reference itself (the golden's state norm grows from 5.5e2 to 1.8e7, elements up to 8.9e4: x0 =
    16426 (x - eps) at the first step and nothing pulls it back), which is beyond fp16's range (65504): this trajectory can only be
    followed with bf16 activations (fp32 exponent range, 8 mantissa bits).  A trained checkpoint keeps |x| ~ 1, where sampling
    runs in fp16 as in the ancestral test above.  The test reports the end-to-end error of the bf16 run; it is a range /
    plumbing check of the DDIM path at full size, not a precision claim."""
    path = os.path.join(golden_dir, "full_sample_ddim.pt")
    if not os.path.exists(path):
        pytest.skip("tests/golden/full_sample_ddim.pt not generated yet (oracle/make_golden_full.py ddim)")
    g = torch.load(path)
    model, gd = build(torch.bfloat16, sampling_T=250)
    assert gd.is_ddim_sampling
    cond = g["cond"].cuda()
    with SeededNoise():
        out = gd.sample(cond=cond, guidance_scale=5.0)
    e_final = rel(out, g["final"])
    gd.use_cuda_graph = True
    with SeededNoise():
        out_graph = gd.sample(cond=cond, guidance_scale=5.0)
    e_graph = rel(out_graph, g["final"])
    summary = dict(test="full_ddim_250", dtype="bfloat16", rel_l2_final=e_final, rel_l2_final_graph_loop=e_graph, target=1e-3,
                   meets_target=bool(e_final < 1e-3), max_abs_err=float((out.cpu() - g["final"]).abs().max()),
                   golden_abs_max=float(g["final"].abs().max()), note="random-init weights: the reference's own DDIM trajectory diverges to |x| ~ 1e5")
    print(summary)
    report(**summary)
    assert e_final == e_final and e_graph == e_graph, summary            # finite
    assert e_final < 0.5 and e_graph < 0.5, summary
