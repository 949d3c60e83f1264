"""Trainer.train bookkeeping on the CPU (no kernels): the loss and the sampler are replaced by stubs and the fused optimiser
kernel by the oracle's restatement, so that what is tested is the host logic the reference defines at VDDP:1594-1672 and
1534-1592: step counting, EMA cadence (copy phase, then averaging, every `update_ema_every` steps), evaluation every
`save_and_sample_every` steps, the checkpoint written at the final step, and resuming from it (optimiser moments included).
The expected trajectories come from torch.optim.Adam on a plain tensor."""
import os

import numpy as np
import pytest
import torch


def _stub_sample(self, cond=None, batch_size=16, guidance_scale=1.):
    b = cond.shape[0]
    yy, xx = torch.meshgrid(torch.linspace(0, 1, self.image_size), torch.linspace(0, 1, self.image_size), indexing="ij")
    base = torch.stack([torch.sin(3.0 * (k + 1) * xx + 2.0 * yy) for k in range(self.channels)])
    return 0.5 + 0.5 * torch.tanh(base[None, :, None] * (1.0 + cond[:, None, :, None, None]))


def _quadratic_loss(self, x, *args, **kwargs):
    """0.5 |w|^2 over one parameter tensor: its gradient is the parameter itself."""
    assert set(kwargs) <= {"cond", "null_cond_prob", "prob_focus_present", "focus_present_mask"} and kwargs["cond"].shape == (x.shape[0], 11)
    w = self.denoise_fn.init_conv.weight
    return 0.5 * (w ** 2).sum() + 0.0 * x.mean()


def _make_trainer(tmp_path, train_num_steps, monkeypatch):
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import Accelerator, GaussianDiffusion, Trainer, Unet3D, ops
    from videometamaterials_b200.dataset import write_synthetic_dataset
    monkeypatch.setattr(GaussianDiffusion, "forward", _quadratic_loss)
    monkeypatch.setattr(GaussianDiffusion, "sample", _stub_sample)
    monkeypatch.setattr(ops, "adam_ema_step", lambda p, g, m, v, ema, lr, b1, b2, eps, step, gs, mode, beta:
                        O.adam_ema_step(p, g, m, v, ema, step, lr=lr, beta1=b1, beta2=b2, eps=eps, grad_scale=gs, ema_mode=mode, ema_beta=beta))
    data = str(tmp_path / "data") + "/"
    if not os.path.isdir(data):
        write_synthetic_dataset(data + "training/", 4, image_size=12, num_frames=11, seed=0)
        write_synthetic_dataset(data + "validation/", 2, image_size=12, num_frames=11, seed=1)
    torch.manual_seed(0)
    m = Unet3D(dim=16, dim_mults=(1, 2), channels=3, cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, per_frame_cond=True)
    gd = GaussianDiffusion(m, image_size=12, channels=3, num_frames=11, timesteps=8, use_dynamic_thres=True, sampling_timesteps=8)
    t = Trainer(gd, data + "training/", data + "validation/", [0, 1, 3], train_batch_size=2, test_batch_size=2, train_lr=1e-2,
                train_num_steps=train_num_steps, step_start_ema=5, update_ema_every=3, save_and_sample_every=6, results_folder='run', log=True,
                null_cond_prob=0.1, per_frame_cond=True, reference_frame='lagrangian', accelerator=Accelerator(cpu=True), log_every=1)
    return t


def _simulate(w0, steps, lr=1e-2, step_start_ema=5, every=3, beta=0.995, opt_state=None, ema0=None):
    """Reference semantics on one tensor: Adam step, then (VDDP:1637-1639) the model average when step % every == 0."""
    w = torch.nn.Parameter(w0.clone())
    opt = torch.optim.Adam([w], lr=lr)
    if opt_state is not None:
        opt.load_state_dict(opt_state)
    ema = w0.clone() if ema0 is None else ema0.clone()
    for step in steps:
        w.grad = w.detach().clone()
        opt.step()
        if step % every == 0:
            ema = w.detach().clone() if step < step_start_ema else ema * beta + (1 - beta) * w.detach()
    return w.detach(), ema, opt


def test_train_loop_checkpoint_and_resume(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    t = _make_trainer(tmp_path, 12, monkeypatch)
    assert len(t.ds) == 4 and len(t.ds_test) == 2 and t.test_batch_size == 1
    w0 = t.model.denoise_fn.init_conv.weight.detach().clone()
    other0 = t.model.denoise_fn.final_conv[1].weight.detach().clone()
    np.random.seed(0)
    t.train(num_samples=2, num_preds=1)
    # steps 0..12 inclusive are optimised (VDDP:1611, 1659-1665); the final step writes the only checkpoint
    assert t.step == 12 and t.opt.step_count == 13
    w_want, ema_want, opt = _simulate(w0, range(0, 13))
    assert torch.allclose(t.model.denoise_fn.init_conv.weight.detach(), w_want, rtol=1e-5, atol=1e-6)
    assert torch.allclose(t.ema_model.denoise_fn.init_conv.weight.detach(), ema_want, rtol=1e-5, atol=1e-6)
    assert torch.equal(t.model.denoise_fn.final_conv[1].weight.detach(), other0)            # zero gradient, zero moments: untouched
    assert sorted(os.listdir("run/model")) == ["step_12"]
    for s in (6, 12):                                                                        # 0 < step, step % 6 == 0
        assert np.genfromtxt(f"run/training/step_{s}/geometries.csv", delimiter=',').shape == (2, 36)
        assert sorted(os.listdir(f"run/training/step_{s}/gifs")) == [f"prediction_channel_{c}.gif" for c in (0, 1, 3)]
    logs = t.accelerator.logs
    assert [l["step"] for l in logs if "training loss" in l] == list(range(13))
    assert [l["step"] for l in logs if "validation loss" in l] == [6, 12]
    ck = torch.load("run/model/step_12/checkpoint.pt", map_location="cpu")
    assert set(ck) == {"model", "optimizer", "steps", "ema"} and ck["steps"] == 12
    assert list(ck["model"].keys()) == list(t.model.state_dict().keys()) and len(ck["model"]) == len(ck["ema"])
    assert ck["optimizer"]["param_groups"][0]["lr"] == 1e-2 and len(ck["optimizer"]["state"]) == len(t.opt.arena.params)

    # resume: a fresh process image (new random weights) continues from the checkpoint.  The reference advances `step` twice per
    # iteration when resuming (VDDP:1612-1616 and 1659-1660), so from 12 with train_num_steps = 15 it optimises steps 13 and 15.
    t2 = _make_trainer(tmp_path, 15, monkeypatch)
    torch.manual_seed(5)
    with torch.no_grad():
        t2.model.denoise_fn.init_conv.weight.add_(1.0)
    t2.train(load_model_step=12, num_samples=0)
    assert t2.step == 15 and t2.opt.step_count == 15
    w_res, ema_res, _ = _simulate(w_want, [13, 15], opt_state=opt.state_dict(), ema0=ema_want)
    assert torch.allclose(t2.model.denoise_fn.init_conv.weight.detach(), w_res, rtol=1e-5, atol=1e-6)
    assert torch.allclose(t2.ema_model.denoise_fn.init_conv.weight.detach(), ema_res, rtol=1e-5, atol=1e-6)
    assert sorted(os.listdir("run/model")) == ["step_12", "step_15"]
    # already trained: nothing happens, nothing is written
    t3 = _make_trainer(tmp_path, 15, monkeypatch)
    t3.train(load_model_step=15, num_samples=0)
    assert t3.step == 15 and t3.opt.step_count == 15 and sorted(os.listdir("run/model")) == ["step_12", "step_15"]
    t4 = _make_trainer(tmp_path, 15, monkeypatch)
    with pytest.raises(FileNotFoundError):
        t4.train(load_model_step=14)


def test_checkpoint_loads_ddp_prefixed_state_dict(tmp_path, monkeypatch):
    """Checkpoints saved by the reference under DistributedDataParallel carry a `module.` prefix on every key (main.py:32)."""
    monkeypatch.chdir(tmp_path)
    t = _make_trainer(tmp_path, 3, monkeypatch)
    sd = {("module." + k): v.clone() + 0.5 for k, v in t.model.state_dict().items() if v.is_floating_point()}
    sd.update({("module." + k): v.clone() for k, v in t.model.state_dict().items() if not v.is_floating_point()})
    os.makedirs("run/model/step_2")
    torch.save(dict(model=sd, optimizer=None, steps=2, ema=sd), "run/model/step_2/checkpoint.pt")
    t.step = 2
    before = t.model.denoise_fn.init_conv.weight.detach().clone()
    t.load()
    assert torch.allclose(t.model.denoise_fn.init_conv.weight.detach(), before + 0.5)
    assert torch.allclose(t.ema_model.denoise_fn.init_conv.weight.detach(), before + 0.5)


def test_optimizer_state_interchanges_with_torch_adam():
    """The reference's checkpoint stores `torch.optim.Adam(model.parameters()).state_dict()` (VDDP:1478, 1547): indices run over
    all parameters, the frozen rotary table included, and parameters that never had a gradient have no entry.  FusedAdam must
    read that layout into the right arena slots and write a state that torch's Adam loads back."""
    from videometamaterials_b200 import Unet3D
    from videometamaterials_b200.trainer import FusedAdam
    torch.manual_seed(0)
    kw = dict(dim=16, dim_mults=(1, 2), channels=3, cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, per_frame_cond=True)
    ref_model, own_model = Unet3D(**kw), Unet3D(**kw)
    own_model.load_state_dict(ref_model.state_dict())
    names = [n for n, _ in ref_model.named_parameters()]
    assert names[5].endswith("rotary_emb.freqs") and not dict(ref_model.named_parameters())[names[5]].requires_grad
    opt = torch.optim.Adam(ref_model.parameters(), lr=3e-4)
    never = lambda n: "to_q." in n or "sign_emb_CNN" in n                      # tensors without a gradient under the shipped config
    for step in range(3):
        for n, p in ref_model.named_parameters():
            p.grad = None if (never(n) or not p.requires_grad) else torch.randn_like(p) * (step + 1)
        opt.step()
    sd = opt.state_dict()
    assert 5 not in sd["state"] and len(sd["param_groups"][0]["params"]) == len(names)
    fused = FusedAdam(own_model, lr=1e-4)
    fused.load_state_dict(sd)
    assert fused.step_count == 3 and fused.lr == 3e-4
    o = 0
    by_name = dict(own_model.named_parameters())
    for p in fused.arena.params:
        n = next(k for k, v in by_name.items() if v is p)
        i = names.index(n)
        k = p.numel()
        if i in sd["state"]:
            assert torch.equal(fused.m[o:o + k].view(p.shape), sd["state"][i]["exp_avg"]), n
            assert torch.equal(fused.v[o:o + k].view(p.shape), sd["state"][i]["exp_avg_sq"]), n
        else:
            assert never(n) and float(fused.m[o:o + k].abs().sum()) == 0.0
        o += k
    # and back: torch's Adam accepts what FusedAdam writes, with the moments under the same indices
    back = torch.optim.Adam(ref_model.parameters(), lr=1e-4)
    out = fused.state_dict()
    back.load_state_dict(out)
    assert back.param_groups[0]["lr"] == 3e-4
    for i, st in sd["state"].items():
        assert torch.equal(out["state"][i]["exp_avg"], st["exp_avg"]) and float(out["state"][i]["step"]) == 3.0
