"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: gradient all-reduce over the arena,
gather / pad_across_processes, parameter broadcast, and the contiguous split of conditionings over ranks."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from videometamaterials_b200.accel import Accelerator, broadcast_object_list
    from videometamaterials_b200.blocks_bwd import GradArena
    acc = Accelerator(cpu=True)
    assert acc.num_processes == world and acc.process_index == rank
    torch.manual_seed(rank)                       # different initial weights per rank: prepare() must broadcast rank 0's
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    net = acc.prepare(net)
    w0 = net[0].weight.detach().clone()
    arena = GradArena(net)
    net._vmm_arena = arena
    arena.flat_grad.fill_(float(rank + 1))        # rank-dependent gradients
    acc.all_reduce_gradients()
    mean = sum(range(1, world + 1)) / world
    ok_grad = bool(torch.allclose(arena.flat_grad, torch.full_like(arena.flat_grad, mean)))
    g = acc.gather(torch.tensor(float(rank)))
    v = torch.ones(rank + 1, 2) * rank
    padded = acc.pad_across_processes(v, dim=0)
    gathered = acc.gather(padded)
    obj = broadcast_object_list([{"cond": [1, 2, 3]} if rank == 0 else None])[0]

    class T:      # minimal stand-in carrying what cond_to_gpu reads
        accelerator, test_batch_size = acc, 2
    from videometamaterials_b200.trainer import Trainer
    chunks = Trainer.cond_to_gpu(T, torch.arange(10).reshape(5, 2))
    acc.wait_for_everyone()
    q.put((rank, w0.numpy(), ok_grad, g.tolist(), tuple(padded.shape), tuple(gathered.shape), obj, [c.tolist() for c in chunks]))
    dist.destroy_process_group()


def test_two_rank_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, w0, ok0, g0, ps0, gs0, o0, c0), (r1, w1, ok1, g1, ps1, gs1, o1, c1) = res
    assert torch.equal(torch.from_numpy(w0), torch.from_numpy(w1))                                    # parameters broadcast from rank 0
    assert ok0 and ok1                                            # averaged gradient on both ranks
    assert g0 == g1 == [0.0, 1.0]
    assert ps0 == ps1 == (2, 2) and gs0 == gs1 == (4, 2)
    assert o0 == o1 == {"cond": [1, 2, 3]}
    # VDDP:1506-1532: rank 0 gets rows [0, 2), the last rank the remainder [2, 5), chunked by test_batch_size
    assert c0 == [[[0, 1], [2, 3]]] and c1 == [[[4, 5], [6, 7]], [[8, 9]]]


def _stub_sample(cond, guidance_scale=1.):
    """Deterministic stand-in for GaussianDiffusion.sample (a smooth function of the conditioning only)."""
    size, ch, nf = 12, 3, 11
    yy, xx = torch.meshgrid(torch.linspace(0, 1, size), torch.linspace(0, 1, size), indexing="ij")
    base = torch.stack([torch.sin(3.0 * (k + 1) * xx + 2.0 * yy) for k in range(ch)])
    return 0.5 + 0.5 * torch.tanh(base[None, :, None] * (1.0 + cond[:, None, :nf, None, None]) * guidance_scale * 0.3)


def _eval_target_run(workdir, targets):
    """Build a small Trainer on the CPU accelerator, replace the sampler by the stub and run eval_target + eval_network."""
    import numpy as np
    from videometamaterials_b200 import Accelerator, GaussianDiffusion, Trainer, Unet3D
    from videometamaterials_b200.dataset import fixed_normalization
    os.chdir(workdir)
    m = Unet3D(dim=16, dim_mults=(1, 2), channels=3, cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, per_frame_cond=True)
    gd = GaussianDiffusion(m, image_size=12, channels=3, num_frames=11, timesteps=8, use_dynamic_thres=True, sampling_timesteps=8)
    t = Trainer(gd, None, None, [0, 1, 3], train_batch_size=2, test_batch_size=4, results_folder='run', log=True, null_cond_prob=0.1,
                per_frame_cond=True, reference_frame='lagrangian', accelerator=Accelerator(cpu=True))
    t.step = 3
    t.ds.labels_scaling = fixed_normalization(0., 100.)
    t.ema_model.sample = _stub_sample
    t.model = lambda **kw: kw["x"].mean() * 0 + 0.25 * (1 + t.accelerator.process_index)       # per-rank validation "loss"
    np.random.seed(4)
    gathered = t.eval_target(targets, guidance_scale=5., num_preds=1)
    t.eval_network(0., None, num_samples=3, num_preds=1)
    return t, gathered


def _eval_worker(rank, world, port, workdir, targets, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    t, gathered = _eval_target_run(workdir, targets)
    q.put((rank, tuple(gathered.shape), [l for l in t.accelerator.logs if 'validation loss' in l]))
    dist.destroy_process_group()


@pytest.mark.parametrize("rows", [5, 1])
def test_two_rank_eval_target_matches_single_process(tmp_path, rows):
    """Five conditionings over two ranks (2 + 3, the shorter rank padded for the gather), and a single conditioning (rank 0 gets
    none and contributes an empty, fully padded block): rank 0 must write the same geometries.csv and GIFs as a single process
    does (VDDP:1755-1846), and the validation loss is the mean over ranks."""
    import numpy as np
    targets = str(tmp_path / "targets.csv")
    rng = np.random.default_rng(0)
    np.savetxt(targets, np.cumsum(rng.random((rows, 11)), axis=1) * 15.0, delimiter=',')
    one, two = tmp_path / "one", tmp_path / "two"
    one.mkdir()
    two.mkdir()
    cwd = os.getcwd()
    try:
        t, g1 = _eval_target_run(str(one), targets)
    finally:
        os.chdir(cwd)
    assert tuple(g1.shape) == (rows, 3, 11, 12, 12)
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_eval_worker, args=(r, world, port, str(two), targets, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == (2 * (rows - rows // 2), 3, 11, 12, 12)     # 2 x the longer rank's rows; the rest is padding
    assert res[0][2] and abs(res[0][2][0]['validation loss'] - 0.375) < 1e-6 and not res[1][2]
    sub = "eval_target_w_5.0_0/step_3"
    a = np.genfromtxt(str(one / "run" / sub / "geometries.csv"), delimiter=',')
    b = np.genfromtxt(str(two / "run" / sub / "geometries.csv"), delimiter=',')
    assert np.atleast_2d(a).shape == np.atleast_2d(b).shape == (rows, 36) and np.array_equal(a, b)
    for ch in (0, 1, 3):
        fa = open(str(one / "run" / sub / f"gifs/prediction_channel_{ch}.gif"), "rb").read()
        fb = open(str(two / "run" / sub / f"gifs/prediction_channel_{ch}.gif"), "rb").read()
        assert fa == fb, ch
    # eval_network: the validation loader is sharded over ranks (as accelerate's prepared loader is), so the sampled conditionings
    # differ from the single-process run; the files of the three requested samples must exist with the right shapes
    for root in (one, two):
        assert np.genfromtxt(str(root / "run" / "training/step_3" / "geometries.csv"), delimiter=',').shape == (3, 36)
        assert all(os.path.getsize(str(root / "run" / "training/step_3" / f"gifs/prediction_channel_{ch}.gif")) > 0 for ch in (0, 1, 3))


def _rank_loss(self, x, *args, **kwargs):
    """0.5 |w|^2 * (1 + rank): rank-dependent gradients (rank + 1) * w, whose mean over two ranks is 1.5 w."""
    import torch.distributed as dist
    w = self.denoise_fn.init_conv.weight
    return 0.5 * (1 + dist.get_rank()) * (w ** 2).sum() + 0.0 * x.mean()


def _train_worker(rank, world, port, workdir, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from oracle import vdm_oracle as O
    from videometamaterials_b200 import Accelerator, GaussianDiffusion, Trainer, Unet3D, ops
    os.chdir(workdir)
    import videometamaterials_b200.trainer as trainer_mod
    trainer_mod.REDUCE_CHUNKS = 3                     # the pipelined all-reduce / optimiser path (default 1 = one collective, covered by test_two_rank_gloo)
    GaussianDiffusion.forward = _rank_loss
    GaussianDiffusion.sample = lambda self, cond=None, batch_size=16, guidance_scale=1.: _stub_sample(cond, guidance_scale)
    ops.adam_ema_step = lambda p, g, m, v, ema, lr, b1, b2, eps, step, gs, mode, beta: \
        O.adam_ema_step(p, g, m, v, ema, step, lr=lr, beta1=b1, beta2=b2, eps=eps, grad_scale=gs, ema_mode=mode, ema_beta=beta)
    torch.manual_seed(100 + rank)                     # different initial weights per rank: prepare() broadcasts rank 0's
    m = Unet3D(dim=16, dim_mults=(1, 2), channels=3, cond_bias=True, cond_attention='self-stacked', use_temporal_attention_cond=True, per_frame_cond=True)
    gd = GaussianDiffusion(m, image_size=12, channels=3, num_frames=11, timesteps=8, use_dynamic_thres=True, sampling_timesteps=8)
    t = Trainer(gd, None, None, [0, 1, 3], train_batch_size=2, test_batch_size=4, train_lr=1e-2, train_num_steps=4, step_start_ema=2,
                update_ema_every=2, save_and_sample_every=100, results_folder='run', log=True, null_cond_prob=0.1, per_frame_cond=True,
                reference_frame='lagrangian', accelerator=Accelerator(cpu=True), log_every=1)
    w0 = t.model.denoise_fn.init_conv.weight.detach().clone()
    seen = []
    dl = t.dl

    def recording():
        while True:
            x, c = next(dl)
            seen.append(c.clone())
            yield x, c
    t.dl = recording()
    t.train(num_samples=0)
    # numpy arrays are pickled by value: a torch tensor on an mp.Queue travels as a shared-memory file that is gone when this process
    # exits before the parent has read it (seen once as FileNotFoundError on a loaded machine)
    q.put((rank, w0.numpy(), t.model.denoise_fn.init_conv.weight.detach().numpy().copy(), t.ema_model.denoise_fn.init_conv.weight.detach().numpy().copy(),
           torch.cat(seen).numpy(), os.path.isfile("run/model/step_4/checkpoint.pt")))
    dist.destroy_process_group()


def test_two_rank_training_loop(tmp_path):
    """Trainer.train on two gloo ranks with a stub loss: both ranks start from rank 0's weights, see disjoint shards of the data,
    apply the MEAN of their gradients (VDDP:1629 through DDP; here one all-reduce over the arena) and stay bit-identical; the
    trajectory equals torch.optim.Adam on the averaged gradient, the EMA follows the reference's cadence, one checkpoint is written."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = [tuple(torch.from_numpy(v) if hasattr(v, "dtype") and not torch.is_tensor(v) else v for v in r) for r in res]
    (_, w0a, wa, ea, ca, cka), (_, w0b, wb, eb, cb, ckb) = res
    assert torch.equal(w0a, w0b) and torch.equal(wa, wb) and torch.equal(ea, eb)
    assert cka and ckb                                          # rank 0 wrote it; both see the same folder
    # disjoint data shards: no conditioning row was seen by both ranks within the first epoch (5 steps x 2 clips of 512 each)
    rows_a, rows_b = {tuple(r.tolist()) for r in ca}, {tuple(r.tolist()) for r in cb}
    assert len(rows_a) == len(rows_b) == 10 and not (rows_a & rows_b)
    w = torch.nn.Parameter(w0a.clone())
    opt = torch.optim.Adam([w], lr=1e-2)
    ema = w0a.clone()
    for step in range(5):                                       # steps 0..4 inclusive
        w.grad = 1.5 * w.detach()
        opt.step()
        if step % 2 == 0:
            ema = w.detach().clone() if step < 2 else ema * 0.995 + (1 - 0.995) * w.detach()
    assert torch.allclose(wa, w.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(ea, ema, rtol=1e-5, atol=1e-6)
