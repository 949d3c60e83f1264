"""Generate tests/golden/host_*.pt from the UNMODIFIED reference's host-side code (run in the build container only).

    python oracle/make_golden_host.py

Pins the rows either side of the hot path to the reference itself:
  * `Dataset.__getitem__` / labels / `min_max_values.csv` (VDDP:1126-1397) on small synthetic folders in the reference's
    on-disk format (Lagrangian and Eulerian frames, per-frame and whole-curve labels, GIFs that need resize + crop + frame
    padding, a validation folder normalised with the training folder's scaling);
  * `clean_pred` (src/utils.py:32-82) on random and tie-breaking geometries;
  * `Trainer.save_preds` / `remove_padding` (VDDP:1848-1919): tiled GIFs and `geometries.csv`;
  * `Trainer.eval_target` (VDDP:1755-1846): conditioning interpolation / normalisation / fan-out, with a stub sampler;
  * `Trainer.cond_to_gpu` (VDDP:1506-1532) for several (rows, ranks, test batch) combinations.
The reference classes are driven through a stand-in `self` carrying just the attributes those methods read; nothing of the
reference is modified.  Inputs are stored next to the outputs so the tests need neither the reference nor a particular
numpy random stream.
"""
import functools
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)                                   # the reference's `denoising_diffusion_pytorch` shadows this repo's
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
os.chdir(REF)                                             # the reference imports `src.*` relative to its repo root

from denoising_diffusion_pytorch.video_denoising_diffusion_pytorch import Dataset as RefDataset  # noqa: E402
from denoising_diffusion_pytorch.video_denoising_diffusion_pytorch import Trainer as RefTrainer  # noqa: E402
from src.utils import clean_pred as ref_clean_pred  # noqa: E402

from videometamaterials_b200.accel import Accelerator  # noqa: E402
from videometamaterials_b200.dataset import synthetic_dataset_arrays, write_dataset  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def gif_to_array(path):
    from PIL import Image
    img, frames, i = Image.open(path), [], 0
    while True:
        try:
            img.seek(i)
        except EOFError:
            break
        frames.append(np.asarray(img.convert('L'), dtype=np.uint8).copy())
        i += 1
    return np.stack(frames)


# ------------------------------------------------------------------------------------------------ Dataset
def dataset_cases():
    cases = []
    specs = [
        dict(name="lagrangian_per_frame", frame="lagrangian", n=3, size=12, hw=None, gif_frames=11, num_frames=11, per_frame=True, sel=[0, 1, 3], points=51),
        dict(name="lagrangian_resize_crop_pad", frame="lagrangian", n=2, size=12, hw=(15, 18), gif_frames=9, num_frames=11, per_frame=True, sel=[0, 1, 2, 3], points=11),
        dict(name="eulerian_whole_curve", frame="eulerian", n=2, size=12, hw=None, gif_frames=11, num_frames=11, per_frame=False, sel=[0, 1, 2, 3], points=21),
        dict(name="lagrangian_single_frame", frame="lagrangian", n=2, size=12, hw=None, gif_frames=3, num_frames=1, per_frame=False, sel=[0, 1, 2], points=21),
    ]
    for k, sp in enumerate(specs):
        tmp = tempfile.mkdtemp() + "/"
        train, val = tmp + "training/", tmp + "validation/"
        arrays_t = synthetic_dataset_arrays(sp["n"], sp["size"], sp["gif_frames"], sp["points"], seed=10 + k, reference_frame=sp["frame"], size_hw=sp["hw"])
        arrays_v = synthetic_dataset_arrays(2, sp["size"], sp["gif_frames"], sp["points"], seed=50 + k, reference_frame=sp["frame"], size_hw=sp["hw"])
        write_dataset(train, *arrays_t)
        write_dataset(val, *arrays_v)
        ds = RefDataset(train, sp["size"], labels_scaling=None, selected_channels=list(sp["sel"]), num_frames=sp["num_frames"],
                        per_frame_cond=sp["per_frame"], reference_frame=sp["frame"])
        dv = RefDataset(val, sp["size"], labels_scaling=ds.labels_scaling, selected_channels=list(sp["sel"]), num_frames=sp["num_frames"],
                        per_frame_cond=sp["per_frame"], reference_frame=sp["frame"])
        items = [ds[i] for i in range(len(ds))]
        items_v = [dv[i] for i in range(len(dv))]
        case = dict(spec=sp,
                    train=dict(fields={k2: torch.from_numpy(v) for k2, v in arrays_t[0].items()}, ranges=torch.from_numpy(arrays_t[1]), curves=torch.from_numpy(arrays_t[2])),
                    val=dict(fields={k2: torch.from_numpy(v) for k2, v in arrays_v[0].items()}, ranges=torch.from_numpy(arrays_v[1]), curves=torch.from_numpy(arrays_v[2])),
                    x=torch.stack([it[0] for it in items]), labels=torch.stack([it[1] for it in items]),
                    x_val=torch.stack([it[0] for it in items_v]), labels_val=torch.stack([it[1] for it in items_v]),
                    min_max_csv=open(train + "min_max_values.csv").read(),
                    zero_u_2=None if ds.zero_u_2 is None else ds.zero_u_2.clone(),
                    globalmin=ds.labels_scaling.globalmin.clone(), globalmax=ds.labels_scaling.globalmax.clone(),
                    selected_after=list(ds.selected_channels))
        print(f"dataset case {sp['name']}: x {tuple(case['x'].shape)} labels {tuple(case['labels'].shape)}")
        cases.append(case)
    return cases


# ------------------------------------------------------------------------------------------------ clean_pred
def clean_pred_cases():
    rng = np.random.default_rng(7)
    p = 10
    geoms = []
    while len(geoms) < 48:
        dens = rng.uniform(0.15, 0.75)
        g = (rng.random((p, p)) < dens).astype(float)
        if (g[1:] * g[:-1]).sum() + (g[:, 1:] * g[:, :-1]).sum() == 0:
            continue                                          # the reference raises IndexError without any edge
        geoms.append(g * rng.uniform(0.51, 1.0, (p, p)) + (1 - g) * rng.uniform(0.0, 0.49, (p, p)))
    # equal-size components: a single-row pair before (raster order) a single-column pair, and the other way round
    t1 = np.zeros((p, p)); t1[0, 0:2] = 1; t1[5:7, 5] = 1
    t2 = np.zeros((p, p)); t2[0:2, 8] = 1; t2[4, 0:2] = 1; t2[8, 3:5] = 1
    t3 = np.zeros((p, p)); t3[2, 2:5] = 1; t3[6, 1:4] = 1; t3[9, 9] = 1; t3[0, 9] = 1
    t4 = np.zeros((p, p)); t4[1:3, 1:3] = 1; t4[5:7, 5:7] = 1; t4[0, 5:9] = 1
    geoms += [t1, t2, t3, t4]
    geoms = np.stack(geoms)
    out = ref_clean_pred(geoms.copy(), p)
    print("clean_pred cases:", geoms.shape, "->", out.shape)
    return dict(geoms=torch.from_numpy(geoms), pixels=p, out=torch.from_numpy(out))


# ------------------------------------------------------------------------------------------------ Trainer host methods
class Stand:
    """Carries the attributes the reference's Trainer methods read from `self`."""


def stand_in(results_folder, frame, num_frames, selected, zero_u_2, test_batch_size, per_frame, labels_scaling=None, sample_fn=None):
    s = Stand()
    s.accelerator = Accelerator(cpu=True)
    s.device = torch.device("cpu")
    s.results_folder = Path(results_folder)
    s.step = 7
    s.selected_channels = selected
    s.reference_frame = frame
    s.num_frames = num_frames
    s.per_frame_cond = per_frame
    s.test_batch_size = test_batch_size
    s.log_fn = lambda *a, **k: None
    s.ds = Stand()
    s.ds.zero_u_2 = zero_u_2
    s.ds.labels_scaling = labels_scaling
    s.ema_model = Stand()
    s.ema_model.sample = sample_fn
    for m in ("remove_padding", "save_preds", "cond_to_gpu"):
        setattr(s, m, functools.partial(getattr(RefTrainer, m), s))
    return s


def read_outputs(save_dir, selected):
    gifs = {int(c): torch.from_numpy(gif_to_array(f"{save_dir}/gifs/prediction_channel_{c}.gif")) for c in selected}
    geom = torch.from_numpy(np.atleast_2d(np.genfromtxt(f"{save_dir}/geometries.csv", delimiter=',')))
    return gifs, geom


def save_preds_cases():
    cases = []
    g = torch.Generator().manual_seed(3)
    for frame, num_frames, sel in (("lagrangian", 11, [0, 1, 3]), ("eulerian", 11, [0, 1, 2, 3]), ("lagrangian", 1, [0, 1])):
        os.chdir(tempfile.mkdtemp())
        n_real, lengths, max_length, num_samples = 6, torch.tensor([4, 2]), 4, 3      # two ranks, the second padded by two rows
        c, px = len(sel), 12
        vids = torch.rand(n_real, c, num_frames, px, px, generator=g)
        zero_u_2 = torch.tensor([0.43])
        # material / void structure: u_2 close to its zero value in every frame on a random mask (Lagrangian), first channel
        # of the first frame near 0 / 1 (Eulerian)
        mask = torch.rand(n_real, 1, px, px, generator=g) < 0.45
        if frame == "lagrangian" and num_frames != 1:
            vids[:, 1] = torch.where(mask, zero_u_2 + (torch.rand(n_real, num_frames, px, px, generator=g) - 0.5) * 0.03, vids[:, 1])
        else:
            vids[:, 0, 0] = torch.where(mask[:, 0], vids[:, 0, 0] * 0.4, 0.6 + 0.4 * vids[:, 0, 0])
        gathered = torch.zeros(2 * max_length, c, num_frames, px, px)
        gathered[:4] = vids[:4]
        gathered[4:6] = vids[4:]
        gathered[6:] = 0.77                                                            # padding rows must be dropped
        s = stand_in("run", frame, num_frames, sel, zero_u_2, 2, True)
        os.makedirs("run/training/step_7/gifs", exist_ok=True)
        RefTrainer.save_preds(s, gathered.clone(), lengths, max_length, num_samples=num_samples, mode="training")
        gifs, geom = read_outputs("run/training/step_7", sel)
        print(f"save_preds {frame} f={num_frames}: gif {tuple(next(iter(gifs.values())).shape)} geometries {tuple(geom.shape)}")
        cases.append(dict(frame=frame, num_frames=num_frames, selected=sel, gathered=gathered, lengths=lengths, max_length=max_length,
                          num_samples=num_samples, zero_u_2=zero_u_2, gifs=gifs, geometries=geom))
    return cases


def stub_sample(image_size, channels, num_frames):
    """Deterministic stand-in for GaussianDiffusion.sample: a smooth function of the conditioning only."""
    def sample(cond, guidance_scale=1.):
        b = cond.shape[0]
        yy, xx = torch.meshgrid(torch.linspace(0, 1, image_size), torch.linspace(0, 1, image_size), indexing="ij")
        base = torch.stack([torch.sin(3.0 * (k + 1) * xx + 2.0 * yy) for k in range(channels)])            # (c, h, w)
        cf = cond[:, :num_frames] if cond.shape[1] >= num_frames else cond[:, :1].expand(b, num_frames)
        v = 0.5 + 0.5 * torch.tanh(base[None, :, None] * (1.0 + cf[:, None, :, None, None]) * guidance_scale * 0.3)
        return v
    return sample


def eval_target_cases():
    from src.normalization import Normalization as RefNormalization
    cases = []
    targets11 = np.array([[2.81, 30.4, 49.7, 60.1, 67.4, 73.1, 77.4, 81.3, 84.3, 86.1, 85.1],
                          [0.979, 8.76, 12.8, 14.7, 16.4, 18.5, 26.3, 33.5, 41.6, 48.5, 54.1],
                          [0.222, 0.742, 0.612, 0.524, 0.454, 0.401, 0.362, 0.334, 2.33, 4.13, 4.69]])
    targets51 = np.cumsum(np.random.default_rng(5).random((2, 51)), axis=1) * 3.0
    targets51[:, 0] = 0.
    for name, target, per_frame, num_preds in (("per_frame_11_points", targets11, True, 2), ("per_frame_51_points", targets51, True, 1),
                                               ("single_row", targets11[:1], True, 3)):
        os.chdir(tempfile.mkdtemp())
        np.savetxt("targets.csv", target, delimiter=',')
        scaling = RefNormalization(torch.tensor([[0.0, 25.0], [100.0, 50.0]]), ['continuous'] * 2, 'global-min-max-2')
        scaling.cols = 11                                        # the reference normalises column by column over `cols`
        scaling.dataType = ['continuous'] * 11
        seen = []
        base = stub_sample(12, 3, 11)

        def sample(cond, guidance_scale=1., _seen=seen, _base=base):
            _seen.append(cond.clone())
            return _base(cond, guidance_scale)

        s = stand_in("run", "lagrangian", 11, [0, 1, 3], torch.tensor([0.5]), 2, per_frame, labels_scaling=scaling, sample_fn=sample)
        s.accelerator.unwrap_model = lambda m: m
        RefTrainer.eval_target(s, "targets.csv", guidance_scale=5., num_preds=num_preds)
        mode = "eval_target_w_5.0_0"
        gifs, geom = read_outputs(f"run/{mode}/step_7", [0, 1, 3])
        print(f"eval_target {name}: {len(seen)} sampler calls, geometries {tuple(geom.shape)}")
        cases.append(dict(name=name, target=torch.from_numpy(target), num_preds=num_preds, conds=seen, mode=mode, gifs=gifs, geometries=geom,
                          globalmin=0.0, globalmax=100.0))
    return cases


def cond_to_gpu_cases():
    cases = []
    for rows, ranks, tb in ((4, 1, 2), (5, 2, 2), (3, 4, 2), (9, 4, 1), (8, 8, 2), (7, 2, 3), (1, 2, 2)):
        cond = torch.arange(rows * 11, dtype=torch.float32).reshape(rows, 11)
        per_rank = []
        for r in range(ranks):
            s = Stand()
            s.accelerator = Stand()
            s.accelerator.process_index, s.accelerator.num_processes = r, ranks
            s.test_batch_size = tb
            per_rank.append([[int(v) for v in (c[:, 0] / 11).tolist()] for c in RefTrainer.cond_to_gpu(s, cond)])
        cases.append(dict(rows=rows, ranks=ranks, test_batch_size=tb, row_ids=per_rank))
    return cases


if __name__ == "__main__":
    here = os.getcwd()
    torch.save(dict(cases=dataset_cases()), os.path.join(OUT, "host_dataset.pt"))
    torch.save(dict(clean_pred=clean_pred_cases(), save_preds=save_preds_cases(), eval_target=eval_target_cases(),
                    cond_to_gpu=cond_to_gpu_cases()), os.path.join(OUT, "host_trainer.pt"))
    os.chdir(here)
    for f in ("host_dataset.pt", "host_trainer.pt"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
