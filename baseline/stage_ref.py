"""Stage the UNMODIFIED reference into the git-ignored baseline/_ref/ so that it travels to the GPU box with the snapshot.

    python baseline/stage_ref.py

Copies /root/reference/{denoising_diffusion_pytorch,src,data/target_responses.csv,model.yaml,main.py} byte for byte (a
manifest with sha256 sums is written next to them; nothing is edited) and places the stand-in packages of oracle/shims
(einops_exts, rotary_embedding_torch, accelerate, imageio, matplotlib: not installed in this image, no network) beside
them.  bench.py's `--impl reference` (CPU) and `--impl torch-gpu` (the reference's own torch path on the B200) import the
reference from there.  The product never imports it.  pip-installing the reference is not possible: it ships no
setup.py / pyproject.toml, hence the plain copy.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VMM_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
ITEMS = ["denoising_diffusion_pytorch", "src", "data/target_responses.csv", "model.yaml", "main.py"]


def stage() -> str:
    if not os.path.isdir(REF):
        raise FileNotFoundError(f"{REF} not found: the reference can only be staged in the build container")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    manifest = {}
    for item in ITEMS:
        src, dst = os.path.join(REF, item), os.path.join(DST, item)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.isdir(src):
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__"))
        else:
            shutil.copy2(src, dst)
    for base, _, files in os.walk(DST):
        for f in files:
            p = os.path.join(base, f)
            manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    shims = os.path.join(DST, "_shims")
    shutil.copytree(os.path.join(ROOT, "oracle", "shims"), shims, ignore=shutil.ignore_patterns("__pycache__"))
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1)
    return DST


def import_reference():
    """(module VDDP, path) with baseline/_ref first on sys.path; the caller's cwd is switched to baseline/_ref because the
    reference imports `src.*` relative to its repository root."""
    if not os.path.isdir(os.path.join(DST, "denoising_diffusion_pytorch")):
        raise FileNotFoundError("baseline/_ref is not staged (run python baseline/stage_ref.py in the build container)")
    for k in [k for k in sys.modules if k == "denoising_diffusion_pytorch" or k.startswith("denoising_diffusion_pytorch.") or k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    sys.path.insert(0, os.path.join(DST, "_shims"))
    sys.path.insert(0, DST)
    os.chdir(DST)
    import denoising_diffusion_pytorch.video_denoising_diffusion_pytorch as vddp
    assert os.path.realpath(vddp.__file__).startswith(os.path.realpath(DST)), vddp.__file__
    return vddp


if __name__ == "__main__":
    print(stage())
