"""ctypes binding of libvmm_sm100.so (the C ABI declared in include/vmm.h).

The library is the product's only compute path.  If it is missing it is built with nvcc when a
toolkit is present; otherwise import fails loudly.  There is no Python / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvmm_sm100.so")

MAX_VIEWS, MAX_PHASES, MAX_TAPS = 4, 4, 20
FMT_F16, FMT_BF16 = 0, 1


class View4(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dims", C.c_int32 * 4), ("strides", C.c_int64 * 3)]


class Tap(C.Structure):
    _fields_ = [("src", C.c_int32), ("dy", C.c_int32), ("dx", C.c_int32), ("kofs", C.c_int32), ("c", C.c_int32)]


class CgemmParams(C.Structure):
    _fields_ = [
        ("fmt", C.c_int32), ("n_views", C.c_int32), ("a", View4 * MAX_VIEWS),
        ("n_phases", C.c_int32), ("n_taps", C.c_int32 * MAX_PHASES), ("taps", (Tap * MAX_TAPS) * MAX_PHASES),
        ("phase_oy", C.c_int32 * MAX_PHASES), ("phase_ox", C.c_int32 * MAX_PHASES),
        ("w", C.c_void_p), ("n", C.c_int32), ("ktot", C.c_int32),
        ("bf", C.c_int32), ("oh", C.c_int32), ("ow", C.c_int32),
        ("tf", C.c_int32), ("th", C.c_int32), ("tw", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_fp32", C.c_int32),
        ("ohs", C.c_int32), ("ows", C.c_int32), ("sy", C.c_int32), ("sx", C.c_int32),
        ("out2", C.c_void_p), ("ldo2", C.c_int64), ("nsplit", C.c_int32),
        ("bias", C.c_void_p), ("res", C.c_void_p), ("ldr", C.c_int64), ("res2", C.c_void_p), ("ldr2", C.c_int64),
        ("alpha", C.c_float),
        ("gn_stats", C.c_void_p), ("gn_group", C.c_int32), ("frames_per_sample", C.c_int32),
        ("rot", C.c_void_p), ("rot_frames", C.c_int32), ("rot_hw", C.c_int32), ("rot_cols", C.c_int32), ("rot_qcols", C.c_int32),
    ]


class WgradTap(C.Structure):
    _fields_ = [("a_src", C.c_int32), ("b_src", C.c_int32), ("dy", C.c_int32), ("dx", C.c_int32), ("c", C.c_int32),
                ("wofs", C.c_int64)]


class WgradParams(C.Structure):
    _fields_ = [
        ("fmt", C.c_int32), ("n_a_views", C.c_int32), ("n_b_views", C.c_int32),
        ("a", View4 * MAX_VIEWS), ("b", View4 * MAX_VIEWS),
        ("n_taps", C.c_int32), ("taps", WgradTap * MAX_TAPS),
        ("n", C.c_int32), ("bf", C.c_int32), ("oh", C.c_int32), ("ow", C.c_int32),
        ("tf", C.c_int32), ("th", C.c_int32), ("tw", C.c_int32),
        ("dw", C.c_void_p), ("s_m", C.c_int64), ("s_c", C.c_int64), ("s_c2", C.c_int64),
        ("cmod", C.c_int32), ("c_valid", C.c_int32), ("k_valid", C.c_int32),
    ]


COND_MAX_BLOCKS = 24


class CondParams(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("T", C.c_int32), ("dim", C.c_int32), ("td", C.c_int32), ("heads", C.c_int32), ("frames", C.c_int32),
        ("n_res", C.c_int32), ("n_attn", C.c_int32),
        ("time", C.c_void_p), ("cond", C.c_void_p), ("null_mask", C.c_void_p), ("param", C.c_void_p), ("grad", C.c_void_p),
        ("freqs", C.c_void_p), ("buckets", C.c_void_p),
        ("o_w1", C.c_int64), ("o_b1", C.c_int64), ("o_w2", C.c_int64), ("o_b2", C.c_int64), ("o_wse", C.c_int64), ("o_bse", C.c_int64),
        ("o_lng", C.c_int64), ("o_lnb", C.c_int64), ("o_w3", C.c_int64), ("o_b3", C.c_int64), ("o_w4", C.c_int64), ("o_b4", C.c_int64),
        ("o_ntok", C.c_int64), ("o_nhid", C.c_int64), ("o_table", C.c_int64),
        ("res_w", C.c_int64 * COND_MAX_BLOCKS), ("res_b", C.c_int64 * COND_MAX_BLOCKS), ("res_out", C.c_int64 * COND_MAX_BLOCKS),
        ("res_c2", C.c_int32 * COND_MAX_BLOCKS),
        ("att_wk", C.c_int64 * COND_MAX_BLOCKS), ("att_wv", C.c_int64 * COND_MAX_BLOCKS), ("att_out", C.c_int64 * COND_MAX_BLOCKS),
        ("att_temporal", C.c_int32 * COND_MAX_BLOCKS),
        ("bias_out", C.c_int64), ("rot_out", C.c_int64), ("out", C.c_void_p), ("ws", C.c_void_p),
    ]


class GifFrame(C.Structure):        # vmm_gif_frame
    _fields_ = [("data_ofs", C.c_uint32), ("pal_ofs", C.c_uint32), ("px_ofs", C.c_uint32),
                ("x", C.c_uint16), ("y", C.c_uint16), ("w", C.c_uint16), ("h", C.c_uint16), ("pal_size", C.c_uint16),
                ("min_code", C.c_uint8), ("interlace", C.c_uint8), ("disposal", C.c_uint8), ("has_transp", C.c_uint8),
                ("transp", C.c_uint8), ("background", C.c_uint8), ("reserved", C.c_uint32)]


class GifInfo(C.Structure):         # vmm_gif_info
    _fields_ = [("width", C.c_uint16), ("height", C.c_uint16), ("n_frames", C.c_int32)]


class VmmError(RuntimeError):
    pass


ABI_VERSION = 8          # must equal vmm_abi_version() of the loaded library (bumped whenever a params struct or signature changes)


def _load() -> C.CDLL:
    # Where nvcc exists the build is always consulted: it is a no-op when the source digest matches the stamp, and it replaces a
    # stale .so (older csrc, drifted struct layouts) otherwise.  Ranks of one node serialise on a file lock.
    try:
        from . import build as _build
        if os.path.exists(_build.NVCC) or not os.path.exists(LIB_PATH):
            import fcntl
            os.makedirs(_build.OBJ, exist_ok=True)
            with open(os.path.join(_build.OBJ, ".lock"), "w") as lock:
                fcntl.flock(lock, fcntl.LOCK_EX)
                try:
                    _build.build()
                finally:
                    fcntl.flock(lock, fcntl.LOCK_UN)
    except Exception as e:  # noqa: BLE001
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"libvmm_sm100.so is missing at {LIB_PATH} and could not be built ({e}). "
                "This package has no CPU or PyTorch fallback: build it with "
                "`python -m videometamaterials_b200.build` where nvcc is available.") from e
        import warnings
        warnings.warn(f"videometamaterials_b200: could not verify / rebuild libvmm_sm100.so ({e}); loading the existing file")
    lib = C.CDLL(LIB_PATH)
    lib.vmm_last_error.restype = C.c_char_p
    lib.vmm_abi_version.restype = C.c_int
    lib.vmm_launch_count.restype = C.c_uint64
    if int(lib.vmm_abi_version()) != ABI_VERSION:
        raise ImportError(f"libvmm_sm100.so reports ABI version {int(lib.vmm_abi_version())}, this package expects {ABI_VERSION}: "
                          "rebuild it with `python -m videometamaterials_b200.build --force`")
    return lib


lib = _load()

_P, _I, _L, _F, _Z = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
# name -> argtypes (restype int unless listed in _RESTYPES); mirrors include/vmm.h one to one
_SIGNATURES = {
    "vmm_cgemm": [C.POINTER(CgemmParams), _P],
    "vmm_wgrad": [C.POINTER(WgradParams), _P],
    "vmm_colsum": [_P, _L, _I, _L, _I, _P, _P],
    "vmm_qkv_bwd": [_P, _P, _P, _P, _P, _L, _I, _P],
    "vmm_qkv_ln_bwd": [_P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _L, _I, _P],
    "vmm_gn_silu_fwd": [_P, _P, _P, _I, _I, _L, _I, _I, _P, _P, _P, _P, _F, _I, _P],
    "vmm_gn_silu_bwd_workspace": [_I, _I, _I],
    "vmm_gn_silu_bwd": [_P, _P, _P, _I, _I, _L, _I, _I, _P, _P, _P, _P, _F, _I, _P, _P, _P, _P, _P, _Z, _P],
    "vmm_ln_fwd": [_P, _P, _I, _L, _I, _P, _F, _P, _P],
    "vmm_ln_bwd": [_P, _P, _P, _P, _I, _L, _I, _P, _F, _P, _P],
    "vmm_tattn_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _I, _P],
    "vmm_ftattn_workspace": [_I],
    "vmm_ftattn_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _I, _I, _I, _I, _I, _I, _F, _P],
    "vmm_ftattn_ctas_per_sm": [],
    "vmm_ftattn_diag": [_P],
    "vmm_flattn_workspace": [_I],
    "vmm_flattn_fwd": [_P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _Z, _I, _I, _I, _I, _I, _I, _F, _F, _P],
    "vmm_cond_workspace": [_I, _I, _I, _I],
    "vmm_cond_fwd": [C.POINTER(CondParams), _P],
    "vmm_cond_bwd": [C.POINTER(CondParams), _P],
    "vmm_lattn_fwd": [_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "vmm_sattn_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "vmm_tattn_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _I, _P],
    "vmm_lattn_bwd": [_P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _P],
    "vmm_sattn_bwd": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "vmm_prep_input": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "vmm_loss": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    "vmm_cfg_x0": [_P, _P, _I, _F, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "vmm_abs_quantile_workspace": [_I],
    "vmm_abs_quantile": [_P, _I, _L, _L, _F, _F, _P, _P, _Z, _P],
    "vmm_posterior_step": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _L, _P],
    "vmm_axpby": [_P, _P, _F, _F, _F, _P, _L, _P],
    "vmm_gather_cast": [_P, _P, _P, _L, _I, _P],
    "vmm_adam_ema_step": [_P, _P, _P, _P, _P, _L, _F, _F, _F, _F, _I, _F, _I, _F, _P],
    "vmm_gif_scan": [_P, _Z, _I, C.POINTER(GifInfo), C.POINTER(GifFrame), _I],
    "vmm_gif_decode": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "vmm_dataset_items": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P],
}
_RESTYPES = {"vmm_abs_quantile_workspace": C.c_size_t, "vmm_gn_silu_bwd_workspace": C.c_size_t, "vmm_ftattn_workspace": C.c_size_t, "vmm_cond_workspace": C.c_size_t,
             "vmm_flattn_workspace": C.c_size_t}
for _name, _args in _SIGNATURES.items():
    _fn = getattr(lib, _name)     # AttributeError here == the .so is stale: rebuild it
    _fn.argtypes = _args
    _fn.restype = _RESTYPES.get(_name, C.c_int)

# every symbol include/vmm.h declares; tests assert that the library exports all of them
EXPORTS = ["vmm_last_error", "vmm_abi_version", "vmm_launch_count", *_SIGNATURES.keys()]


def check(rc: int, what: str = "vmm call") -> None:
    if rc != 0:
        raise VmmError(f"{what} failed ({rc}): {lib.vmm_last_error().decode()}")


_replayed_launches = 0      # kernel launches executed by CUDA-graph replays (the library only counts host-side launch calls)


def add_replayed_launches(n: int) -> None:
    global _replayed_launches
    _replayed_launches += int(n)


def launch_count() -> int:
    return int(lib.vmm_launch_count()) + _replayed_launches
