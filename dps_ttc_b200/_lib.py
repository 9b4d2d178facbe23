"""ctypes binding of libdpsttc.so (include/dpsttc.h).  Plumbing only: torch owns every buffer and
the stream; this module turns tensors into (pointer, stride) arguments and error codes into
exceptions.  There is NO fallback: if the library is missing or a kernel fails, it raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DPSTTC_LIB") or os.path.join(_HERE, "libdpsttc.so")  # env override: experiment builds
CSRC_DIR = os.path.join(_HERE, "csrc")

DPS_COEF_NORM = 1
DPS_COEF_NORM_SQ = 2
DPS_COEF_GLOBAL_NORM = 3
OP_KINDS = {1: "inpainting", 2: "blur_separable", 3: "blur_sparse", 4: "resize", 5: "phase"}


class DpsError(RuntimeError):
    pass


class Source(C.Structure):
    _fields_ = [("x", C.c_void_p), ("eps", C.c_void_p), ("x_stride", C.c_int64), ("eps_stride", C.c_int64),
                ("c1", C.c_float), ("c2", C.c_float), ("clip", C.c_int32), ("pad_", C.c_int32)]


class StepConstsC(C.Structure):
    _fields_ = [("p1", C.c_float), ("p2", C.c_float), ("max_log", C.c_float), ("min_log", C.c_float),
                ("ddim_sa", C.c_float), ("ddim_sb", C.c_float), ("ddim_sigma", C.c_float),
                ("noise_on", C.c_int32), ("var_mode", C.c_int32), ("mean_mode", C.c_int32)]


class UpdateExt(C.Structure):
    _fields_ = [("partials", C.c_void_p), ("P", C.c_int32), ("coef_mode", C.c_int32), ("scale", C.c_float),
                ("use_philox", C.c_int32), ("l2_out", C.c_void_p), ("philox_seed", C.c_uint64), ("philox_step", C.c_uint64),
                ("particle_offset", C.c_int64)]


class OperatorInfo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("out_C", C.c_int32), ("out_H", C.c_int32), ("out_W", C.c_int32),
                ("partials_per_particle", C.c_int32), ("aux_floats_per_particle", C.c_int64),
                ("taps", C.c_int32), ("guidance_partials", C.c_int32)]


_P = C.c_void_p
_I, _L, _F = C.c_int, C.c_int64, C.c_float
# name -> (restype, argtypes); must list every symbol include/dpsttc.h declares (tests check this)
SIGNATURES = {
    "dps_last_error": (C.c_char_p, []),
    "dps_version": (_I, []),
    "dps_compiled_sm": (_I, []),
    "dps_device_sm": (_I, [C.POINTER(C.c_int)]),
    "dps_launch_count": (_L, []),
    "dps_launch_count_reset": (None, []),
    "dps_x0_from_eps": (_I, [C.POINTER(Source), _P, _I, _L, _P]),
    "dps_posterior_update_ddpm": (_I, [C.POINTER(Source), _P, _L, _P, _P, _L, _P, C.POINTER(StepConstsC), _P, _P, _P,
                                       _I, _L, _P]),
    "dps_posterior_update_ddim": (_I, [C.POINTER(Source), _P, _P, _L, _P, C.POINTER(StepConstsC), _P, _P, _P, _I, _L,
                                       _P]),
    "dps_posterior_update_ddpm_ext": (_I, [C.POINTER(Source), _P, _L, _P, _P, _L, _P, C.POINTER(StepConstsC),
                                           C.POINTER(UpdateExt), _P, _I, _L, _P]),
    "dps_posterior_update_ddim_ext": (_I, [C.POINTER(Source), _P, _P, _L, _P, C.POINTER(StepConstsC), C.POINTER(UpdateExt), _P,
                                           _I, _L, _P]),
    "dps_guidance_grad": (_I, [_P, _L, _P, _F, _F, _P, _I, _L, _P]),
    "dps_apply_gradient": (_I, [_P, _P, _L, _P, _I, _L, _P]),
    "dps_q_sample": (_I, [_P, _P, _F, _F, _P, _L, _P]),
    "dps_particle_sqdiff": (_I, [_P, _L, _P, _L, _I, _L, _P, _I, _P]),
    "dps_operator_create_inpainting": (_I, [_P, _I, _I, _I, C.POINTER(_P)]),
    "dps_operator_create_blur": (_I, [_P, _I, _I, _I, _I, _I, C.POINTER(_P)]),
    "dps_operator_create_resize": (_I, [_P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _I, C.POINTER(_P)]),
    "dps_operator_create_phase": (_I, [_I, _I, _I, _I, C.POINTER(_P)]),
    "dps_operator_destroy": (None, [_P]),
    "dps_operator_get_info": (_I, [_P, C.POINTER(OperatorInfo)]),
    "dps_operator_forward": (_I, [_P, C.POINTER(Source), _P, _L, _P, _P, _P, _I, _P]),
    "dps_operator_adjoint": (_I, [_P, _P, _P, C.POINTER(Source), _P, _L, _P, _L, _P, _I, _P]),
    "dps_operator_guidance": (_I, [_P, C.POINTER(Source), _P, _L, _P, _P, _L, _P, _P, _I, _P]),
    "dps_operator_project": (_I, [_P, _P, _L, _P, _L, _I, _P, _L, _P, _I, _P]),
    "dps_particle_norms": (_I, [_P, _I, _I, _P, _P, _P]),
    "dps_guidance_coef": (_I, [_P, _I, _I, _I, _F, _P, _P, _P]),
    "dps_particle_logweights": (_I, [_P, _P, _I, _F, _F, _I, _F, _I, _P, _P]),
    "dps_weights_cdf": (_I, [_P, _I, _I, _P, _P, _P, _P, _P]),
    "dps_ancestors_multinomial": (_I, [_P, _I, _P, _I, _P, _P, _P]),
    "dps_ancestors_systematic": (_I, [_P, _I, _P, _I, _P, _P, _P]),
    "dps_gather_particles": (_I, [_P, _P, _P, _I, _L, _P]),
    "dps_gather_particles_p2p": (_I, [_P, _I, _P, _P, _I, _L, _P]),
    "dps_exchange_particles_p2p": (_I, [_P, _P, _I, _I, C.c_uint32, _L, _I, _P, _P, _I, _L, _P]),
    "dps_argmin": (_I, [_P, _I, _P, _P, _P]),
    "dps_broadcast_particle": (_I, [_P, _P, _P, _I, _L, _P]),
}

_lock = threading.Lock()
_lib = None


def build(verbose: bool = False) -> str:
    """Compile libdpsttc.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if res.returncode != 0:
        raise DpsError("building libdpsttc.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise DpsError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               f"or `make -C {CSRC_DIR}`. dps_ttc_b200 has no CPU/PyTorch fallback.")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().dps_last_error().decode(errors="replace")
        raise DpsError(f"{what or 'libdpsttc'} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(lib().dps_launch_count())


def reset_launch_count() -> None:
    lib().dps_launch_count_reset()


# ------------------------------------------------------------------------------------------------
# tensor → argument helpers
# ------------------------------------------------------------------------------------------------
def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise DpsError(f"{name} must be a CUDA tensor: dps_ttc_b200 kernels have no CPU path")
    if t.dtype != torch.float32:
        raise DpsError(f"{name} must be float32, got {t.dtype}")
    return t


def particle_view(t: torch.Tensor, name: str):
    """(pointer, particle stride) of an (N, ...) fp32 CUDA tensor whose trailing dims are dense.
    Channel-slice views such as model_output[:, :3] qualify without a copy."""
    require_cuda_f32(t, name)
    inner = 1
    for size, stride in zip(reversed(t.shape[1:]), reversed(t.stride()[1:])):
        if size != 1 and stride != inner:
            raise DpsError(f"{name}: trailing dimensions must be contiguous (got strides {t.stride()})")
        inner *= size
    stride0 = t.stride(0) if t.shape[0] > 1 else inner
    if t.data_ptr() % 16 or stride0 % 4:
        raise DpsError(f"{name}: needs 16-byte alignment and a particle stride that is a multiple of 4")
    return t.data_ptr(), stride0


def dense(t: torch.Tensor, name: str) -> torch.Tensor:
    require_cuda_f32(t, name)
    return t if t.is_contiguous() else t.contiguous()


def make_source(x: torch.Tensor, eps=None, c1: float = 1.0, c2: float = 0.0, clip: bool = False) -> Source:
    xp, xs = particle_view(x, "x")
    s = Source()
    s.x, s.x_stride = xp, xs
    if eps is not None:
        if eps.shape != x.shape:
            raise DpsError(f"eps shape {tuple(eps.shape)} != x shape {tuple(x.shape)}")
        ep, es = particle_view(eps, "eps")
        s.eps, s.eps_stride = ep, es
    else:
        s.eps, s.eps_stride = None, 0
    s.c1, s.c2, s.clip, s.pad_ = float(c1), float(c2), int(bool(clip)), 0
    return s


def make_consts(k, var_mode: int = 0, max_log=None) -> StepConstsC:
    c = StepConstsC()
    c.p1, c.p2 = k.p1, k.p2
    c.max_log = k.max_log if max_log is None else max_log
    c.min_log = k.min_log
    c.ddim_sa, c.ddim_sb, c.ddim_sigma = k.ddim_sa, k.ddim_sb, k.ddim_sigma
    c.noise_on, c.var_mode = int(k.noise_on), int(var_mode)
    c.mean_mode = int(getattr(k, "mean_mode", 0))
    return c


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()
