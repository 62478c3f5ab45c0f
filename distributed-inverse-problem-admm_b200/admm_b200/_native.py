"""ctypes binding of libadmm_b200.so (C ABI in include/admm_b200.h).

There is no CPU fallback: if the shared library is missing it is built with nvcc for sm_100a; if a compute
entry point is called without a CUDA device it raises.  PyTorch only supplies device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)                      # distributed-inverse-problem-admm_b200/
CSRC = os.path.join(_ROOT, "csrc")
LIB_PATH = os.path.join(_ROOT, "libadmm_b200.so")
HEADER = os.path.join(os.path.dirname(_ROOT), "include", "admm_b200.h")
SOURCES = ["api.cu", "projector.cu", "solver_kernels.cu", "tv_helpers.cu", "dense.cu", "rotsum.cu", "pixel_masks.cu", "tma.cu", "pdhg.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

NSCAL = 16
S_RR0, S_RR1, S_PHP, S_RHP, S_HPHP, S_TV, S_GN2, S_IMG, S_MSE = range(9)
INFO_N, INFO_D, INFO_V, INFO_A, INFO_PART_FLOATS, INFO_FWD_SPAN, INFO_FWD_NREC, INFO_BACK_SPAN, INFO_WS_BYTES = range(9)


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> libadmm_b200.so (in-tree)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


class State(ctypes.Structure):
    """struct admm_state (include/admm_b200.h)."""
    _fields_ = [(k, ctypes.c_void_p) for k in
                ("x", "r", "r1", "p0", "p1", "hp", "rhs0", "tvterm", "atb", "w0", "w1", "rhoD_vec", "rhoD_s", "prec",
                 "xtrue", "q", "ax", "b", "scal", "part", "counter")] + [
        ("stride", ctypes.c_longlong), ("rho", ctypes.c_float), ("lam", ctypes.c_float), ("mu", ctypes.c_float),
        ("q_uniform", ctypes.c_float), ("w_parity", ctypes.c_int), ("fuse_pupdate", ctypes.c_int),
        ("defer_tv", ctypes.c_int), ("reuse_ax", ctypes.c_int), ("ctl", ctypes.c_void_p), ("masked", ctypes.c_int),
        ("carry_r", ctypes.c_int), ("reuse_r", ctypes.c_int), ("iter_dev", ctypes.c_void_p),
        ("hist_stride", ctypes.c_longlong), ("accept_mode", ctypes.c_int), ("max_tighten", ctypes.c_int),
        ("eps_target", ctypes.c_double), ("skip_mse", ctypes.c_int)]


EDGE_FIELDS = ("xi", "xj", "yi", "yj", "z", "ai", "aj", "Wi", "Wj", "qij", "qji", "vi", "vj")  # struct admm_edge (u64 each)
PACK_FIELDS = ("x", "y", "out")                                                    # struct admm_pack_item

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if _stale():
        build()
    L = ctypes.CDLL(LIB_PATH)
    vp, i, ll, d = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_double
    L.admm_version.restype = i
    L.admm_abi_sizeof.restype = ll
    L.admm_abi_sizeof.argtypes = [i]
    L.admm_last_error.restype = ctypes.c_char_p
    L.admm_device_count.restype = i
    L.admm_launch_count.restype = ll
    L.admm_plan_create.restype = vp
    L.admm_plan_create.argtypes = [i, i, d, i, vp, vp, vp, i]
    L.admm_plan_create_dense.restype = vp
    L.admm_plan_create_dense.argtypes = [i, i, vp, i]
    L.admm_plan_upload_dense.restype = i
    L.admm_plan_upload_dense.argtypes = [vp, i, vp]
    L.admm_plan_destroy.argtypes = [vp]
    L.admm_plan_destroy.restype = None
    L.admm_plan_info.restype = ll
    L.admm_plan_info.argtypes = [vp, i]
    L.admm_plan_set.restype = i
    L.admm_plan_set.argtypes = [vp, i, ll]
    L.admm_forward.argtypes = [vp, vp, ll, i, i, vp, vp]
    L.admm_adjoint.argtypes = [vp, vp, vp, vp, ll, i, i, vp]
    L.admm_colnorm2.argtypes = [vp, vp, ll, i, i, vp]
    L.admm_forward_host.argtypes = [vp, i, vp, vp]
    L.admm_adjoint_host.argtypes = [vp, i, vp, vp]
    L.admm_colnorm2_host.argtypes = [vp, i, vp]
    L.admm_colnorm2_host.restype = i
    L.admm_rhs0.argtypes = [vp, ctypes.POINTER(State), vp, vp, vp, vp, i, i, vp]
    L.admm_x_update.argtypes = [vp, ctypes.POINTER(State), i, i, i, i, vp]
    L.admm_tv_pass.argtypes = [vp, ctypes.POINTER(State), i, i, i, vp]
    L.admm_accept.argtypes = [vp, ctypes.POINTER(State), i, i, d, i, i, vp]
    L.admm_edge_update.argtypes = [vp, ctypes.POINTER(State), vp, i, vp, vp]
    L.admm_pixel_masks.argtypes = [i, ll, i, i, i, vp, vp, vp, vp]
    L.admm_pixel_masks.restype = i
    L.admm_pack.argtypes = [vp, vp, i, vp]
    L.admm_push_copy.argtypes = [vp, vp, i, vp]
    L.admm_finalize.argtypes = [vp, ctypes.POINTER(State), vp, vp, vp, vp, i, i, vp, vp, vp, vp, i, vp, vp]
    L.admm_grad2d_host.argtypes = [i, vp, vp, vp]
    L.admm_div2d_host.argtypes = [i, vp, vp, i, vp]
    L.admm_kt_subgrad_host.argtypes = [i, vp, d, i, vp, vp]
    L.admm_ipc_alloc.argtypes = [ll, ctypes.POINTER(vp), vp]
    L.admm_ipc_open.argtypes = [vp, ctypes.POINTER(vp)]
    L.admm_ipc_close.argtypes = [vp]
    L.admm_ipc_free.argtypes = [vp]
    f = ctypes.c_float
    L.admm_pdhg_dual.argtypes = [vp, vp, ll, vp, vp, vp, vp, vp, f, f, i, i, vp]
    L.admm_pdhg_primal.argtypes = [vp, vp, vp, ll, vp, vp, vp, vp, vp, f, f, i, i, vp]
    L.admm_pdhg_normal.argtypes = [vp, vp, ll, vp, vp, vp, i, i, vp]
    L.admm_pdhg_combine.argtypes = [vp, vp, ll, vp, vp, vp, i, vp]
    L.admm_pdhg_sums.argtypes = [vp, vp, ll, vp, vp, vp, vp, i, i, vp]
    for name in ("admm_pdhg_dual", "admm_pdhg_primal", "admm_pdhg_normal", "admm_pdhg_combine", "admm_pdhg_sums"):
        getattr(L, name).restype = i
    L.admm_profile_enable.argtypes = [i]
    L.admm_profile_read.argtypes = [vp, vp]
    for name in ("admm_ipc_alloc", "admm_ipc_open", "admm_ipc_close", "admm_ipc_free", "admm_grad2d_host", "admm_div2d_host", "admm_kt_subgrad_host", "admm_profile_enable", "admm_profile_read", "admm_forward", "admm_adjoint", "admm_colnorm2", "admm_forward_host", "admm_adjoint_host",
                 "admm_rhs0", "admm_x_update", "admm_tv_pass", "admm_accept", "admm_edge_update", "admm_pack", "admm_push_copy", "admm_finalize"):
        getattr(L, name).restype = i
    if L.admm_abi_sizeof(0) != ctypes.sizeof(State):
        raise RuntimeError(f"libadmm_b200.so admm_state is {L.admm_abi_sizeof(0)} bytes, the binding's mirror "
                           f"{ctypes.sizeof(State)}: stale library?")
    _lib = L
    return L


EXPORTS = ("admm_version", "admm_abi_sizeof", "admm_last_error", "admm_device_count", "admm_plan_create",
           "admm_plan_create_dense", "admm_plan_upload_dense", "admm_plan_destroy",
           "admm_plan_info", "admm_plan_set", "admm_forward", "admm_adjoint", "admm_colnorm2", "admm_forward_host",
           "admm_adjoint_host", "admm_colnorm2_host", "admm_rhs0", "admm_x_update", "admm_edge_update", "admm_pack", "admm_push_copy", "admm_finalize", "admm_pixel_masks",
           "admm_tv_pass", "admm_accept", "admm_launch_count", "admm_profile_enable", "admm_profile_read", "admm_grad2d_host",
           "admm_div2d_host", "admm_kt_subgrad_host", "admm_ipc_alloc", "admm_ipc_open", "admm_ipc_close", "admm_ipc_free",
           "admm_pdhg_dual", "admm_pdhg_primal", "admm_pdhg_normal", "admm_pdhg_combine", "admm_pdhg_sums")

OPT_PACK_BLOCKS = 0
OPT_IMPL = 1

KC_NAMES = ("fwd", "fwd_reduce", "back_plain", "back_hp", "back_resid0", "colnorm2", "tv", "cg_update", "p_update",
            "sino_axpy", "sino_resid", "rhs0", "edge", "pack", "finalize", "fwd_fused", "accept")


_profiling = False


def profile_enable(on: bool) -> None:
    global _profiling
    check(lib().admm_profile_enable(1 if on else 0), "admm_profile_enable")
    _profiling = bool(on)


def profiling() -> bool:
    """True while the per-launch CUDA-event profiler is on (event records cannot be captured into a CUDA graph)."""
    return _profiling


def profile_read() -> dict:
    """{kernel class: (launches, total ms)} since the last read (synchronises the device)."""
    n = len(KC_NAMES)
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_longlong * n)()
    check(lib().admm_profile_read(ms, cnt), "admm_profile_read")
    return {KC_NAMES[k]: (int(cnt[k]), float(ms[k])) for k in range(n) if cnt[k]}


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = lib().admm_last_error()
        raise RuntimeError(f"libadmm_b200 {what} failed ({code}): {msg.decode() if msg else ''}")


def require_cuda() -> None:
    if lib().admm_device_count() < 1:
        raise RuntimeError("libadmm_b200: no CUDA device visible -- this path is CUDA-only (no CPU fallback)")


def launch_count() -> int:
    return int(lib().admm_launch_count())
