"""Drop-in for the reference's block_4_tv_helpers.py: the NumPy helpers (:17-46) keep their names, argument meaning
and float64 in/out, but run as CUDA kernels (C ABI admm_grad2d_host / admm_div2d_host / admm_kt_subgrad_host; fp64 on
the device, bit-identical to NumPy).  `isotropic_tv_on_vector` (:5-14) built a CVXPY expression; here it evaluates
the canonical isotropic TV of a concrete vector."""
from __future__ import annotations

import numpy as np

from admm_b200 import _native as nat


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _grad_forward_2d_from_vec(x_vec, N):
    """block_4_tv_helpers.py:17-23."""
    nat.require_cuda()
    x = _f64(x_vec).reshape(-1)
    if x.size != N * N:
        raise ValueError("x_vec must have N*N entries")
    gx, gy = np.empty((N, N)), np.empty((N, N))
    nat.check(nat.lib().admm_grad2d_host(N, x.ctypes.data, gx.ctypes.data, gy.ctypes.data), "admm_grad2d_host")
    return gx, gy


def _div_backward_2d_to_vec(px, py, N, exact_adjoint=False):
    """block_4_tv_helpers.py:25-35.  Default reproduces the reference (border rows/columns have the opposite sign of
    K^T); `exact_adjoint=True` gives the true adjoint the solver uses."""
    nat.require_cuda()
    px, py = _f64(px).reshape(-1), _f64(py).reshape(-1)
    out = np.empty(N * N)
    nat.check(nat.lib().admm_div2d_host(N, px.ctypes.data, py.ctypes.data, int(bool(exact_adjoint)), out.ctypes.data),
              "admm_div2d_host")
    return out


def kt_subgrad_isotropic_tv_from_x(x_vec, N, eps=1e-12, exact_adjoint=False):
    """block_4_tv_helpers.py:37-46."""
    nat.require_cuda()
    x = _f64(x_vec).reshape(-1)
    out = np.empty(N * N)
    nat.check(nat.lib().admm_kt_subgrad_host(N, x.ctypes.data, float(eps), int(bool(exact_adjoint)), out.ctypes.data,
                                             None), "admm_kt_subgrad_host")
    return out


def edge_map_from_vector(x_vec, N, normalize=True):
    """block_4_tv_helpers_with_plot.py:23-46: gradient magnitude image (optionally scaled to [0, 1])."""
    nat.require_cuda()
    x = _f64(x_vec).reshape(-1)
    mag = np.empty((N, N))
    nat.check(nat.lib().admm_kt_subgrad_host(N, x.ctypes.data, 1e-12, 0, None, mag.ctypes.data), "admm_kt_subgrad_host")
    if normalize:
        mx = mag.max()
        if mx > 0:
            mag = mag / mx
    return mag


def isotropic_tv_on_vector(x_vec, N):
    """Canonical isotropic TV value sum_k |(Dx_k, Dy_k)|_2 of a concrete vector (block_4_tv_helpers.py:5-14 builds the
    CVXPY expression; its Fortran-order pairing quirk is not reproduced)."""
    return float(edge_map_from_vector(x_vec, N, normalize=False).sum())
