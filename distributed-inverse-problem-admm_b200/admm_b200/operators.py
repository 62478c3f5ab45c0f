"""ODL-shaped, matrix-free parallel-beam operators backed by the sm_100a kernels.

`RayTransformCUDA` stands where the reference uses `odl.tomo.RayTransform` objects
(block_2_load_odl_data.py:51-56, Gen_Sino_Partitioned.py:133) *and* where it uses the dense matrices
`A_dense_list[i]` (`.shape`, `A @ x`, `A.T @ r`; block_6_admm_loop_ver2.py:26,145,193): the dense
materialisation (block_2_load_odl_data.py:68-96) is deliberately not reproduced.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import _native as nat
from .geometry import trig_table32


def _torch():
    import torch
    return torch


def _impl_name(impl):
    """ODL backend names -> the two discretisations this build has: "joseph" (= ASTRA `linear`, ODL's default 2-D
    parallel-beam backend when ASTRA is present) and "skimage" (rotate-and-sum bilinear)."""
    if impl in ("joseph", "astra_cuda", "astra_cpu", None):
        return "joseph"
    if impl == "skimage":
        return "skimage"
    raise ValueError(f"unknown impl {impl!r}")


class Plan:
    """Geometry + projector workspace of the nodes resident on one GPU (wraps `admm_plan`)."""

    def __init__(self, N, thetas, D=None, det_w=2.0, device=0, impl="joseph"):
        nat.require_cuda()
        self.impl = _impl_name(impl)
        self.N = int(N)
        self.D = int(D if D is not None else N)
        self.det_w = float(det_w)
        self.device = int(device)
        self.thetas = [np.asarray(t, dtype=np.float64).reshape(-1) for t in thetas]
        self.V = len(self.thetas)
        self.ang_ptr = np.zeros(self.V + 1, dtype=np.int32)
        self.ang_ptr[1:] = np.cumsum([len(t) for t in self.thetas])
        self.A = int(self.ang_ptr[-1])
        allth = np.concatenate(self.thetas) if self.A else np.zeros(0)
        c32, s32 = trig_table32(allth)
        self._c32, self._s32 = c32, s32
        L = nat.lib()
        h = L.admm_plan_create(self.N, self.D, self.det_w, self.V, self.ang_ptr.ctypes.data,
                               c32.ctypes.data, s32.ctypes.data, self.device)
        if not h:
            raise RuntimeError("admm_plan_create failed: " + L.admm_last_error().decode())
        self.handle = ctypes.c_void_p(h)
        self.n = self.N * self.N
        if self.impl == "skimage":
            nat.check(L.admm_plan_set(self.handle, nat.OPT_IMPL, 1), "admm_plan_set")
        self.part_floats = int(L.admm_plan_info(self.handle, nat.INFO_PART_FLOATS))

    def info(self, what):
        return int(nat.lib().admm_plan_info(self.handle, what))

    def close(self):
        if getattr(self, "handle", None):
            nat.lib().admm_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- device-pointer operator calls (torch CUDA float32 tensors) -------------------------------------
    def _stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream().cuda_stream)

    def forward(self, img, sino, node0=0, nodes=None, stride=None):
        nodes = self.V - node0 if nodes is None else nodes
        stride = self.n if stride is None else stride
        nat.check(nat.lib().admm_forward(self.handle, img.data_ptr(), stride, node0, nodes, sino.data_ptr(),
                                         self._stream()), "admm_forward")

    def adjoint(self, sino, img, prec=None, node0=0, nodes=None, stride=None):
        nodes = self.V - node0 if nodes is None else nodes
        stride = self.n if stride is None else stride
        nat.check(nat.lib().admm_adjoint(self.handle, sino.data_ptr(), prec.data_ptr() if prec is not None else None,
                                         img.data_ptr(), stride, node0, nodes, self._stream()), "admm_adjoint")

    def colnorm2(self, img, node0=0, nodes=None, stride=None):
        nodes = self.V - node0 if nodes is None else nodes
        stride = self.n if stride is None else stride
        nat.check(nat.lib().admm_colnorm2(self.handle, img.data_ptr(), stride, node0, nodes, self._stream()),
                  "admm_colnorm2")


class DensePlan(Plan):
    """Plan over explicit dense matrices (the reference's literal `A_dense_list` ndarrays, uploaded once): same
    interface as `Plan` with D = 1 and A = total matrix rows.  Tiny problems only -- 4 m_i n bytes per node."""

    def __init__(self, N, mats, device=0):
        nat.require_cuda()
        self.N, self.D, self.det_w, self.device = int(N), 1, 2.0, int(device)
        self.n = self.N * self.N
        mats = [np.ascontiguousarray(np.asarray(A, dtype=np.float32)) for A in mats]
        for A in mats:
            if A.ndim != 2 or A.shape[1] != self.n:
                raise ValueError(f"dense operator of shape {A.shape} does not map {self.N}x{self.N} images")
        self.V = len(mats)
        self.thetas = None
        self.ang_ptr = np.zeros(self.V + 1, dtype=np.int32)
        self.ang_ptr[1:] = np.cumsum([A.shape[0] for A in mats])
        self.A = int(self.ang_ptr[-1])
        L = nat.lib()
        h = L.admm_plan_create_dense(self.N, self.V, self.ang_ptr.ctypes.data, self.device)
        if not h:
            raise RuntimeError("admm_plan_create_dense failed: " + L.admm_last_error().decode())
        self.handle = ctypes.c_void_p(h)
        for i, A in enumerate(mats):
            nat.check(L.admm_plan_upload_dense(self.handle, i, A.ctypes.data), "admm_plan_upload_dense")
        self.part_floats = int(L.admm_plan_info(self.handle, nat.INFO_PART_FLOATS))


def make_plan(N, ops, D=None, det_w=2.0, device=0, impl="joseph"):
    """One plan over the operators of the nodes resident on a GPU: matrix-free (`RayTransformCUDA`: angle arrays) or
    dense (`DenseOperatorCUDA` / 2-D ndarrays)."""
    dense = [isinstance(o, DenseOperatorCUDA) or (isinstance(o, np.ndarray) and o.ndim == 2) for o in ops]
    if all(dense):
        return DensePlan(N, [o.matrix if isinstance(o, DenseOperatorCUDA) else o for o in ops], device)
    if any(dense):
        raise TypeError("mixing dense matrices and matrix-free operators in one A_dense_list is not supported")
    impls = {getattr(o, "impl", impl) for o in ops}
    if len(impls) > 1:
        raise TypeError("all node operators must use the same projector discretisation (impl)")
    return Plan(N, [np.asarray(o.angles if hasattr(o, "angles") else o, dtype=np.float64) for o in ops], D, det_w, device,
                impl=impls.pop())


class DenseOperatorCUDA:
    """A dense (m, n) matrix behind the operator interface the drop-ins use (`.shape`, `@`, `.T @`, `.colnorm2()`),
    evaluated by the dense kernels of libadmm_b200.so.  Stands for an `A_dense_list[i]` ndarray
    (block_2_load_odl_data.py:68-96)."""

    is_dense = True

    def __init__(self, A, N=None, device=0):
        self.matrix = np.ascontiguousarray(np.asarray(A, dtype=np.float32))
        if self.matrix.ndim != 2:
            raise ValueError("a dense operator is a 2-D matrix")
        self.shape = self.matrix.shape
        self.N = int(N) if N is not None else int(round(math.sqrt(self.shape[1])))
        if self.N * self.N != self.shape[1]:
            raise ValueError(f"{self.shape[1]} columns is not a square image")
        self.D, self.det_w, self.device = 1, 2.0, int(device)
        self._plan = None

    def plan(self):
        if self._plan is None:
            self._plan = DensePlan(self.N, [self.matrix], self.device)
        return self._plan

    def _forward_np(self, x):
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1))
        out = np.empty(self.shape[0], dtype=np.float32)
        nat.check(nat.lib().admm_forward_host(self.plan().handle, 0, x.ctypes.data, out.ctypes.data), "admm_forward_host")
        return out

    def _adjoint_np(self, y):
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float32).reshape(-1))
        out = np.empty(self.shape[1], dtype=np.float32)
        nat.check(nat.lib().admm_adjoint_host(self.plan().handle, 0, y.ctypes.data, out.ctypes.data), "admm_adjoint_host")
        return out

    def __matmul__(self, x):
        x = np.asarray(x)
        out = self._forward_np(x)
        return out.astype(np.float64) if x.dtype == np.float64 else out

    @property
    def T(self):
        op = self

        class _T:
            shape = (op.shape[1], op.shape[0])

            def __matmul__(self, y):
                y = np.asarray(y)
                out = op._adjoint_np(y)
                return out.astype(np.float64) if y.dtype == np.float64 else out
        return _T()

    def colnorm2(self):
        out = np.empty(self.shape[1], dtype=np.float32)
        nat.check(nat.lib().admm_colnorm2_host(self.plan().handle, 0, out.ctypes.data), "admm_colnorm2_host")
        return out.astype(np.float64)


# ---- ODL-shaped spaces / elements (block_2_load_odl_data.py:87-93,145-154; ADMM_Tomo_Only.py usage) -------
class Element:
    def __init__(self, space, arr):
        self.space = space
        self._a = arr

    def asarray(self):
        return self._a

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    @property
    def shape(self):
        return self._a.shape

    def set_zero(self):
        self._a[...] = 0
        return self

    def copy(self):
        return Element(self.space, self._a.copy())

    def __getitem__(self, k):
        return self._a[k]

    def __setitem__(self, k, v):
        self._a[k] = v

    def _coerce(self, o):
        return o._a if isinstance(o, Element) else o

    def __add__(self, o):
        return Element(self.space, self._a + self._coerce(o))

    __radd__ = __add__

    def __sub__(self, o):
        return Element(self.space, self._a - self._coerce(o))

    def __mul__(self, o):
        return Element(self.space, self._a * self._coerce(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return Element(self.space, self._a / self._coerce(o))

    def __neg__(self):
        return Element(self.space, -self._a)

    def norm(self):
        return float(np.sqrt(np.sum(self._a.astype(np.float64) ** 2) * self.space.cell_volume))

    def inner(self, o):
        return float(np.sum(self._a.astype(np.float64) * self._coerce(o)) * self.space.cell_volume)


class DiscreteSpace:
    """Uniformly discretised space with ODL's `.shape/.size/.element/.zero/.one` and cell-volume weighting."""

    def __init__(self, shape, cell_sides, dtype=np.float32):
        self.shape = tuple(int(s) for s in shape)
        self.size = int(np.prod(self.shape))
        self.cell_sides = tuple(float(c) for c in cell_sides)
        self.cell_volume = float(np.prod(self.cell_sides))
        self.dtype = np.dtype(dtype)

    def element(self, arr=None):
        if arr is None:
            return self.zero()
        if isinstance(arr, Element):
            arr = arr.asarray()
        return Element(self, np.array(arr, dtype=self.dtype).reshape(self.shape))

    def zero(self):
        return Element(self, np.zeros(self.shape, dtype=self.dtype))

    def one(self):
        return Element(self, np.ones(self.shape, dtype=self.dtype))


class _Functional:
    """`op.adjoint` / `op.T` view: callable on sinogram elements / `@` on flat vectors."""

    def __init__(self, op, scale):
        self._op = op
        self._scale = scale
        self.domain, self.range = op.range, op.domain
        self.shape = (op.shape[1], op.shape[0])

    def __call__(self, y):
        arr = y.asarray() if isinstance(y, Element) else np.asarray(y)
        out = self._op._adjoint_np(arr) * self._scale
        return Element(self.range, out.astype(self.range.dtype).reshape(self.range.shape))

    def __matmul__(self, y):
        y = np.asarray(y)
        out = self._op._adjoint_np(y) * self._scale
        return out.astype(np.float64 if y.dtype == np.float64 else np.float32).reshape(-1)

    @property
    def T(self):
        return self._op

    @property
    def adjoint(self):
        return self._op


class RayTransformCUDA:
    """A_i : (N, N) image on [-1,1]^2 -> (M_i, D) sinogram, 2-D parallel beam, Joseph discretisation
    (SURVEY App. C), matrix-free on the GPU.

    ODL shape: `op(x)`, `op.adjoint(y)` (ODL-weighted: (w_Y / w_X) A^T), `op.domain`, `op.range`.
    Dense-matrix shape: `op.shape == (M_i*D, N*N)`, `op @ x`, `op.T @ r` (plain transpose, what CG uses),
    `op.colnorm2()` (= np.sum(A*A, axis=0), block_3_graph_and_precisions.py:22).
    """

    def __init__(self, N, theta, D=None, det_w=2.0, device=0, impl="joseph", angle_cell=None):
        self.impl = _impl_name(impl)     # "skimage": the rotate-and-sum variant (csrc/rotsum.cu), otherwise Joseph
        self.N = int(N)
        self.D = int(D if D is not None else N)
        self.det_w = float(det_w)
        self.device = int(device)
        self.theta = np.asarray(theta, dtype=np.float64).reshape(-1)
        self.nang = len(self.theta)
        self.shape = (self.nang * self.D, self.N * self.N)
        h = 2.0 / self.N
        self.domain = DiscreteSpace((self.N, self.N), (h, h))
        # angular cell of the range space (it weights `.adjoint`): pi / m_k for a node that covers the whole half
        # circle (`uniform_partition(0, pi, m_k)`, block_2_load_odl_data.py:51), pi / angles_total for a node that
        # holds a contiguous block of the aggregate grid -- the caller passes it
        dth = float(angle_cell) if angle_cell is not None else math.pi / max(self.nang, 1)
        self.range = DiscreteSpace((self.nang, self.D), (dth, self.det_w / self.D))
        self._plan = None

    # geometry accessors used by the solver to build one plan over many nodes
    @property
    def angles(self):
        return self.theta

    def plan(self):
        if self._plan is None:
            self._plan = Plan(self.N, [self.theta], self.D, self.det_w, self.device, impl=self.impl)
        return self._plan

    def _forward_np(self, x):
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1))
        if x.size != self.shape[1]:
            raise ValueError(f"expected {self.shape[1]} image values, got {x.size}")
        out = np.empty(self.shape[0], dtype=np.float32)
        nat.check(nat.lib().admm_forward_host(self.plan().handle, 0, x.ctypes.data, out.ctypes.data),
                  "admm_forward_host")
        return out

    def _adjoint_np(self, y):
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float32).reshape(-1))
        if y.size != self.shape[0]:
            raise ValueError(f"expected {self.shape[0]} sinogram values, got {y.size}")
        out = np.empty(self.shape[1], dtype=np.float32)
        nat.check(nat.lib().admm_adjoint_host(self.plan().handle, 0, y.ctypes.data, out.ctypes.data),
                  "admm_adjoint_host")
        return out

    def __call__(self, x):
        arr = x.asarray() if isinstance(x, Element) else np.asarray(x)
        return Element(self.range, self._forward_np(arr).reshape(self.range.shape))

    def __matmul__(self, x):
        x = np.asarray(x)
        out = self._forward_np(x)
        return out.astype(np.float64) if x.dtype == np.float64 else out

    @property
    def T(self):
        return _Functional(self, 1.0)

    @property
    def adjoint(self):
        """ODL adjoint w.r.t. the weighted inner products: (w_Y / w_X) A^T."""
        return _Functional(self, self.range.cell_volume / self.domain.cell_volume)

    def colnorm2(self):
        out = np.empty(self.shape[1], dtype=np.float32)
        nat.check(nat.lib().admm_colnorm2_host(self.plan().handle, 0, out.ctypes.data), "admm_colnorm2_host")
        return out.astype(np.float64)


def stack_operators(ops):
    """Aggregate operator = vstack of node operators (block_2_load_odl_data.py:58-63 intent)."""
    N, D, det_w = ops[0].N, ops[0].D, ops[0].det_w
    theta = np.concatenate([o.theta for o in ops])
    return RayTransformCUDA(N, theta, D, det_w, ops[0].device, impl=ops[0].impl, angle_cell=ops[0].range.cell_sides[0])
