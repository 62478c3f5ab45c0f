"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

fp64 CPU restatement of the reference's decentralized TV-ADMM tomography path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import this module; the product package never does.

PARITY STATUS: *parity unpinned* at the ODL ``RayTransform`` and CVXPY/SCS boundaries (neither is
vendored under /root/reference, installable here, nor covered by any golden vector in the reference).
Pinned against the reference's own NumPy code (see tests/golden/make_golden.py): block_4 TV helpers,
block_3 ``make_precisions``, the block_6_ver2 z / y / residual / history loop, the block_2 angle split,
``ConstIm`` and ``psnr``.  The projector is anchored independently by closed-form Radon transforms of
ellipses and by exact-transpose checks (tests/test_oracle.py).

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_SRC_PATH = os.path.join(_HERE, "admm_oracle.c")
_lib = None


def build(force: bool = False) -> str:
    """Compile admm_oracle.c -> liboracle.so (gcc -O3 -fopenmp)."""
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(_SRC_PATH)):
        return _LIB_PATH
    cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-o", _LIB_PATH, _SRC_PATH, "-lm"]
    try:
        subprocess.run(cmd, check=True, capture_output=True, text=True)
    except subprocess.CalledProcessError:
        cmd.remove("-march=native")
        subprocess.run(cmd, check=True, capture_output=True, text=True)
    return _LIB_PATH


_DP = ctypes.POINTER(ctypes.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i, d, l = ctypes.c_int, ctypes.c_double, ctypes.c_long
        L.orc_num_threads.restype = i
        L.orc_set_num_threads.argtypes = [i]
        L.orc_forward.argtypes = [_DP, i, i, d, _DP, _DP, i, _DP]
        L.orc_adjoint.argtypes = [_DP, i, i, d, _DP, _DP, i, _DP]
        L.orc_colnorm2.argtypes = [i, i, d, _DP, _DP, i, _DP]
        L.orc_forward_rs.argtypes = [_DP, i, i, d, _DP, _DP, i, _DP]
        L.orc_adjoint_rs.argtypes = [_DP, i, i, d, _DP, _DP, i, _DP]
        L.orc_colnorm2_rs.argtypes = [i, i, d, _DP, _DP, i, _DP]
        L.orc_grad.argtypes = [_DP, i, _DP, _DP]
        L.orc_gradT.argtypes = [_DP, _DP, i, _DP]
        L.orc_div_reference.argtypes = [_DP, _DP, i, _DP]
        L.orc_tv_value.argtypes = [_DP, i]
        L.orc_tv_value.restype = d
        L.orc_x_update.argtypes = [i, i, d, _DP, _DP, i, d, _DP, _DP, d, d, d, i, i, _DP, _DP, _DP,
                                   _DP, _DP, _DP]
        L.orc_edge_update.argtypes = [l, _DP, _DP, _DP, _DP, _DP, _DP, _DP, _DP, _DP, d, _DP]
        L.orc_accum_cons.argtypes = [l, d, _DP, d, _DP, _DP, _DP]
        _lib = L
        L.orc_set_num_threads(default_threads())
    return _lib


def default_threads() -> int:
    """Host threads the oracle uses: $ORACLE_THREADS, else all cores (one left free on small shared hosts,
    where a descheduled OpenMP worker stalls every barrier)."""
    env = os.environ.get("ORACLE_THREADS")
    if env:
        return max(1, int(env))
    nc = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return nc - 1 if 2 < nc <= 8 else nc


def num_threads() -> int:
    return int(lib().orc_num_threads())


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_DP)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ----------------------------------------------------------------------------------------------
# Geometry (a1).  block_2_load_odl_data.py:31-38 (split), :51-54 (per-node geometry), App. B-1.
# ----------------------------------------------------------------------------------------------
def default_angles_total(N: int) -> int:
    """block_2_load_odl_data.py:31-33."""
    return max(180, 3 * N)


def angle_split(angles_total: int, num_nodes: int) -> list:
    """block_2_load_odl_data.py:36-38 -- integer, bit-exact."""
    per = [angles_total // num_nodes] * num_nodes
    for i in range(angles_total % num_nodes):
        per[i] += 1
    return per


def node_angles(angles_total: int, num_nodes: int, partition: str = "contiguous") -> list:
    """Angles (radians, fp64) of every node.

    ``contiguous``: node k takes the k-th contiguous block of the aggregate midpoint grid
    theta_a = (a+.5) pi / M  (north_star "angle-partitioned"; aggregate == vstack of the nodes).
    ``reference_literal``: block_2_load_odl_data.py:51 as shipped -- every node gets
    ``uniform_partition(0, pi, m_k)`` midpoints, i.e. the full half circle at coarser sampling.
    """
    per = angle_split(angles_total, num_nodes)
    out, start = [], 0
    for m_k in per:
        if partition == "contiguous":
            idx = np.arange(start, start + m_k, dtype=np.float64)
            out.append((idx + 0.5) * math.pi / angles_total)
        elif partition == "reference_literal":
            out.append((np.arange(m_k, dtype=np.float64) + 0.5) * math.pi / m_k)
        else:
            raise ValueError(partition)
        start += m_k
    return out


def trig_table(theta) -> tuple:
    """(cos, sin) computed in fp64, rounded to fp32, returned as fp64 (oracle and
    kernels use bit-identical trig values and take identical dominant-axis decisions)."""
    theta = np.asarray(theta, dtype=np.float64)
    c = np.cos(theta).astype(np.float32).astype(np.float64)
    s = np.sin(theta).astype(np.float32).astype(np.float64)
    return np.ascontiguousarray(c), np.ascontiguousarray(s)


def node_to_gpu(num_nodes: int, num_gpus: int) -> list:
    """deterministic contiguous blocks gpu(i) = (i*G)//V."""
    return [(i * num_gpus) // num_nodes for i in range(num_nodes)]


# ----------------------------------------------------------------------------------------------
# Projector (a2, a3, a7)
# ----------------------------------------------------------------------------------------------
class JosephOperator:
    """Matrix-free A_i: (N,N) image -> (M_i, D) sinogram, with .T and column norms."""

    def __init__(self, N, theta, D=None, det_w=2.0, impl="joseph"):
        """impl: "joseph" (the canonical contract) or "skimage" (rotate-and-sum with bilinear
        interpolation, the backend Gen_Sino_Partitioned.py:133 pins; see admm_oracle.c orc_forward_rs)."""
        self.N = int(N)
        self.D = int(D if D is not None else N)
        self.det_w = float(det_w)
        self.theta = np.asarray(theta, dtype=np.float64)
        self.c, self.s = trig_table(self.theta)
        self.nang = len(self.theta)
        self.shape = (self.nang * self.D, self.N * self.N)
        if impl not in ("joseph", "skimage"):
            raise ValueError(impl)
        self.impl = impl

    def forward(self, x):
        x = _f64(np.asarray(x).reshape(-1))
        out = np.empty(self.nang * self.D)
        fn = lib().orc_forward if self.impl == "joseph" else lib().orc_forward_rs
        fn(_p(x), self.N, self.D, self.det_w, _p(self.c), _p(self.s), self.nang, _p(out))
        return out

    def adjoint(self, q):
        q = _f64(np.asarray(q).reshape(-1))
        out = np.empty(self.N * self.N)
        fn = lib().orc_adjoint if self.impl == "joseph" else lib().orc_adjoint_rs
        fn(_p(q), self.N, self.D, self.det_w, _p(self.c), _p(self.s), self.nang, _p(out))
        return out

    def colnorm2(self):
        out = np.empty(self.N * self.N)
        fn = lib().orc_colnorm2 if self.impl == "joseph" else lib().orc_colnorm2_rs
        fn(self.N, self.D, self.det_w, _p(self.c), _p(self.s), self.nang, _p(out))
        return out

    def __matmul__(self, x):
        return self.forward(x)

    @property
    def T(self):
        return _Transposed(self)

    def dense(self):
        """Dense (m, n) matrix by probing unit vectors with the pure-NumPy forward (small N only);
        layout of block_2_load_odl_data.py:87-96 (row = angle*D + det, col = ix*N + iy)."""
        n = self.N * self.N
        cols = [np_forward(np.eye(1, n, j).reshape(self.N, self.N), self.c, self.s, self.D, self.det_w).reshape(-1)
                for j in range(n)]
        return np.stack(cols, axis=1)


class _Transposed:
    def __init__(self, op):
        self.op = op
        self.shape = (op.shape[1], op.shape[0])

    def __matmul__(self, q):
        return self.op.adjoint(q)


def np_forward_rs(X, c, s, D, det_w=2.0):
    """Pure-NumPy twin of orc_forward_rs: rotate-and-sum with bilinear interpolation (literal statement)."""
    X = np.asarray(X, dtype=np.float64)
    N = X.shape[0]
    h, ds, smin, x0 = 2.0 / N, det_w / D, -0.5 * det_w, -1.0 + 1.0 / N
    P = int(math.ceil(math.sqrt(2.0) * N))
    sj = smin + (np.arange(D) + 0.5) * ds
    tk = (np.arange(P) - 0.5 * (P - 1)) * h
    out = np.zeros((len(c), D))
    Xp = np.zeros((N + 2, N + 2))
    Xp[1:-1, 1:-1] = X                      # zero ring: taps outside the image contribute 0
    for a, (ca, sa) in enumerate(zip(c, s)):
        px = ((sj[:, None] * ca - tk[None, :] * sa) - x0) / h
        py = ((sj[:, None] * sa + tk[None, :] * ca) - x0) / h
        i0, j0 = np.floor(px).astype(np.int64), np.floor(py).astype(np.int64)
        fx, fy = px - i0, py - j0
        acc = np.zeros_like(px)
        for di, wx in ((0, 1.0 - fx), (1, fx)):
            for dj, wy in ((0, 1.0 - fy), (1, fy)):
                ii, jj = np.clip(i0 + di + 1, 0, N + 1), np.clip(j0 + dj + 1, 0, N + 1)
                acc += wx * wy * Xp[ii, jj]
        out[a] = acc.sum(axis=1) * h
    return out


def np_forward(X, c, s, D, det_w=2.0):
    """Pure-NumPy twin of orc_forward (literal SURVEY App. C statement); used to cross-check the C code."""
    X = np.asarray(X, dtype=np.float64)
    N = X.shape[0]
    h, ds, smin, x0 = 2.0 / N, det_w / D, -0.5 * det_w, -1.0 + 1.0 / N
    sj = smin + (np.arange(D) + 0.5) * ds
    coord = x0 + np.arange(N) * h
    out = np.zeros((len(c), D))
    for a, (ca, sa) in enumerate(zip(c, s)):
        xdom = abs(ca) > abs(sa)
        major, minor = (ca, sa) if xdom else (sa, ca)
        w = h / abs(major)
        # position on the interpolated axis for every (bin, step)
        pos = (sj[:, None] - coord[None, :] * minor) / major
        t = (pos - x0) / h
        i0 = np.floor(t).astype(np.int64)
        f = t - i0
        step = np.broadcast_to(np.arange(N)[None, :], t.shape)
        acc = np.zeros(D)
        for idx, wt in ((i0, 1.0 - f), (i0 + 1, f)):
            ok = (idx >= 0) & (idx < N)
            ii = np.clip(idx, 0, N - 1)
            vals = X[ii, step] if xdom else X[step, ii]
            acc += np.sum(np.where(ok, wt * w * vals, 0.0), axis=1)
        out[a] = acc
    return out


# ----------------------------------------------------------------------------------------------
# Phantoms.  Gen_Sino_Partitioned.py:5-64 (ConstIm), :67-122 (randIm), Shepp-Logan (BASELINE configs)
# ----------------------------------------------------------------------------------------------
_SHEPP_LOGAN_MODIFIED = [  # value, a, b, x0, y0, phi(deg)  (Toft's modified table)
    (1.0, .69, .92, 0.0, 0.0, 0.0),
    (-.8, .6624, .8740, 0.0, -.0184, 0.0),
    (-.2, .1100, .3100, .22, 0.0, -18.0),
    (-.2, .1600, .4100, -.22, 0.0, 18.0),
    (.1, .2100, .2500, 0.0, .35, 0.0),
    (.1, .0460, .0460, 0.0, .1, 0.0),
    (.1, .0460, .0460, 0.0, -.1, 0.0),
    (.1, .0460, .0230, -.08, -.605, 0.0),
    (.1, .0230, .0230, 0.0, -.606, 0.0),
    (.1, .0230, .0460, .06, -.605, 0.0),
]


def ellipse_image(N, ellipses):
    """Sample sum of ellipse indicators at pixel centres of [-1,1]^2; array [ix, iy]."""
    h = 2.0 / N
    g = -1.0 + (np.arange(N) + 0.5) * h
    X, Y = np.meshgrid(g, g, indexing="ij")
    img = np.zeros((N, N))
    for v, a, b, x0, y0, phi in ellipses:
        ph = math.radians(phi)
        xr = (X - x0) * math.cos(ph) + (Y - y0) * math.sin(ph)
        yr = -(X - x0) * math.sin(ph) + (Y - y0) * math.cos(ph)
        img[(xr / a) ** 2 + (yr / b) ** 2 <= 1.0] += v
    return img


def shepp_logan(N):
    return ellipse_image(N, _SHEPP_LOGAN_MODIFIED)


def ellipse_sinogram(theta, D, ellipses, det_w=2.0):
    """Closed-form Radon transform of the ellipse sum on the (theta, s_j) grid (independent anchor)."""
    theta = np.asarray(theta, dtype=np.float64)
    sj = -0.5 * det_w + (np.arange(D) + 0.5) * det_w / D
    out = np.zeros((len(theta), D))
    for v, a, b, x0, y0, phi in ellipses:
        ph = math.radians(phi)
        th = theta - ph
        a2 = (a * np.cos(th)) ** 2 + (b * np.sin(th)) ** 2
        s0 = x0 * np.cos(theta) + y0 * np.sin(theta)
        t = sj[None, :] - s0[:, None]
        arg = a2[:, None] - t ** 2
        out += v * 2.0 * a * b * np.sqrt(np.maximum(arg, 0.0)) / a2[:, None]
    return out


def ConstIm(N):
    """Gen_Sino_Partitioned.py:5-64 restated (fixed-layout rectangle + four discs, values 80-400)."""
    return _rect_disc_phantom(N, (N // 6, N // 5), (N // 3, N // 3), (3 * N // 5, 3 * N // 5),
                              (N // 10, N - N // 6), (N - N // 6, N // 10))


def randIm(N, rng=None):
    """Gen_Sino_Partitioned.py:67-122 restated; ``rng`` is a numpy RandomState-like with ``randint``
    (the reference draws from the unseeded global ``np.random``)."""
    rng = np.random if rng is None else rng
    ofs = rng.randint(N // 8, N // 4 + N // 8, size=2)
    c1 = rng.randint(N // 4, N // 2, size=2)
    c2 = rng.randint(N // 2, 3 * N // 4, size=2)
    c3 = rng.randint(0, N // 4, size=2) + np.array([0, N - N // 4])
    c4 = rng.randint(0, N // 4, size=2) + np.array([N - N // 4, 0])
    return _rect_disc_phantom(N, tuple(ofs), tuple(c1), tuple(c2), tuple(c3), tuple(c4))


def _disc(N, ctr, rad, val):
    """The reference's disc stamp: meshgrid over a clipped index box, written transposed via np.ix_(I2, I1)
    (Gen_Sino_Partitioned.py:21-28)."""
    tmp = np.zeros((N, N))
    I1 = np.arange(max(ctr[0] - rad, 0), min(ctr[0] + rad, N))
    I2 = np.arange(max(ctr[1] - rad, 0), min(ctr[1] + rad, N))
    Xg, Yg = np.meshgrid(I1, I2)
    cir = ((Xg - ctr[0]) ** 2 + (Yg - ctr[1]) ** 2) <= rad ** 2
    tmp[np.ix_(I2, I1)] = cir.astype(float) * val
    return tmp


def _rect_disc_phantom(N, rec, c1, c2, c3, c4):
    Im = np.zeros((N, N))
    Im[rec[0]:N, rec[1]:N] = 200
    t = _disc(N, c1, N // 2, 80)
    Im = np.where(t == 0, Im, t)
    Im = np.maximum(Im, _disc(N, c2, N // 8, 300))
    Im = np.maximum(Im, _disc(N, c3, N // 16, 400))
    Im = np.maximum(Im, _disc(N, c4, N // 16, 400))
    return Im


def psnr(x_hat, x_true, data_range=1.0):
    """test_final_integration.py:41-45."""
    mse = np.mean((np.asarray(x_hat, dtype=np.float64) - np.asarray(x_true, dtype=np.float64)) ** 2)
    if mse == 0:
        return float("inf")
    return 20.0 * np.log10(data_range) - 10.0 * np.log10(mse)


# ----------------------------------------------------------------------------------------------
# TV helpers (a9-a12).  block_4_tv_helpers.py
# ----------------------------------------------------------------------------------------------
def grad_forward(x_vec, N):
    """block_4_tv_helpers.py:17-23."""
    X = np.asarray(x_vec, dtype=np.float64).reshape(N, N)
    gx = np.zeros_like(X)
    gy = np.zeros_like(X)
    gx[:-1, :] = X[1:, :] - X[:-1, :]
    gy[:, :-1] = X[:, 1:] - X[:, :-1]
    return gx, gy


def grad_T(px, py, N):
    """Exact K^T (adjoint of grad_forward)."""
    out = np.zeros((N, N))
    out[:-1, :] -= px[:-1, :]
    out[1:, :] += px[:-1, :]
    out[:, :-1] -= py[:, :-1]
    out[:, 1:] += py[:, :-1]
    return out.reshape(-1)


def div_reference(px, py, N):
    """block_4_tv_helpers.py:25-35 as shipped (border rows/cols sign-flipped vs K^T, App. B-4)."""
    px = _f64(px)
    py = _f64(py)
    out = np.empty(N * N)
    lib().orc_div_reference(_p(px.reshape(-1)), _p(py.reshape(-1)), N, _p(out))
    return out


def kt_subgrad(x_vec, N, eps=1e-12, exact_adjoint=False):
    """block_4_tv_helpers.py:37-46."""
    gx, gy = grad_forward(x_vec, N)
    mag = np.sqrt(gx ** 2 + gy ** 2)
    mask = mag > eps
    px = np.zeros_like(gx)
    py = np.zeros_like(gy)
    px[mask] = gx[mask] / mag[mask]
    py[mask] = gy[mask] / mag[mask]
    return grad_T(px, py, N) if exact_adjoint else div_reference(px, py, N)


def tv_canonical(x_vec, N):
    gx, gy = grad_forward(x_vec, N)
    return float(np.sum(np.sqrt(gx ** 2 + gy ** 2)))


def tv_reference_pairing(x_vec, N):
    """block_4_tv_helpers.py:5-14 evaluated numerically with CVXPY's Fortran-order reshapes
    (mis-paired differences, App. B-3).  Evaluator only."""
    X = np.asarray(x_vec, dtype=np.float64).reshape((N, N), order="F")
    Dx = (X[1:, :] - X[:-1, :]).reshape(-1, order="F")
    Dy = (X[:, 1:] - X[:, :-1]).reshape(-1, order="F")
    return float(np.sum(np.sqrt(Dx ** 2 + Dy ** 2)))


# ----------------------------------------------------------------------------------------------
# Precisions and graphs (a7, a8)
# ----------------------------------------------------------------------------------------------
def make_precisions(Wi_raw, q_mode="arithmetic"):
    """block_3_graph_and_precisions.py:11-43 with W_i supplied as column norms^2 (dtype preserved: the reference's
    W, Q are float32 when A is)."""
    eps = 1e-12
    Wi_list = [np.maximum(np.asarray(w), eps) for w in Wi_raw]
    if q_mode == "harmonic":
        def Q(i, j):
            return np.maximum(Wi_list[i] * Wi_list[j] / (Wi_list[i] + Wi_list[j]), eps)
    elif q_mode == "arithmetic":
        def Q(i, j):
            return np.maximum(0.5 * (Wi_list[i] + Wi_list[j]), eps)
    else:
        raise ValueError("q_mode must be 'harmonic' or 'arithmetic'")
    return Wi_list, Q


def make_graph(kind, V, seed=0, p=0.1, degree=4):
    """Node-level graphs of the BASELINE configs; networkx generators, same seeds."""
    import networkx as nx
    if kind == "ring":
        return nx.cycle_graph(V)
    if kind == "regular":
        return nx.random_regular_graph(degree, V, seed=seed)
    if kind == "er":
        s = seed
        while True:
            G = nx.erdos_renyi_graph(V, p, seed=s)
            if nx.is_connected(G):
                return G
            s += 1
    if kind == "complete":
        return nx.complete_graph(V)
    if kind == "path":
        return nx.path_graph(V)
    raise ValueError(kind)


def graph_csr(G):
    """Edge list in G.edges() order keyed (min,max) (block_6_admm_loop_ver2.py:39-40) and neighbour lists in
    G.neighbors(i) order (:87) as CSR: nbr_ptr, nbr_idx, nbr_edge, nbr_end (0 if i is the min end)."""
    V = G.number_of_nodes()
    edges = [(min(i, j), max(i, j)) for i, j in G.edges()]
    eid = {e: k for k, e in enumerate(edges)}
    ptr, idx, ed, end = [0], [], [], []
    for i in range(V):
        for j in G.neighbors(i):
            key = (min(i, j), max(i, j))
            idx.append(j)
            ed.append(eid[key])
            end.append(0 if i == key[0] else 1)
        ptr.append(len(idx))
    return (np.array(edges, dtype=np.int32).reshape(-1, 2), np.array(ptr, dtype=np.int32),
            np.array(idx, dtype=np.int32), np.array(ed, dtype=np.int32), np.array(end, dtype=np.int32))


# ----------------------------------------------------------------------------------------------
# Node x-update (a13/a14 replacement) and the outer loop (a15-a19)
# ----------------------------------------------------------------------------------------------
def x_update(op, prec, rhs0, rhoD, mu, lam, S, C, x, d, w):
    """In-place S sweeps x C CG iterations (C implementation).  Returns (Ax, r_final, tvrhs)."""
    n = op.N * op.N
    Ax = np.empty(op.nang * op.D)
    r = np.empty(n)
    tvrhs = np.empty(n)
    vec = None if np.isscalar(rhoD) else _f64(rhoD)
    lib().orc_x_update(op.N, op.D, op.det_w, _p(op.c), _p(op.s), op.nang, float(prec), _p(_f64(rhs0)),
                       _p(vec), float(rhoD) if vec is None else 0.0, float(mu), float(lam), int(S), int(C),
                       _p(x), _p(d), _p(w), _p(Ax), _p(r), _p(tvrhs))
    return Ax, r, tvrhs


def np_x_update(op, prec, rhs0, rhoD, mu, lam, S, C, x, d, w):
    """Pure-NumPy twin of orc_x_update (same arithmetic, readable); in-place on x, d, w."""
    N, n = op.N, op.N * op.N

    def H(v):
        Av = op.forward(v)
        V2 = v.reshape(N, N)
        gx, gy = grad_forward(v, N)
        return op.adjoint(prec * Av) + rhoD * v + mu * grad_T(gx, gy, N), Av

    for _ in range(S):
        tvrhs = mu * grad_T((d[:n] - w[:n]).reshape(N, N), (d[n:] - w[n:]).reshape(N, N), N)
        rhs = rhs0 + tvrhs
        Hx, Ax = H(x)
        r = rhs - Hx
        p = r.copy()
        rr = float(r @ r)
        for _ in range(C):
            Hp, Ap = H(p)
            pHp = float(p @ Hp)
            alpha = rr / pHp if pHp > 0 else 0.0
            x += alpha * p
            r -= alpha * Hp
            Ax += alpha * Ap
            rr_new = float(r @ r)
            beta = rr_new / rr if rr > 0 else 0.0
            p = r + beta * p
            rr = rr_new
        gx, gy = grad_forward(x, N)
        g1 = gx.reshape(-1) + w[:n]
        g2 = gy.reshape(-1) + w[n:]
        nrm = np.sqrt(g1 ** 2 + g2 ** 2)
        kappa = lam / mu
        sc = np.where(nrm > kappa, 1.0 - kappa / np.maximum(nrm, 1e-300), 0.0)
        d[:n] = sc * g1
        d[n:] = sc * g2
        w[:n] = g1 - d[:n]
        w[n:] = g2 - d[n:]
    return Ax, r, tvrhs


def decentralized_admm(ops, sinograms, G, Wi_list, Qij_diag_fn, N, lam_tv=0.01, rho=1.0, max_iters=10,
                       eps_pri=1e-1, eps_dual=1e-1, phantom_true=None, node_prec=None, tv_mu=None,
                       tv_sweeps=1, cg_iters=2, weighted_z=False, uniform_q=None, stop=True,
                       x_update_fn=None, node_subset=None, acceptance=True, max_tighten=2, on_iteration=None):
    """Array restatement of block_6_admm_loop_ver2.py:15-326 with the SCS solve (:97-176) replaced by the
    TV-split + CG x-update.  Same initialisation (:36-46), Jacobi node sweep (:81-97,187), metrics (:189-206),
    midpoint z (:210-223) [W-weighted PDF eq. (2) if ``weighted_z``], duals (:225-230), residuals (:232-264),
    stop test (:286-289) and history keys (:310-326).

    ``acceptance`` (default on, as in the reference): restate the reference's accept / tighten-and-retry rule (:100-108, :155-176): after a solve the
    stationarity norm |g_x,i| (:137-149) is compared with eps_target = 2/(k+1)^1.005; a node that misses it is solved
    again (warm-started, like the re-solve of the same ``cp.Problem`` with ``warm_start=True``) with eps/5, at most
    ``max_tighten`` = 2 more times, and ``eps_used_history`` records the eps of the accepted try (:161,170).  One
    "solve" here is ``tv_sweeps`` x ``cg_iters`` of the TV-split + CG x-update; eps itself only labels the try.
    ``history['tighten_history']`` holds the number of extra solves per node.

    ``uniform_q``: scalar q used instead of calling ``Qij_diag_fn`` (then D_i = deg_i * q).
    ``node_subset``: if given, only those nodes are x-updated (bounded CPU-baseline sample); others keep x.
    """
    V = len(ops)
    n = N * N
    mu = float(tv_mu if tv_mu is not None else rho)
    edges, ptr, nidx, nedge, nend = graph_csr(G)
    E = len(edges)
    node_prec = [1.0] * V if node_prec is None else list(node_prec)
    b = [np.asarray(s, dtype=np.float64).reshape(-1) for s in sinograms]
    x = [np.zeros(n) for _ in range(V)]
    d = [np.zeros(2 * n) for _ in range(V)]
    w = [np.zeros(2 * n) for _ in range(V)]
    z = [np.zeros(n) for _ in range(E)]
    y = [[np.zeros(n), np.zeros(n)] for _ in range(E)]
    Atb = [ops[i].adjoint(node_prec[i] * b[i]) for i in range(V)]
    if uniform_q is None:
        Qd = {}
        for i in range(V):
            for k in range(ptr[i], ptr[i + 1]):
                Qd[(i, int(nidx[k]))] = np.asarray(Qij_diag_fn(i, int(nidx[k])), dtype=np.float64)
    xupd = x_update_fn or x_update
    hist = {k: [] for k in ("primal", "dual", "pri_per_node", "dual_per_node", "obj_per_node", "obj_total",
                            "mse_sino_per_node", "mse_sino_total", "img_mse_per_node", "img_mse_total",
                            "g_norm_history", "eps_used_history", "eps_target_history", "tighten_history")}
    phantom_vec = None if phantom_true is None else np.asarray(phantom_true, dtype=np.float64).reshape(-1)
    nodes = range(V) if node_subset is None else node_subset
    L = lib()
    for k in range(max_iters):
        new_x = [xi for xi in x]
        mse_i = np.zeros(V)
        tvv = np.zeros(V)
        g_norm = np.zeros(V)
        eps_used = np.zeros(V)
        tighten = np.zeros(V, dtype=np.int64)
        eps_target = 2.0 / ((k + 1) ** 1.005)  # block_6_admm_loop_ver2.py:101-103
        for i in nodes:
            cons = np.zeros(n)
            if uniform_q is None:
                Dv = np.zeros(n)
            for kk in range(ptr[i], ptr[i + 1]):
                e, end = int(nedge[kk]), int(nend[kk])
                if uniform_q is None:
                    q = Qd[(i, int(nidx[kk]))]
                    Dv += q
                    L.orc_accum_cons(n, rho, _p(q), 0.0, _p(z[e]), _p(y[e][end]), _p(cons))
                else:
                    L.orc_accum_cons(n, rho, None, float(uniform_q), _p(z[e]), _p(y[e][end]), _p(cons))
            deg = ptr[i + 1] - ptr[i]
            rhoD = rho * Dv if uniform_q is None else rho * deg * float(uniform_q)
            xi = x[i].copy()
            eps_try = min(1e-2, eps_target)  # :106-108
            tries = 0
            while True:
                Ax, r, tvrhs = xupd(ops[i], node_prec[i], Atb[i] + cons, rhoD, mu, lam_tv, tv_sweeps, cg_iters,
                                    xi, d[i], w[i])
                # a14 stationarity (:137-149) through the identity
                #   A^T P(Ax-b) + rho(Dx - sum q v) = tvrhs - r - mu K^T K x
                gx, gy = grad_forward(xi, N)
                g_vec = tvrhs - r - mu * grad_T(gx, gy, N) + lam_tv * kt_subgrad(xi, N)
                g_norm[i] = float(np.linalg.norm(g_vec))
                if not acceptance or g_norm[i] <= eps_target or tries >= max_tighten:   # :155-172
                    break
                tries += 1                   # :175-176
                eps_try /= 5.0
            eps_used[i], tighten[i] = eps_try, tries
            new_x[i] = xi
            res = Ax - b[i]
            mse_i[i] = float(res @ res)  # :190-194
            tvv[i] = tv_canonical(xi, N)
        x = new_x
        hist["g_norm_history"].append(g_norm)
        hist["eps_used_history"].append(eps_used)
        hist["eps_target_history"].append(np.full(V, eps_target))
        hist["tighten_history"].append(tighten)
        hist["mse_sino_per_node"].append(mse_i.copy())
        hist["mse_sino_total"].append(float(np.sum(mse_i)))
        if phantom_vec is not None:
            img = np.array([float((xi - phantom_vec) @ (xi - phantom_vec)) for xi in x])
        else:
            img = np.zeros(V)
        hist["img_mse_per_node"].append(img)
        hist["img_mse_total"].append(float(np.sum(img)))
        # edges
        r2 = s2 = 0.0
        pri_node = np.zeros(V)
        dual_node = np.zeros(V)
        pen = np.zeros(V)
        sums = np.zeros(5)
        for e, (i, j) in enumerate(edges):
            i, j = int(i), int(j)
            if uniform_q is None:
                qij, qji, qs = _p(Qd[(i, j)]), _p(Qd[(j, i)]), 0.0
            else:
                qij = qji = None
                qs = float(uniform_q)
            Wi = _p(_f64(Wi_list[i])) if weighted_z else None
            Wj = _p(_f64(Wi_list[j])) if weighted_z else None
            L.orc_edge_update(n, _p(x[i]), _p(x[j]), _p(y[e][0]), _p(y[e][1]), _p(z[e]), Wi, Wj, qij, qji, qs,
                              _p(sums))
            r2 += sums[0] + sums[1]
            pri_node[i] += sums[0]
            pri_node[j] += sums[1]
            s2 += rho * rho * sums[2]
            dual_node[i] += rho * rho * sums[2]
            dual_node[j] += rho * rho * sums[2]
            pen[i] += sums[3]
            pen[j] += sums[4]
        obj_i = 0.5 * np.array(node_prec) * mse_i + lam_tv * tvv + 0.5 * rho * pen
        pri_norm, dual_norm = math.sqrt(r2), math.sqrt(s2)
        hist["primal"].append(pri_norm)
        hist["dual"].append(dual_norm)
        hist["obj_per_node"].append(obj_i)
        hist["obj_total"].append(float(np.sum(obj_i)))
        hist["pri_per_node"].append(np.sqrt(pri_node))
        hist["dual_per_node"].append(np.sqrt(dual_node))
        if on_iteration is not None:
            on_iteration(k, x)
        if stop and pri_norm < eps_pri and dual_norm < eps_dual:
            break
    return x, hist


# ---- PDHG consensus variant (ADMM_Tomo_Only.py:89-148) -- NumPy fp64 restatement -----------------
# Test infrastructure like the rest of this file.  PARITY UNPINNED at odl.solvers.pdhg / odl.Gradient /
# odl.power_method_opnorm (ODL is not installable here and the reference holds no vectors for this script): the
# conventions below are ODL's as recalled and are stated, not verified -- X = uniform_discr([-1,1]^2, (N,N)), cell
# h = 2/N; Gradient = forward differences / h with zero padding beyond the last index; adjoint of the ray transform
# = (w_Y / w_X) A^T; prox formulas of L2NormSquared.translated and GroupL1Norm (weights cancel in both).
def odl_grad(x, N):
    """odl.Gradient(space)(x): forward differences / h, zero padding (ADMM_Tomo_Only.py:63)."""
    X = x.reshape(N, N)
    ih = N / 2.0
    g1 = np.vstack([X[1:], np.zeros((1, N))]) - X
    g2 = np.hstack([X[:, 1:], np.zeros((N, 1))]) - X
    return g1 * ih, g2 * ih


def odl_grad_adjoint(p1, p2, N):
    """Gradient.adjoint (= -Divergence with the matching zero padding): the plain transpose of odl_grad."""
    ih = N / 2.0
    a = np.vstack([np.zeros((1, N)), p1[:-1]]) - p1
    b = np.hstack([np.zeros((N, 1)), p2[:, :-1]]) - p2
    return ((a + b) * ih).reshape(-1)


def pdhg_opnorm(op, adj_scale, N, iters=30):
    """odl.power_method_opnorm(BroadcastOperator(A, Gradient)) (:128,:145) from a deterministic start (the reference's
    is random): `iters` steps of x <- L* L x / |.|, returns sqrt(|L* L x| / |x|) of the last step."""
    n = N * N
    x = 1.0 + 0.5 * np.cos(0.37 * np.arange(n))
    est = 0.0
    for _ in range(iters):
        x = x / np.linalg.norm(x)
        g1, g2 = odl_grad(x, N)
        y = adj_scale * op.adjoint(op.forward(x)) + odl_grad_adjoint(g1, g2, N)
        est = math.sqrt(np.linalg.norm(y))
        x = y
    return est


def pdhg_steps(op, adj_scale, b, x, y1, y2, niter, tau, sigma, gamma, pull, lam_d, lam_t, N, theta=1.0):
    """`niter` steps of odl.solvers.pdhg (:132-133) for gamma|x - pull|^2 + lam_d|A x - b|^2 + lam_t|G x|_{2,1};
    x_relax starts at x (pdhg's default).  `op` may be a list (aggregate problem: rows stacked, `b`, `y1` lists)."""
    ops = op if isinstance(op, (list, tuple)) else [op]
    bs = b if isinstance(b, (list, tuple)) else [b]
    y1s = y1 if isinstance(y1, (list, tuple)) else [y1]
    sc = adj_scale if isinstance(adj_scale, (list, tuple)) else [adj_scale] * len(ops)
    xbar = x.copy()
    for _ in range(niter):
        for k, o in enumerate(ops):
            y1s[k] = (y1s[k] + sigma * (o.forward(xbar) - bs[k])) / (1.0 + sigma / (2.0 * lam_d))
        g1, g2 = odl_grad(xbar, N)
        t1, t2 = y2[0] + sigma * g1, y2[1] + sigma * g2
        s = np.maximum(1.0, np.sqrt(t1 * t1 + t2 * t2) / lam_t)
        y2 = (t1 / s, t2 / s)
        back = sum(sc[k] * o.adjoint(y1s[k]) for k, o in enumerate(ops))
        v = x - tau * (back + odl_grad_adjoint(y2[0], y2[1], N))
        w = 2.0 * tau * gamma
        xn = (v + (w * pull if pull is not None else 0.0)) / (1.0 + w)
        xbar = xn + theta * (xn - x)
        x = xn
    return x, (y1s if isinstance(y1, (list, tuple)) else y1s[0]), y2


def pdhg_consensus(ops, sinograms, phantom, N, niter=100, lambda_penalty=0.005, alpha_tv=0.0, lambda_agg=0.005,
                   gamma=2.0, node_niter=5, agg_niter=15, opnorm_iters=30, angle_cells=None, agg_angle_cell=None):
    """ADMM_Tomo_Only.py:89-148: per outer iteration the error-weighted convex combination x_a of the node iterates
    (:100-118, ground-truth dependent), 5 cold-dual PDHG steps per node pulled towards x_a (:121-133), 15 warm-dual PDHG
    steps of the aggregate problem (:142-148), and the four metric lists (:134-139, :152-159)."""
    V, n = len(ops), N * N
    ph = np.asarray(phantom, dtype=np.float64).reshape(-1)
    b = [np.asarray(s, dtype=np.float64).reshape(-1) for s in sinograms]
    hx2 = (2.0 / N) ** 2
    cells = [math.pi / o.nang for o in ops] if angle_cells is None else list(angle_cells)
    adj = [c * (o.det_w / o.D) / hx2 for c, o in zip(cells, ops)]
    agg_cell = (math.pi / sum(o.nang for o in ops)) if agg_angle_cell is None else agg_angle_cell
    adj_agg = [agg_cell * (o.det_w / o.D) / hx2 for o in ops]
    cn = [np.sqrt(o.colnorm2()) for o in ops]                            # :55 np.linalg.norm(A_i, axis=0)
    norms = [pdhg_opnorm(o, a, N, opnorm_iters) for o, a in zip(ops, adj)]

    class _Agg:   # the stacked operator with the aggregate range weighting
        def forward(self, x): return np.concatenate([o.forward(x) for o in ops])
        def adjoint(self, q):
            out, pos = np.zeros(n), 0
            for o in ops:
                out += o.adjoint(q[pos:pos + o.shape[0]])
                pos += o.shape[0]
            return out
    norm_agg = pdhg_opnorm(_Agg(), adj_agg[0], N, opnorm_iters)
    x = [np.zeros(n) for _ in range(V)]
    x_agg = np.zeros(n)
    y1_agg = [np.zeros_like(bi) for bi in b]
    y2_agg = (np.zeros((N, N)), np.zeros((N, N)))
    out = {"mse_lists": [[] for _ in range(V)], "mse_sino_lists": [[] for _ in range(V)], "mse_agg_list": [],
           "mse_agg_sino_list": [], "op_norms": norms, "op_norm_agg": norm_agg}
    for k in range(niter):
        lam = lambda_penalty * math.exp(alpha_tv * k)
        eta = np.stack([cn[i] / (np.abs(x[i] - ph) + 1e-8) for i in range(V)])
        xa = np.sum((eta / (eta.sum(axis=0) + 1e-8)) * np.stack(x), axis=0)
        for i in range(V):
            t = 1.0 / norms[i]
            x[i], _, _ = pdhg_steps(ops[i], adj[i], b[i], x[i], np.zeros_like(b[i]),
                                    (np.zeros((N, N)), np.zeros((N, N))), node_niter, t, t, gamma, xa, lam, lam, N)
            out["mse_lists"][i].append(float(np.mean((x[i] - ph) ** 2)))
            out["mse_sino_lists"][i].append(float(np.linalg.norm(ops[i].forward(x[i]) - b[i])))
        t = 1.0 / norm_agg
        x_agg, y1_agg, y2_agg = pdhg_steps(list(ops), adj_agg, b, x_agg, y1_agg, y2_agg, agg_niter, t, t, 0.0, None,
                                           1.0, lambda_agg, N)
        out["mse_agg_list"].append(float(np.mean((x_agg - ph) ** 2)))
        out["mse_agg_sino_list"].append(float(math.sqrt(sum(np.sum((o.forward(x_agg) - bi) ** 2)
                                                            for o, bi in zip(ops, b)))))
    out["x_vars"], out["x_agg"] = x, x_agg
    return out
