"""Drop-in for the reference's Gen_Sino_Partitioned.py: phantoms (`ConstIm` :5-64, `randIm` :67-122) and
`generate_sinogram` (:124-147) with the ODL RayTransform replaced by the CUDA operator.  The dense matrix the
reference builds by probing every unit vector (:138-145) is not materialised: the returned `A` is the matrix-free
operator itself (`A @ x`, `A.T @ r`, `A.shape`)."""
from __future__ import annotations

import numpy as np

from admm_b200 import RayTransformCUDA


def _disc(N, ctr, rad, val):
    # the reference's disc stamp: meshgrid over the clipped index box, written through np.ix_(I2, I1) (:21-28)
    tmp = np.zeros((N, N))
    I1 = np.arange(max(ctr[0] - rad, 0), min(ctr[0] + rad, N))
    I2 = np.arange(max(ctr[1] - rad, 0), min(ctr[1] + rad, N))
    Xg, Yg = np.meshgrid(I1, I2)
    tmp[np.ix_(I2, I1)] = (((Xg - ctr[0]) ** 2 + (Yg - ctr[1]) ** 2) <= rad ** 2).astype(float) * val
    return tmp


def _phantom(N, rec, c1, c2, c3, c4):
    Im = np.zeros((N, N))
    Im[rec[0]:N, rec[1]:N] = 200
    t = _disc(N, c1, N // 2, 80)
    Im = np.where(t == 0, Im, t)
    Im = np.maximum(Im, _disc(N, c2, N // 8, 300))
    Im = np.maximum(Im, _disc(N, c3, N // 16, 400))
    Im = np.maximum(Im, _disc(N, c4, N // 16, 400))
    return Im


def ConstIm(N):
    """Gen_Sino_Partitioned.py:5-64."""
    return _phantom(N, (N // 6, N // 5), (N // 3, N // 3), (3 * N // 5, 3 * N // 5), (N // 10, N - N // 6),
                    (N - N // 6, N // 10))


def randIm(N, seed=None):
    """Gen_Sino_Partitioned.py:67-122.  Draws from the global np.random like the reference; `seed` (which
    block_2_load_odl_data.py:137 passes but the reference signature lacks) selects a private
    RandomState instead."""
    rng = np.random if seed is None else np.random.RandomState(seed)
    ofs = rng.randint(N // 8, N // 4 + N // 8, size=2)
    c1 = rng.randint(N // 4, N // 2, size=2)
    c2 = rng.randint(N // 2, 3 * N // 4, size=2)
    c3 = rng.randint(0, N // 4, size=2) + np.array([0, N - N // 4])
    c4 = rng.randint(0, N // 4, size=2) + np.array([N - N // 4, 0])
    return _phantom(N, tuple(ofs), tuple(c1), tuple(c2), tuple(c3), tuple(c4))


def generate_sinogram(image, angles, plot_sparsity=True):
    """Gen_Sino_Partitioned.py:124-147: sinogram of `image` for len(angles) views over [0, pi] (midpoints of
    uniform_partition(0, pi, len(angles)), :129) on an N-bin detector over [-1, 1] (:130), plus the reference's
    5e-13 white noise (:135).  Returns (noisy_sinogram, ray_trafo, geometry, space, A)."""
    image = np.asarray(image)
    N = image.shape[0]
    m = len(angles)
    theta = (np.arange(m, dtype=np.float64) + 0.5) * np.pi / m
    ray_trafo = RayTransformCUDA(N, theta, D=N, det_w=2.0, impl="skimage")
    phantom = ray_trafo.domain.element(image)
    sinogram = ray_trafo(phantom)
    noisy = sinogram + ray_trafo.range.element(np.random.normal(0.0, 1.0, size=ray_trafo.range.shape)) * 0.0000000000005
    geometry = {"angles": theta, "det_min": -1.0, "det_max": 1.0, "det_pixels": N}
    return noisy, ray_trafo, geometry, ray_trafo.domain, ray_trafo
