"""Drop-in for the reference's block_1_env_and_imports.py: the three array helpers other blocks import from it
(:10-18), plus the phantom generators that block_2_load_odl_data.py:13 expects to find here.  Imports neither cvxpy
nor anything else the hot path does not need."""
import numpy as np

from Gen_Sino_Partitioned import ConstIm, randIm  # noqa: F401


def vec(img_2d):
    """(N, N) image -> length N*N vector, C order (ix*N + iy)."""
    return np.reshape(img_2d, -1)


def unvec(x_vec, N):
    """Inverse of `vec`."""
    return np.reshape(x_vec, (N, N))


def diag_from_column_norms(A_dense):
    """eta_p = ||A(:, p)||_2^2.  Matrix-free operators answer through the K2b kernel (admm_colnorm2); a dense array is
    reduced with einsum."""
    if hasattr(A_dense, "colnorm2"):
        return A_dense.colnorm2()
    A = np.asarray(A_dense)
    return np.einsum("rp,rp->p", A, A)
