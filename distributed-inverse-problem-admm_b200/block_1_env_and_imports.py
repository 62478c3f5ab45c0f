"""Drop-in for block_1_env_and_imports.py:10-18 (small helpers; no cvxpy import)."""
import math  # noqa: F401
import os  # noqa: F401
import pickle  # noqa: F401

import numpy as np
import networkx as nx  # noqa: F401

from Gen_Sino_Partitioned import ConstIm, randIm  # noqa: F401  (block_2_load_odl_data.py:13 imports them from here)


def vec(img_2d):
    return img_2d.reshape(-1)


def unvec(x_vec, N):
    return x_vec.reshape(N, N)


def diag_from_column_norms(A_dense):
    """eta_j = ||A(:, j)||_2^2 -- operator objects answer through the K2b kernel (block_1:16-18)."""
    if hasattr(A_dense, "colnorm2"):
        return A_dense.colnorm2()
    return np.sum(A_dense * A_dense, axis=0)
