"""Drop-in for block_4_tv_helpers_with_plot.py: same TV helpers + edge maps (:23-62).  PNG output needs matplotlib,
which the hot path does not depend on; `save_edge_map` writes .npy and a PNG only if matplotlib is importable."""
import os

import numpy as np

from block_4_tv_helpers import (_div_backward_2d_to_vec, _grad_forward_2d_from_vec, edge_map_from_vector,  # noqa: F401
                                isotropic_tv_on_vector, kt_subgrad_isotropic_tv_from_x)


def save_edge_map(x_vec, N, out_path, show=False, cmap="gray", dpi=300):
    """block_4_tv_helpers_with_plot.py:48-62."""
    em = edge_map_from_vector(x_vec, N, normalize=True)
    os.makedirs(os.path.dirname(os.path.abspath(out_path)) or ".", exist_ok=True)
    np.save(os.path.splitext(out_path)[0] + ".npy", em)
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        plt.figure(figsize=(5, 5))
        plt.imshow(em, cmap=cmap)
        plt.title("Discrete gradient magnitude")
        plt.axis("off")
        plt.tight_layout()
        plt.savefig(out_path, dpi=dpi)
        plt.close()
    except ImportError:
        pass
    return em
