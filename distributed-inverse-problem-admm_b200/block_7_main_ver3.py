"""Drop-in for the reference's live driver block_7_main_ver3.py: `save_recons` (:16-27), `run_one_strategy` (:30-329)
and `main` (:332-371) with the same signatures, call sequence (block_2 -> block_3 -> block_6) and output files
(`run_parameters.txt`, the per-node reconstructions and every history as `.npy`).  The reference's ~12 matplotlib
figures are written only when matplotlib is importable (it is not a dependency of the hot path).

The reference's own file also runs unchanged against these modules (put this directory first on sys.path); this
copy exists so that the end-to-end flow can be run where matplotlib is absent."""
import os
from datetime import datetime

import numpy as np

from block_2_load_odl_data import load_odl_data
from block_3_graph_and_precisions import build_pixel_connected_Q_provider
from block_6_admm_loop_ver2 import decentralized_admm

try:  # optional
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt
except ImportError:  # pragma: no cover
    plt = None


def _png(path, draw):
    if plt is None:
        return
    plt.figure(figsize=(7, 4))
    draw()
    plt.tight_layout()
    plt.savefig(path, dpi=220)
    plt.close()


def save_recons(x_list, N, out_dir, tag):
    os.makedirs(out_dir, exist_ok=True)
    for i, x_i in enumerate(x_list):
        img = np.asarray(x_i).reshape(N, N)
        np.save(os.path.join(out_dir, f"{tag}_node_{i}.npy"), img)
        _png(os.path.join(out_dir, f"{tag}_node_{i}.png"), lambda: (plt.imshow(img, cmap="gray"), plt.axis("off")))


def run_one_strategy(strategy, k, data, N, lam_tv, rho,
                     max_iters, max_inner_iters, eps_pri, eps_dual,
                     base_dir, out_root, show_plots=False, verbose=True, snapshot_div=5, phantom_true=None):
    tag = f"{strategy}_k{k}" if strategy == "knn" else strategy
    out_dir = os.path.join(out_root, tag)
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "run_parameters.txt"), "w") as f:
        f.write("===== Global Parameters =====\n")
        for name, val in (("Strategy", strategy), ("Image side N", N), ("Lambda_TV", lam_tv), ("Rho", rho),
                          ("Max ADMM iterations", max_iters), ("Max Inner iterations", max_inner_iters),
                          ("Date-Time", datetime.now().strftime("%Y-%m-%d %H:%M:%S"))):
            f.write(f"{name}: {val}\n")
        f.write("\n===== Solver =====\nTV-split + CG node updates on the GPU (libadmm_b200)\n")
        f.write(f"\n===== Data Info =====\nOutput directory: {out_dir}\n")

    G_union, Wi_list, Qij_diag_fn_masked, keep = build_pixel_connected_Q_provider(
        base_dir=base_dir, strategy=strategy, k=k, seed=123, q_mode="arithmetic", verbose=verbose, plot_union=True,
        show_plots=show_plots, output_dir=os.path.join(out_dir, "union_figs"))

    snap_dir = os.path.join(out_dir, "snapshots")
    os.makedirs(snap_dir, exist_ok=True)
    snap_every = max(1, max_iters // snapshot_div)
    x_list, hist = decentralized_admm(
        A_dense_list=data["A_dense_list"], sinograms=data["sinograms"], G=G_union, Wi_list=Wi_list,
        Qij_diag_fn=Qij_diag_fn_masked, N=N, lam_tv=lam_tv, rho=rho, max_iters=max_iters,
        max_inner_iters=max_inner_iters, eps_pri=eps_pri, eps_dual=eps_dual, verbose=verbose, snapshot_dir=snap_dir,
        snapshot_every=snap_every, snapshot_div=snapshot_div, phantom_true=phantom_true)

    save_recons(x_list, N, out_dir, tag)
    dumps = {"obj_per_node": np.vstack(hist["obj_per_node"]), "obj_total": np.array(hist["obj_total"]),
             "pri_per_node": np.vstack(hist["pri_per_node"]), "dual_per_node": np.vstack(hist["dual_per_node"]),
             "primal_hist": np.array(hist["primal"]), "dual_hist": np.array(hist["dual"]),
             "sino_mse_per_node": np.vstack(hist["mse_sino_per_node"]), "sino_mse_total": np.array(hist["mse_sino_total"]),
             "g_norm_per_node": np.vstack(hist["g_norm_history"])}
    if phantom_true is not None:
        dumps["img_mse_per_node"] = np.vstack(hist["img_mse_per_node"])
        dumps["img_mse_total"] = np.array(hist["img_mse_total"])
    for name, arr in dumps.items():
        np.save(os.path.join(out_dir, f"{tag}_{name}.npy"), arr)
        _png(os.path.join(out_dir, f"{tag}_{name}.png"), lambda a=arr: (plt.semilogy(np.maximum(a, 1e-300)), plt.grid(True)))
    print(f"[Done] Saved outputs to {out_dir}")
    return x_list, hist


def main(N=64, num_nodes=5, lam_tv=0.02, rho=2.0, max_iters=200, max_inner_iters=100, eps_pri=1e-3, eps_dual=1e-3,
         noise_level=0.005, base_dir="saved_operators_Incmp_Span", snapshot_div=2, out_root=None):
    """block_7_main_ver3.py:332-371 (same settings)."""
    data = load_odl_data(base_dir=base_dir, N=N, num_nodes=num_nodes, noise_level=noise_level)
    phantom_true = data.get("phantom", None)
    out_root = out_root or f"Recon_Out_ADMM_{datetime.now().strftime('%Y%m%d_%H%M%S')}"
    return run_one_strategy("knn", k=2, data=data, N=N, lam_tv=lam_tv, rho=rho, max_iters=max_iters,
                            max_inner_iters=max_inner_iters, eps_pri=eps_pri, eps_dual=eps_dual, base_dir=base_dir,
                            out_root=out_root, snapshot_div=snapshot_div, phantom_true=phantom_true)


if __name__ == "__main__":
    main()
