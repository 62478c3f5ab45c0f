"""Drop-in for the reference's block_3_graph_and_precisions.py: precisions W_i / Q_ij (:11-43) and the
Q-provider entry point (:262-319), fed by matrix-free operators (W_i from the K2b kernel instead of
np.sum(A_i*A_i, axis=0) on a dense matrix).

Node-level graphs (`ring`, `regular`, `er`, `complete`, `path`; the BASELINE configs) are built on the host with
networkx.  The reference's per-pixel strategies (`knn`, `mst`, `chain`; :62-187, a V*V*n boolean mask) are kept as a
host-side option for small problems only (17 GB at V=64, N=2048)."""
from __future__ import annotations

import numpy as np
import networkx as nx

from admm_b200 import make_graph

_EPS = 1e-12


def _colnorm2(A_i):
    if hasattr(A_i, "colnorm2"):
        return np.asarray(A_i.colnorm2(), dtype=np.float64)
    A_i = np.asarray(A_i)
    if A_i.ndim == 1:            # already a W vector
        return A_i.astype(np.float64)
    return np.sum(A_i * A_i, axis=0)


def make_precisions(A_dense_list, q_mode="arithmetic"):
    """block_3_graph_and_precisions.py:11-43.
    Wi[p] = ||A_i[:, p]||_2^2 floored at 1e-12; Qij harmonic Wi*Wj/(Wi+Wj) or arithmetic 0.5*(Wi+Wj), floored."""
    Wi_list = [np.maximum(_colnorm2(A_i), _EPS) for A_i in A_dense_list]
    if q_mode == "harmonic":
        def Qij_diag(i, j):
            return np.maximum((Wi_list[i] * Wi_list[j]) / (Wi_list[i] + Wi_list[j]), _EPS)
    elif q_mode == "arithmetic":
        def Qij_diag(i, j):
            return np.maximum(0.5 * (Wi_list[i] + Wi_list[j]), _EPS)
    else:
        raise ValueError("q_mode must be 'harmonic' or 'arithmetic'")
    # lets the CUDA engine form Q_ij on the device from the V uploaded W vectors instead of calling back into Python
    # for every directed edge (same arithmetic, fp32)
    Qij_diag._admm_b200_spec = (q_mode, Wi_list)
    return Wi_list, Qij_diag


def _precompute_q_cache(num_nodes, Qij_diag):
    """block_3_graph_and_precisions.py:48-58."""
    return {(i, j): Qij_diag(i, j) for i in range(num_nodes) for j in range(num_nodes) if i != j}


# ---- per-pixel strategies (:62-187), small problems only -------------------------------------------------
def _complete_weighted(w, V):
    Gp = nx.Graph()
    Gp.add_nodes_from(range(V))
    for i in range(V):
        for j in range(i + 1, V):
            Gp.add_edge(i, j, weight=float(w[i, j]))
    return Gp


def _pixel_edges(strategy, w, V, k, rng):
    """Connected edge set {(a, b)} on V nodes for one pixel from the symmetric weight matrix w (V x V, zero
    diagonal).  Same selection rules as the reference: `mst` = maximum spanning tree of the complete graph (:114-130);
    `chain` = consecutive pairs of one rng.permutation(V) (:134-146); `knn` = each node's k heaviest neighbours by
    np.argpartition, symmetrised, plus every maximum-spanning-tree edge when that is not connected (:62-110)."""
    if strategy == "mst":
        return set(nx.maximum_spanning_tree(_complete_weighted(w, V), weight="weight").edges())
    if strategy == "chain":
        order = rng.permutation(V)
        return {(int(order[t]), int(order[t + 1])) for t in range(V - 1)}
    if strategy == "knn":
        k_eff = min(k, V - 1)
        Gp = nx.Graph()
        Gp.add_nodes_from(range(V))
        if k_eff > 0:
            for i in range(V):
                cand = w[i, :].copy()
                cand[i] = -np.inf
                for j in np.argpartition(cand, -k_eff)[-k_eff:]:
                    Gp.add_edge(i, int(j))
        if not nx.is_connected(Gp):
            Gp.add_edges_from(nx.maximum_spanning_tree(_complete_weighted(w, V), weight="weight").edges())
        return set(Gp.edges())
    raise ValueError("strategy must be one of 'knn', 'mst', or 'chain'")


def _build_all_pixel_masks(q_cache, num_nodes, n, strategy="knn", k=2, seed=0):
    """block_3_graph_and_precisions.py:150-187: keep[V, V, n] bool, symmetric, connected at every pixel; weights are
    symmetrised per pixel (:170-172) and the rng is np.random.default_rng(seed) consumed pixel by pixel (:154)."""
    if num_nodes * num_nodes * n > 2 ** 28:
        raise MemoryError("per-pixel masks are a small-problem option (V*V*n bools); use a node-level strategy")
    rng = np.random.default_rng(seed)
    keep = np.zeros((num_nodes, num_nodes, n), dtype=bool)
    Wm = np.zeros((num_nodes, num_nodes, n))
    for (i, j), q in q_cache.items():
        Wm[i, j] = q
    for p in range(n):
        w = 0.5 * (Wm[:, :, p] + Wm[:, :, p].T)
        np.fill_diagonal(w, 0.0)
        for a, b in _pixel_edges(strategy, w, num_nodes, k, rng):
            keep[a, b, p] = keep[b, a, p] = True
    return keep


def _build_all_pixel_masks_device(Wi_list, q_mode, num_nodes, n, strategy="knn", k=2, seed=0, device=0):
    """The same keep[V, V, n] built on the GPU (csrc/pixel_masks.cu, C ABI `admm_pixel_masks`): one thread per pixel,
    q_ij[p] formed from the W vectors in float32 exactly like make_precisions, Kruskal in networkx's edge order, bit
    rows unpacked here.  Needs Q to be make_precisions' own provider (so that q_ij is a function of W) and V <= 32; the
    `chain` permutations are drawn on the host from np.random.default_rng(seed) pixel by pixel like the reference."""
    import ctypes
    import torch
    from admm_b200 import _native as nat
    nat.require_cuda()
    if num_nodes > 32:
        raise ValueError("the device mask builder packs a node's row into 32 bits (V <= 32)")
    if num_nodes * num_nodes * n > 2 ** 28:
        raise MemoryError("per-pixel masks are a small-problem option (V*V*n bools); use a node-level strategy")
    dev = torch.device(f"cuda:{device}")
    strat = {"knn": 0, "mst": 1, "chain": 2}[strategy]
    W = perm = None
    if strat == 2:
        rng = np.random.default_rng(seed)
        perm = torch.from_numpy(np.stack([rng.permutation(num_nodes) for _ in range(n)]).astype(np.uint8)).to(dev)
    else:
        W = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(w, dtype=np.float32).reshape(-1)
                                                           for w in Wi_list]))).to(dev)
    bits = torch.zeros(num_nodes, n, dtype=torch.int32, device=dev)
    nat.check(nat.lib().admm_pixel_masks(num_nodes, n, strat, int(k), 1 if q_mode == "harmonic" else 0,
                                         W.data_ptr() if W is not None else None,
                                         perm.data_ptr() if perm is not None else None, bits.data_ptr(),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "admm_pixel_masks")
    b = bits.cpu().numpy().view(np.uint32)                                   # [V][n], bit j of row i = keep[i, j, p]
    return ((b[:, None, :] >> np.arange(num_nodes, dtype=np.uint32)[None, :, None]) & 1).astype(bool)


def _union_graph(keep):
    """block_3_graph_and_precisions.py:201-206: node graph with an edge wherever any pixel keeps the pair."""
    V = keep.shape[0]
    G = nx.Graph()
    G.add_nodes_from(range(V))
    for i in range(V):
        for j in range(i + 1, V):
            if keep[i, j].any():
                G.add_edge(i, j)
    return G


def build_pixel_connected_Q_provider(base_dir="saved_operators_Incmp_Span", A_dense_list_pickle="A_dense_list.pkl",
                                     strategy="knn", k=2, seed=0, q_mode="arithmetic", verbose=True, plot_union=True,
                                     show_plots=False, output_dir="pixel_graphs_out", A_dense_list=None, p=0.1,
                                     mask_device="auto"):
    """block_3_graph_and_precisions.py:262-319.  Returns (G_union, Wi_list, Qij_diag_masked, keep).

    `A_dense_list` (list of operators) replaces the pickle the reference loads (:288-291); if it is None the pickle
    path is tried.  Node-level strategies return `keep=None` and an unmasked provider; per-pixel strategies behave
    like the reference.  Unlike the reference (:304-309) a graph is always returned.
    `mask_device`: "auto" (GPU mask builder when a device is present and V <= 32, else the host restatement),
    "device", "host"."""
    if A_dense_list is None:
        import os
        import pickle
        from admm_b200 import registry
        A_dense_list = registry.get(base_dir)          # operators block_2.load_odl_data built for this base_dir
        A_path = os.path.join(base_dir, A_dense_list_pickle)
        if A_dense_list is None and os.path.exists(A_path):
            with open(A_path, "rb") as f:
                A_dense_list = pickle.load(f)
        if A_dense_list is None:
            raise FileNotFoundError(f"no operators registered for {base_dir!r} and no pickle at {A_path}; call "
                                    "load_odl_data(base_dir=...) first or pass A_dense_list=<operators>")
    Wi_list, Qij_diag = make_precisions(A_dense_list, q_mode=q_mode)
    V, n = len(Wi_list), Wi_list[0].shape[0]
    if strategy in ("ring", "regular", "er", "complete", "path"):
        G = make_graph(strategy, V, seed=seed, p=p, degree=k if strategy == "regular" else 4)

        def Qij_diag_masked(i, j):
            if i == j:
                return np.zeros(n, dtype=float)
            return Qij_diag(i, j)
        Qij_diag_masked._admm_b200_spec = Qij_diag._admm_b200_spec
        return G, Wi_list, Qij_diag_masked, None
    keep = None
    if V <= 32 and mask_device != "host":
        # per-pixel graph work on the GPU (the reference's Python loop over pixels with networkx is its second hot loop,
        # SURVEY 3.1); falls back to the host restatement only when there is no CUDA device and "auto" was asked
        try:
            keep = _build_all_pixel_masks_device(Wi_list, q_mode, V, n, strategy=strategy, k=k, seed=seed)
        except RuntimeError:
            if mask_device == "device":
                raise
    q_cache = _LazyQ(Qij_diag)
    if keep is None:
        q_cache = _precompute_q_cache(V, Qij_diag)
        keep = _build_all_pixel_masks(q_cache, V, n, strategy=strategy, k=k, seed=seed)
    G_union = _union_graph(keep)

    def Qij_diag_masked(i, j):  # :312-317
        if i == j:
            return np.zeros(n, dtype=float)
        return np.where(keep[i, j, :], q_cache[(i, j)], 0.0)
    return G_union, Wi_list, Qij_diag_masked, keep


class _LazyQ(dict):
    """q_cache[(i, j)] evaluated on first use (the device mask builder does not need the V^2 host vectors)."""

    def __init__(self, fn):
        super().__init__()
        self._fn = fn

    def __missing__(self, key):
        self[key] = self._fn(*key)
        return self[key]
