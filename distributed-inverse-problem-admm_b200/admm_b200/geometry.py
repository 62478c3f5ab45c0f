"""Host-side integer / geometry logic of the hot path (bit-exact parts of north_star): the angle partition,
neighbour lists, node->GPU map, trig table, phantoms.  Pure NumPy; no device work here."""
from __future__ import annotations

import math

import numpy as np


def default_angles_total(N: int) -> int:
    """block_2_load_odl_data.py:31-33."""
    return max(180, 3 * N)


def angle_split(angles_total: int, num_nodes: int) -> list:
    """block_2_load_odl_data.py:36-38: equal split, the first M % V nodes get one more."""
    per = [angles_total // num_nodes] * num_nodes
    for i in range(angles_total % num_nodes):
        per[i] += 1
    return per


def node_angles(angles_total: int, num_nodes: int, partition: str = "contiguous") -> list:
    """Per-node angles in radians (fp64).

    contiguous (default, north_star "angle-partitioned"): node k owns the k-th contiguous block of the aggregate
    midpoint grid theta_a = (a + 1/2) pi / M, so the aggregate operator is the vstack of the node operators.
    reference_literal: block_2_load_odl_data.py:51 as shipped: every node gets the midpoints of
    uniform_partition(0, pi, m_k)."""
    out, start = [], 0
    for m_k in angle_split(angles_total, num_nodes):
        if partition == "contiguous":
            out.append((np.arange(start, start + m_k, dtype=np.float64) + 0.5) * math.pi / angles_total)
        elif partition == "reference_literal":
            out.append((np.arange(m_k, dtype=np.float64) + 0.5) * math.pi / m_k)
        else:
            raise ValueError(f"unknown partition {partition!r}")
        start += m_k
    return out


def trig_table32(theta) -> tuple:
    """(cos, sin) evaluated in fp64 and rounded once to fp32."""
    theta = np.asarray(theta, dtype=np.float64)
    return (np.ascontiguousarray(np.cos(theta).astype(np.float32)),
            np.ascontiguousarray(np.sin(theta).astype(np.float32)))


def node_to_gpu(num_nodes: int, num_gpus: int) -> list:
    """deterministic contiguous blocks gpu(i) = (i * G) // V."""
    return [(i * num_gpus) // num_nodes for i in range(num_nodes)]


def graph_csr(G):
    """Edge list in G.edges() order keyed (min, max) (block_6_admm_loop_ver2.py:39-40) and neighbour lists in
    G.neighbors(i) order (:87):  edges[E,2], nbr_ptr[V+1], nbr_idx[nnz], nbr_edge[nnz], nbr_end[nnz]
    (nbr_end = 0 when i is the min end of the edge)."""
    V = G.number_of_nodes()
    if sorted(G.nodes()) != list(range(V)):
        raise ValueError("graph nodes must be the integers 0..V-1 (block_6_admm_loop_ver2.py:81)")
    edges = [(min(i, j), max(i, j)) for i, j in G.edges()]
    if any(i == j for i, j in edges):
        raise ValueError("self loops are not supported")
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


def make_graph(kind: str, V: int, seed: int = 0, p: float = 0.1, degree: int = 4):
    """Node-level graphs named by the BASELINE configs."""
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
    raise ValueError(f"unknown graph kind {kind!r}")


_SHEPP_LOGAN_MODIFIED = [
    (1.0, .69, .92, 0.0, 0.0, 0.0), (-.8, .6624, .8740, 0.0, -.0184, 0.0), (-.2, .1100, .3100, .22, 0.0, -18.0),
    (-.2, .1600, .4100, -.22, 0.0, 18.0), (.1, .2100, .2500, 0.0, .35, 0.0), (.1, .0460, .0460, 0.0, .1, 0.0),
    (.1, .0460, .0460, 0.0, -.1, 0.0), (.1, .0460, .0230, -.08, -.605, 0.0), (.1, .0230, .0230, 0.0, -.606, 0.0),
    (.1, .0230, .0460, .06, -.605, 0.0)]


def shepp_logan(N: int) -> np.ndarray:
    """Modified Shepp-Logan on [-1,1]^2 sampled at pixel centres, array [ix, iy], values in [0,1]."""
    h = 2.0 / N
    g = -1.0 + (np.arange(N) + 0.5) * h
    X, Y = np.meshgrid(g, g, indexing="ij")
    img = np.zeros((N, N))
    for v, a, b, x0, y0, phi in _SHEPP_LOGAN_MODIFIED:
        ph = math.radians(phi)
        xr = (X - x0) * math.cos(ph) + (Y - y0) * math.sin(ph)
        yr = -(X - x0) * math.sin(ph) + (Y - y0) * math.cos(ph)
        img[(xr / a) ** 2 + (yr / b) ** 2 <= 1.0] += v
    return img


def psnr(x_hat, x_true, data_range=1.0):
    """test_final_integration.py:41-45."""
    mse = np.mean((np.asarray(x_hat, dtype=np.float64) - np.asarray(x_true, dtype=np.float64)) ** 2)
    if mse == 0:
        return float("inf")
    return 20.0 * np.log10(data_range) - 10.0 * np.log10(mse)
