"""Host-side logic of the product (no GPU): bit-exact integer work (angle partition, neighbour lists, node->GPU map,
per-pixel masks), the C-ABI library (loads, exports every declared symbol, fails loudly without a device), and that
the product never routes through the oracle."""
import ctypes
import os
import re

import numpy as np
import networkx as nx
import pytest

import admm_b200
from admm_b200 import _native as nat
from admm_b200.geometry import angle_split, graph_csr, make_graph, node_angles, node_to_gpu, shepp_logan, trig_table32
from admm_b200.sharding import build_shard_plan, cut_statistics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))


def test_angle_split_matches_reference_and_baseline_table():
    for row in GOLD["b2_split"]:
        N, V, M = int(row[0]), int(row[1]), int(row[2])
        M = admm_b200.default_angles_total(N) if M < 0 else M
        assert angle_split(M, V) == [int(v) for v in row[4:4 + V]]
    assert angle_split(180, 4) == [45] * 4
    assert angle_split(360, 16) == [23] * 8 + [22] * 8
    assert angle_split(720, 32) == [23] * 16 + [22] * 16
    assert angle_split(720, 64) == [12] * 16 + [11] * 48
    th = node_angles(720, 64)
    assert np.array_equal(np.concatenate(th), (np.arange(720) + 0.5) * np.pi / 720)   # aggregate == vstack
    lit = node_angles(90, 2, "reference_literal")
    assert np.array_equal(lit[0], (np.arange(45) + 0.5) * np.pi / 45) and np.array_equal(lit[0], lit[1])
    c, s = trig_table32(th[0])
    assert c.dtype == np.float32 and np.array_equal(c, np.cos(th[0]).astype(np.float32))


@pytest.mark.parametrize("kind,V", [("ring", 4), ("regular", 16), ("er", 64), ("path", 5)])
def test_neighbour_lists_bit_exact_vs_networkx(kind, V):
    G = make_graph(kind, V, seed=0)
    edges, ptr, idx, ed, end = graph_csr(G)
    assert [tuple(e) for e in edges] == [(min(i, j), max(i, j)) for i, j in G.edges()]   # block_6_ver2:39-40
    for i in range(V):
        nb = list(G.neighbors(i))                                                        # :87
        assert list(idx[ptr[i]:ptr[i + 1]]) == nb
        for k, j in zip(range(ptr[i], ptr[i + 1]), nb):
            assert tuple(edges[ed[k]]) == (min(i, j), max(i, j))
            assert end[k] == (0 if i < j else 1)
    assert nx.is_connected(G)
    if kind == "er":
        assert G.number_of_edges() == 201 and max(dict(G.degree()).values()) == 13      # SURVEY 8(d) (s = 1)
    if kind == "regular":
        assert all(d == 4 for _, d in G.degree())


def test_node_to_gpu_map_and_shard_plan():
    assert node_to_gpu(64, 8) == [i // 8 for i in range(64)]
    assert node_to_gpu(5, 2) == [0, 0, 0, 1, 1]
    assert node_to_gpu(16, 1) == [0] * 16
    G = make_graph("er", 64, seed=0)
    edges = graph_csr(G)[0]
    plans = [build_shard_plan(G, 8, r) for r in range(8)]
    assert sorted(g for p in plans for g in p.local_nodes) == list(range(64))
    owned_dual = sorted(le.e for p in plans for le in p.local_edges if le.owns_dual)
    assert owned_dual == list(range(len(edges)))                          # every edge's s2 counted exactly once
    for e, (i, j) in enumerate(edges):
        holders = [p.rank for p in plans if e in p.eslot]
        assert holders == sorted({node_to_gpu(64, 8)[i], node_to_gpu(64, 8)[j]})
    for p in plans:                                                       # both sides agree on the exchange order
        for q in p.peers:
            assert plans[q].exch[p.rank] == p.exch[q]
    st = cut_statistics(G, 8)
    assert st["cut"] == sum(p.n_cut for p in plans) // 2 and st["edges"] == 201
    one = build_shard_plan(G, 1, 0)
    assert one.n_cut == 0 and len(one.local_edges) == 201


def test_block3_masks_and_precisions_match_reference():
    import block_3_graph_and_precisions as b3
    A = list(GOLD["b3_A"])
    for mode in ("arithmetic", "harmonic"):
        Wi, Q = b3.make_precisions(A, q_mode=mode)
        assert np.array_equal(np.stack(Wi), GOLD[f"b3_W_{mode}"])
        qc = b3._precompute_q_cache(len(A), Q)
        for (i, j), q in qc.items():
            assert np.array_equal(q, GOLD[f"b3_Q_{mode}"][i, j])
        for strat in ("knn", "mst", "chain"):
            keep = b3._build_all_pixel_masks(qc, len(A), A[0].shape[1], strategy=strat, k=2, seed=123)
            assert np.array_equal(keep, GOLD[f"b3_keep_{mode}_{strat}"]), (mode, strat)
            # reference invariants (test_block3_structural.py:15-60): symmetric, connected at every pixel
            assert np.array_equal(keep, keep.transpose(1, 0, 2))
            for p in range(keep.shape[2]):
                assert nx.is_connected(nx.from_numpy_array(keep[:, :, p].astype(int)))
                ne = keep[:, :, p].sum() // 2
                assert ne == len(A) - 1 if strat in ("mst", "chain") else ne >= len(A) - 1
    G, Wi, Qm, keep = b3.build_pixel_connected_Q_provider(A_dense_list=A, strategy="mst")
    assert set(G.nodes()) == set(range(len(A))) and keep.shape == (len(A), len(A), A[0].shape[1])
    assert np.array_equal(Qm(0, 1), np.where(keep[0, 1], Q(0, 1), 0.0)) is not None
    G2, _, Q2, keep2 = b3.build_pixel_connected_Q_provider(A_dense_list=A, strategy="ring")
    assert keep2 is None and G2.number_of_edges() == len(A) and not np.any(Q2(0, 0))


def test_phantoms_match_reference():
    import Gen_Sino_Partitioned as gs
    for N in (32, 64):
        assert np.array_equal(gs.ConstIm(N), GOLD[f"ConstIm_{N}"])
        np.random.seed(7 + N)
        assert np.array_equal(gs.randIm(N), GOLD[f"randIm_{N}_seed{7 + N}"])
    img = shepp_logan(64)
    assert img.shape == (64, 64) and abs(img.max() - 1.0) < 1e-12 and img.min() >= -1e-12
    assert admm_b200.psnr(GOLD["psnr_a"], GOLD["psnr_b"]) == GOLD["psnr_val"][0]


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "admm_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(admm_\w+)\s*\(", hdr, flags=re.M))
    declared -= {"admm_plan", "admm_state", "admm_edge", "admm_pack_item", "admm_node_ctl"}
    assert len(declared) >= 20
    L = nat.lib()
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert set(nat.EXPORTS) <= declared
    assert L.admm_version() == 200
    assert ctypes.sizeof(nat.State) == L.admm_abi_sizeof(0) == 272
    assert L.admm_abi_sizeof(1) == 13 * 8 and L.admm_abi_sizeof(2) == 3 * 8 and L.admm_abi_sizeof(3) == 16


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert nat.lib().admm_device_count() == 0
    with pytest.raises(RuntimeError, match="no CUDA device"):
        admm_b200.Plan(32, [np.linspace(0.1, 3.0, 8)])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        admm_b200.RayTransformCUDA(32, np.linspace(0.1, 3.0, 8))(np.zeros((32, 32)))
    import block_4_tv_helpers as b4
    with pytest.raises(RuntimeError, match="no CUDA device"):
        b4._grad_forward_2d_from_vec(np.zeros(16), 4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "distributed-inverse-problem-admm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the fp64 oracle", "").replace("fp64 oracle", ""), os.path.join(dirpath, f)


def test_operator_shapes_without_device():
    op = admm_b200.RayTransformCUDA(64, node_angles(180, 4)[1])
    assert op.shape == (45 * 64, 64 * 64) and op.domain.shape == (64, 64) and op.range.shape == (45, 64)
    assert op.T.shape == (64 * 64, 45 * 64)
    e = op.domain.one()
    assert float((2.0 * e + e).asarray().sum()) == 3 * 64 * 64
    z = op.domain.zero()
    z.asarray().flat[5] = 1.0
    assert z.asarray().sum() == 1.0


def test_boundary_accepts_every_call_the_reference_makes():
    """tests/golden/reference_call_facts.json lists (via ast, from the reference sources) every call the reference's
    drivers and test scripts make to the boundary functions; the drop-in modules must bind all of them."""
    import inspect
    import json
    import block_2_load_odl_data as b2
    import block_3_graph_and_precisions as b3
    import block_4_tv_helpers as b4
    import block_4_tv_helpers_with_plot as b4p
    import block_5_node_problem as b5
    import block_6_admm_loop as b6s
    import block_6_admm_loop_ver2 as b6
    import Gen_Sino_Partitioned as gs
    facts = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_call_facts.json")))
    table = {"load_odl_data": b2.load_odl_data, "build_pixel_connected_Q_provider": b3.build_pixel_connected_Q_provider,
             "make_precisions": b3.make_precisions, "build_node_problem": b5.build_node_problem,
             "decentralized_admm": b6.decentralized_admm, "generate_sinogram": gs.generate_sinogram,
             "kt_subgrad_isotropic_tv_from_x": b4.kt_subgrad_isotropic_tv_from_x,
             "edge_map_from_vector": b4p.edge_map_from_vector}
    assert b6s.decentralized_admm is b6.decentralized_admm
    assert len(facts["calls"]) >= 25
    for c in facts["calls"]:
        f = table[c["func"]]
        kws = {k: None for k in c["keywords"]}
        inspect.signature(f).bind(*([None] * c["n_positional"]), **kws)        # raises TypeError on a mismatch
        if c["func"] == "decentralized_admm":                                   # **kwargs are filtered at run time
            named = set(inspect.signature(f).parameters) | set(b6._IGNORED)
            assert set(c["keywords"]) <= named, (c["file"], c["line"], set(c["keywords"]) - named)
    produced = set(b6s.__dict__["decentralized_admm"].__code__.co_consts) | set(admm_b200.solver.HIST_KEYS) | {
        "primal_res", "dual_res", "obj", "pri_per_node_history", "dual_per_node_history", "obj_per_node_history",
        "obj_total_history"}
    assert set(facts["history_keys"]) <= produced, set(facts["history_keys"]) - produced
    src = open(os.path.join(ROOT, "distributed-inverse-problem-admm_b200", "block_2_load_odl_data.py")).read()
    for key in facts["data_keys"]:
        assert f'"{key}"' in src, key


def test_phased_exchange_orders_match_on_both_sides():
    """send rows of rank a -> p are the recv rows of p <- a, phase by phase, for every pair (8 ranks, 64-node ER)."""
    from admm_b200.sharding import build_shard_plan, phase_bounds
    from oracle import oracle as O
    assert phase_bounds(8, 2) == [0, 4, 8] and phase_bounds(7, 2) == [0, 4, 7] and phase_bounds(1, 2) == [0, 1, 1]
    G = O.make_graph("er", 64, seed=0, p=0.1)
    for phases in (1, 2, 3):
        sps = [build_shard_plan(G, 8, r, phases) for r in range(8)]
        for a in range(8):
            assert sorted(sps[a].node_phase) == list(range(64))
            for p in sps[a].peers:
                assert sps[a].send_rows[p] == sps[p].recv_rows[a]
                assert sps[a].send_rows[p][-1] == len(sps[a].exch[p])
                sa = {le.sslot: (le.e, le.sphase) for le in sps[a].local_edges if le.peer == p}
                rb = {le.rslot: le.e for le in sps[p].local_edges if le.peer == a}
                assert {k: v[0] for k, v in sa.items()} == rb
                for k, (e, ph) in sa.items():
                    assert sps[a].send_rows[p][ph] <= k < sps[a].send_rows[p][ph + 1]


def test_balanced_mincut_partition():
    """partition_nodes: deterministic, balanced, never worse than the contiguous map; ShardPlans built on it are
    consistent (every edge held by the owners of its ends, exchange lists equal on both sides)."""
    from admm_b200.sharding import _map_score, build_shard_plan, partition_nodes
    from oracle import oracle as O
    G = O.make_graph("er", 64, seed=0, p=0.1)
    edges = graph_csr(G)[0]
    for world in (2, 4, 8):
        cont = partition_nodes(G, world, "contiguous")
        assert cont == node_to_gpu(64, world)
        nr = partition_nodes(G, world, "mincut")
        assert nr == partition_nodes(G, world, "auto") == partition_nodes(G, world, "mincut")   # deterministic
        assert sorted(nr.count(k) for k in range(world)) == [64 // world] * world               # balanced
        assert nr[0] == 0
        assert _map_score(edges, nr, world) <= _map_score(edges, cont, world)
        assert _map_score(edges, nr, world)[1] < 0.75 * _map_score(edges, cont, world)[1]        # ER-64: a real gain
        plans = [build_shard_plan(G, world, r, 1, nr) for r in range(world)]
        assert sorted(g for p in plans for g in p.local_nodes) == list(range(64))
        for e, (i, j) in enumerate(edges):
            holders = sorted(r for r in range(world) if e in plans[r].eslot)
            assert holders == sorted({nr[int(i)], nr[int(j)]})
        for a in range(world):
            for p in plans[a].peers:
                assert plans[a].exch[p] == plans[p].exch[a]
    assert partition_nodes(G, 1, "mincut") == [0] * 64
    big = O.make_graph("ring", 300, seed=0)
    assert partition_nodes(big, 4, "auto") == node_to_gpu(300, 4)                                # auto: V > 256 stays contiguous
    with pytest.raises(ValueError):
        partition_nodes(G, 2, "metis")


def test_block3_checker_invariants():
    """The quantitative invariants the reference checks by hand (test_block_3_checker.py:53-124), as asserts:
    total per-pixel edge counts (MST / chain: n(V-1); kNN: between n(V-1) and n*min(Vk, V(V-1)/2)), the harmonic
    bound sum_p keep*Q_ij <= sum_p min(W_i, W_j), and degree conservation of the count / weight matrices."""
    import block_3_graph_and_precisions as b3
    A = list(GOLD["b3_A"])
    V, n, k = len(A), A[0].shape[1], 2
    Wi, Q = b3.make_precisions(A, q_mode="harmonic")
    qc = b3._precompute_q_cache(V, Q)
    for strat in ("mst", "chain", "knn"):
        keep = b3._build_all_pixel_masks(qc, V, n, strategy=strat, k=k, seed=123)
        count = keep.sum(axis=2).astype(np.int64)
        wsum = np.zeros((V, V))
        for i in range(V):
            for j in range(V):
                if i != j:
                    wsum[i, j] = float(np.sum(np.where(keep[i, j], qc[(min(i, j), max(i, j))], 0.0)))
        pairs = int(np.triu(count, 1).sum())
        if strat == "knn":
            assert n * (V - 1) <= pairs <= n * min(V * k, V * (V - 1) // 2)
        else:
            assert pairs == n * (V - 1)
        for i in range(V):
            for j in range(i + 1, V):
                assert wsum[i, j] <= float(np.sum(np.minimum(Wi[i], Wi[j]))) + 1e-12
        assert count.sum() == 2 * np.triu(count, 1).sum()
        assert np.isclose(wsum.sum(), 2.0 * np.triu(wsum, 1).sum())


def test_cfg4_node_to_gpu_maps_fixture():
    """The default (angle-balanced min-cut) node -> GPU maps of BASELINE cfg 4 (64 nodes, ER p = 0.1, 720 angles) at 2, 4
    and 8 GPUs are integer work: bit-exact against the committed maps (tests/golden/cfg4_node_maps.json), 64 / G nodes and
    720 / G angle rows per rank."""
    import json
    from admm_b200 import make_graph
    from admm_b200.sharding import cut_statistics, partition_nodes
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "cfg4_node_maps.json")))
    G = make_graph("er", 64, seed=0, p=0.1)
    assert G.number_of_edges() == fx["edges"] == 201
    w = [len(t) for t in node_angles(720, 64)]
    for world in (2, 4, 8):
        m = fx["maps"][str(world)]
        nr = partition_nodes(G, world, "auto", weights=w)
        assert [int(v) for v in nr] == m["node_rank"]
        assert cut_statistics(G, world, nr)["cut"] == m["cut_edges"]
        assert [nr.count(r) for r in range(world)] == [64 // world] * world
        assert [sum(w[i] for i in range(64) if nr[i] == r) for r in range(world)] == [720 // world] * world
