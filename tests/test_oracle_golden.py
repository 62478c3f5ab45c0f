"""The oracle (and the product's host-side NumPy logic) against fixtures produced by the REFERENCE's own code
(tests/golden/make_golden.py ran the reference modules with odl/cvxpy/matplotlib stubbed)."""
import os

import numpy as np
import networkx as nx
import pytest

from oracle import oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))


@pytest.mark.parametrize("N", [5, 16])
def test_block4_helpers_bit_exact(N):
    x, xf, px, py = G[f"b4_N{N}_x"], G[f"b4_N{N}_xf"], G[f"b4_N{N}_px"], G[f"b4_N{N}_py"]
    gx, gy = O.grad_forward(x, N)
    assert np.array_equal(gx, G[f"b4_N{N}_gx"]) and np.array_equal(gy, G[f"b4_N{N}_gy"])
    assert np.array_equal(O.div_reference(px, py, N), G[f"b4_N{N}_div"])
    assert np.array_equal(O.kt_subgrad(x, N), G[f"b4_N{N}_kt"])
    assert np.array_equal(O.kt_subgrad(xf, N), G[f"b4_N{N}_ktf"])
    # C twins of the numpy restatements
    n = N * N
    g1, g2 = np.empty(n), np.empty(n)
    O.lib().orc_grad(O._p(np.ascontiguousarray(x)), N, O._p(g1), O._p(g2))
    assert np.array_equal(g1.reshape(N, N), gx) and np.array_equal(g2.reshape(N, N), gy)
    # the shipped divergence is K^T in the interior and sign-flipped on the border
    exact = O.grad_T(px, py, N).reshape(N, N)
    ref = G[f"b4_N{N}_div"].reshape(N, N)
    assert np.allclose(exact[1:-1, 1:-1], ref[1:-1, 1:-1])
    assert not np.allclose(exact, ref)
    assert abs(O.tv_canonical(x, N) - G[f"b4_N{N}_edge_raw"].sum()) < 1e-12


def test_block3_precisions_bit_exact():
    A = list(G["b3_A"])
    for mode in ("arithmetic", "harmonic"):
        Wi, Q = O.make_precisions([np.sum(a * a, axis=0) for a in A], mode)
        assert np.array_equal(np.stack(Wi), G[f"b3_W_{mode}"])
        V = len(A)
        for i in range(V):
            for j in range(V):
                if i != j:
                    assert np.array_equal(Q(i, j), G[f"b3_Q_{mode}"][i, j])


def test_angle_split_bit_exact():
    for row in G["b2_split"]:
        N, V, M, agg = (int(v) for v in row[:4])
        M = O.default_angles_total(N) if M < 0 else M
        assert agg == M
        assert O.angle_split(M, V) == [int(v) for v in row[4:4 + V]]


def test_phantoms_and_psnr_bit_exact():
    for N in (32, 64):
        assert np.array_equal(O.ConstIm(N), G[f"ConstIm_{N}"])
        np.random.seed(7 + N)
        assert np.array_equal(O.randIm(N), G[f"randIm_{N}_seed{7 + N}"])
    a, b = G["psnr_a"], G["psnr_b"]
    assert O.psnr(a, b) == G["psnr_val"][0] and O.psnr(a, b, data_range=2.5) == G["psnr_val"][1]


class _DenseOp:
    def __init__(self, A, N):
        self.A, self.N, self.D, self.nang = np.asarray(A, dtype=np.float64), N, N, A.shape[0] // N
        self.shape = A.shape

    def forward(self, v):
        return self.A @ v

    def adjoint(self, q):
        return self.A.T @ q


@pytest.mark.parametrize("tag", ["ring4", "irr5"])
def test_outer_loop_matches_reference_block6(tag):
    """block_6_admm_loop_ver2.decentralized_admm (reference code, executed) vs the oracle's array restatement on the
    same dense operators, graph, Q provider and x-update."""
    N, M, S, C, mu, lam, rho = G["b6_params"]
    N, S, C = int(N), int(S), int(C)
    edges = [tuple(int(v) for v in e) for e in G[f"b6_{tag}_edges"]]
    # same constructions as tests/golden/make_golden.py (adjacency insertion order decides edges()/neighbors() order)
    Gr = nx.cycle_graph(4) if tag == "ring4" else nx.Graph([(0, 3), (3, 1), (1, 4), (4, 0), (2, 3), (2, 1)])
    assert [tuple(e) for e in Gr.edges()] == edges
    rows = G[f"b6_{tag}_rows"]
    dense = [G[f"b6_{tag}_dense"][i][: rows[i]] for i in range(len(rows))]
    sino = np.split(G[f"b6_{tag}_sino"], np.cumsum(rows)[:-1])
    ops = [_DenseOp(d, N) for d in dense]
    Wi, Q = O.make_precisions([np.sum(d * d, axis=0) for d in dense], "arithmetic")
    iters = len(G[f"b6_{tag}_primal"])
    x, h = O.decentralized_admm(ops, sino, Gr, Wi, Q, N, lam_tv=lam, rho=rho, max_iters=iters, eps_pri=1e-9,
                                eps_dual=1e-9, phantom_true=O.shepp_logan(N), tv_mu=mu, tv_sweeps=S, cg_iters=C,
                                x_update_fn=O.np_x_update)
    for key in ("primal", "dual", "pri_per_node", "dual_per_node", "obj_per_node", "obj_total", "mse_sino_per_node",
                "mse_sino_total", "img_mse_per_node", "img_mse_total", "eps_target_history", "eps_used_history"):
        assert np.allclose(np.array(h[key]), G[f"b6_{tag}_{key}"], rtol=1e-9, atol=1e-12), key
    # a14 (block_6_admm_loop_ver2.py:100-176): the reference really retried -- the solves it issued per (iteration,
    # node), and the eps it handed to each, are what the oracle's restatement of the rule produces
    solves = G[f"b6_{tag}_solve_eps"]
    per = np.zeros((iters, len(rows)), dtype=np.int64)
    for k, node, eps in solves:
        assert np.isclose(eps, min(1e-2, 2.0 / ((k + 1) ** 1.005)) / 5.0 ** per[int(k), int(node)], rtol=1e-12)
        per[int(k), int(node)] += 1
    assert np.array_equal(per - 1, np.array(h["tighten_history"]))
    assert per.max() == 3 and per.min() == 1          # the fixture exercises accept-at-once and the tighten cap
    assert np.allclose(np.array(h["g_norm_history"]), G[f"b6_{tag}_g_norm_history"], rtol=1e-7)
    assert np.allclose(np.stack(x), G[f"b6_{tag}_x"], rtol=1e-10, atol=1e-13)
    # the C x-update with the matrix-free projector agrees with the dense float32 matrices to fp32 rounding
    thetas = O.node_angles(int(M), len(rows))
    ops_c = [O.JosephOperator(N, t) for t in thetas]
    x2, h2 = O.decentralized_admm(ops_c, sino, Gr, Wi, Q, N, lam_tv=lam, rho=rho, max_iters=iters, eps_pri=1e-9,
                                  eps_dual=1e-9, phantom_true=O.shepp_logan(N), tv_mu=mu, tv_sweeps=S, cg_iters=C)
    assert np.allclose(np.array(h2["primal"]), G[f"b6_{tag}_primal"], rtol=1e-5)
    assert np.allclose(np.array(h2["dual"]), G[f"b6_{tag}_dual"], rtol=1e-5)


def test_cfg1_default_schedule_fixture_is_the_oracles():
    """tests/golden/cfg1_default_schedule_200.npz (the 200-iteration fp64 run behind the GPU default-schedule parity test,
    written by `tools/carry_study.py 128 200 default-only oracle-only`) starts like the oracle run of the same inputs."""
    import os
    N, M, V = 128, 180, 4
    thetas = O.node_angles(M, V, "contiguous")
    img = O.shepp_logan(N)
    ops = [O.JosephOperator(N, t) for t in thetas]
    sinos = [(op.forward(img) + 0.005 * np.random.default_rng(1234 + i).standard_normal(op.shape[0]))
             .reshape(op.nang, op.D).astype(np.float32) for i, op in enumerate(ops)]
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg1_default_schedule_200.npz"))
    x, h = O.decentralized_admm(ops, sinos, O.make_graph("ring", V), None, None, N, uniform_q=1.0, lam_tv=0.02, rho=2.0,
                                max_iters=3, eps_pri=0.0, eps_dual=0.0, phantom_true=img, tv_sweeps=1, cg_iters=2,
                                acceptance=True)
    assert np.allclose(h["primal"], ref["primal"][:3], rtol=1e-9) and np.allclose(h["dual"], ref["dual"][:3], rtol=1e-9)
    assert np.array_equal(np.array(h["tighten_history"]), ref["tighten"][:3])
    assert ref["primal"].shape == (200,) and ref["x"].shape == (V, N * N)
