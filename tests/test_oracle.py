"""Independent anchors of the oracle: the projector against closed-form Radon transforms and exact-transpose checks,
the x-update against the optimality conditions of eq. (1), the loop against its fixed point."""
import numpy as np
import pytest

from oracle import oracle as O


def test_c_forward_equals_numpy_twin_and_dense_transpose():
    rng = np.random.default_rng(0)
    for N, thetas, D, dw in ((24, O.node_angles(90, 3, "reference_literal")[0], None, 2.0),   # has theta = pi/4 tie
                             (20, O.node_angles(37, 1)[0], 33, 2.6), (16, O.node_angles(20, 1)[0], 40, 2.0)):
        op = O.JosephOperator(N, thetas, D, dw)
        X = rng.standard_normal((N, N))
        q = rng.standard_normal(op.shape[0])
        assert np.abs(op.forward(X) - O.np_forward(X, op.c, op.s, op.D, dw).reshape(-1)).max() < 1e-13
        A = op.dense()
        assert np.abs(A @ X.reshape(-1) - op.forward(X)).max() < 1e-12
        assert np.abs(A.T @ q - op.adjoint(q)).max() < 1e-12           # matched transpose
        assert np.abs((A * A).sum(0) - op.colnorm2()).max() < 1e-13    # block_3:22
    assert np.any(np.isclose(np.abs(np.cos(O.node_angles(90, 3, "reference_literal")[0])),
                             np.abs(np.sin(O.node_angles(90, 3, "reference_literal")[0]))))


def test_forward_converges_to_analytic_ellipse_sinogram():
    errs = []
    for N in (64, 128, 256):
        th = O.node_angles(60, 1)[0]
        sino = O.JosephOperator(N, th).forward(O.shepp_logan(N)).reshape(60, N)
        ana = O.ellipse_sinogram(th, N, O._SHEPP_LOGAN_MODIFIED)
        errs.append(np.linalg.norm(sino - ana) / np.linalg.norm(ana))
    assert errs[0] > errs[1] > errs[2] and errs[2] < 0.03
    # a smooth (bandlimited) test function is integrated to high accuracy: Gaussian blob, closed form
    N, sig = 256, 0.15
    g = -1 + (np.arange(N) + 0.5) * 2 / N
    X, Y = np.meshgrid(g, g, indexing="ij")
    img = np.exp(-((X - 0.2) ** 2 + (Y + 0.1) ** 2) / (2 * sig ** 2))
    th = O.node_angles(45, 1)[0]
    s = -1 + (np.arange(N) + 0.5) * 2 / N
    s0 = 0.2 * np.cos(th) - 0.1 * np.sin(th)
    ana = np.sqrt(2 * np.pi) * sig * np.exp(-(s[None, :] - s0[:, None]) ** 2 / (2 * sig ** 2))
    sino = O.JosephOperator(N, th).forward(img).reshape(45, N)
    assert np.linalg.norm(sino - ana) / np.linalg.norm(ana) < 2e-4


def test_chord_lengths_of_the_unit_square():
    N = 64
    s = O.JosephOperator(N, [1e-9, np.pi / 2, np.pi / 4]).forward(np.ones((N, N))).reshape(3, N)
    assert np.allclose(s[:2, 1:-1], 2.0, atol=1e-6)
    sj = -1 + (np.arange(N) + 0.5) * 2 / N
    assert np.allclose(s[2], 2 * (np.sqrt(2) - np.abs(sj)), atol=0.05)


def test_x_update_c_equals_numpy_twin_and_satisfies_optimality():
    N, n = 16, 256
    op = O.JosephOperator(N, O.node_angles(40, 2)[1])
    rng = np.random.default_rng(3)
    b = op.forward(O.shepp_logan(N)) + 0.01 * rng.standard_normal(op.shape[0])
    v = [0.3 * rng.standard_normal(n) for _ in range(3)]
    q = [0.5 + rng.random(n) for _ in range(3)]
    rho, lam, mu = 2.0, 0.05, 1.0
    rhs0 = op.adjoint(b) + rho * sum(qi * vi for qi, vi in zip(q, v))
    rhoD = rho * sum(q)
    xs = []
    for fn in (O.x_update, O.np_x_update):
        x, d, w = np.zeros(n), np.zeros(2 * n), np.zeros(2 * n)
        for _ in range(3):
            Ax, r, tvrhs = fn(op, 1.0, rhs0, rhoD, mu, lam, 2, 5, x, d, w)
        xs.append((x.copy(), d.copy(), w.copy(), Ax.copy(), r.copy(), tvrhs.copy()))
    for a, c in zip(xs[0], xs[1]):
        assert np.abs(a - c).max() < 1e-12
    # run to convergence: x minimises eq. (1)  <=>  0 in A^T(Ax-b) + rho(Dx - sum q v) + lam K^T s, s in dTV(Kx);
    # at the split-Bregman fixed point s = (mu/lam) w.
    x, d, w = np.zeros(n), np.zeros(2 * n), np.zeros(2 * n)
    for _ in range(300):
        Ax, r, tvrhs = O.x_update(op, 1.0, rhs0, rhoD, mu, lam, 1, 12, x, d, w)
    grad = op.adjoint(op.forward(x) - b) + rhoD * x - rho * sum(qi * vi for qi, vi in zip(q, v))
    sub = O.grad_T((mu / lam) * w[:n].reshape(N, N), (mu / lam) * w[n:].reshape(N, N), N)
    assert np.linalg.norm(grad + lam * sub) < 1e-8 * np.linalg.norm(op.adjoint(b))
    s_norm = np.sqrt(w[:n] ** 2 + w[n:] ** 2) * mu / lam
    assert s_norm.max() <= 1 + 1e-8                       # dual feasibility |s| <= 1
    gx, gy = O.grad_forward(x, N)
    act = np.sqrt(gx ** 2 + gy ** 2).reshape(-1) > 1e-6
    assert np.allclose(s_norm[act], 1.0, atol=1e-5)       # |s| = 1 where the gradient is non-zero
    # and any perturbation increases the objective
    def F(xx):
        return (0.5 * np.sum((op.forward(xx) - b) ** 2) + lam * O.tv_canonical(xx, N)
                + 0.5 * rho * sum(np.sum(qi * (xx - vi) ** 2) for qi, vi in zip(q, v)))
    f0 = F(x)
    for _ in range(5):
        assert F(x + 1e-3 * rng.standard_normal(n)) > f0


def test_admm_fixed_point_on_consistent_data():
    """Noise-free, lam = 0: every x_i -> x_true is not guaranteed with few angles, but consensus is: residuals -> 0
    and all nodes agree; with the full angle set split over nodes the consensus point reproduces the phantom."""
    N, V = 16, 3
    thetas = O.node_angles(48, V)
    ops = [O.JosephOperator(N, t) for t in thetas]
    img = O.shepp_logan(N)
    sinos = [op.forward(img).reshape(op.nang, N) for op in ops]
    Wi, Q = O.make_precisions([op.colnorm2() for op in ops])
    x, h = O.decentralized_admm(ops, sinos, O.make_graph("complete", V), Wi, Q, N, lam_tv=0.0, rho=1.0,
                                max_iters=400, eps_pri=0, eps_dual=0, tv_sweeps=1, cg_iters=10, phantom_true=img, tv_mu=1.0)
    assert h["primal"][-1] < 1e-3 * h["primal"][0] and h["dual"][-1] < 1e-3 * max(h["dual"])
    assert max(np.linalg.norm(x[i] - x[0]) for i in range(V)) < 1e-3 * np.linalg.norm(x[0])
    assert O.psnr(x[0].reshape(N, N), img) > 22.0
    assert h["img_mse_total"][-1] < h["img_mse_total"][10]


def test_tv_pairing_quirk_is_documented_not_reproduced():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(25)
    assert abs(O.tv_reference_pairing(x, 5) - O.tv_canonical(x, 5)) > 1e-3


def test_pdhg_variant_pieces():
    """PDHG consensus twin (ADMM_Tomo_Only.py:89-148): the gradient pair is an exact adjoint pair, the operator-norm
    estimate bounds |L x| / |x|, and the aggregate PDHG iteration decreases its objective."""
    N = 16
    rng = np.random.default_rng(3)
    x, p1, p2 = rng.standard_normal(N * N), rng.standard_normal((N, N)), rng.standard_normal((N, N))
    g1, g2 = O.odl_grad(x, N)
    assert abs((g1 * p1).sum() + (g2 * p2).sum() - x @ O.odl_grad_adjoint(p1, p2, N)) < 1e-10
    # zero padding beyond the last index: the last row / column differences are -x / h
    assert np.allclose(g1[-1], -x.reshape(N, N)[-1] * N / 2) and np.allclose(g2[:, -1], -x.reshape(N, N)[:, -1] * N / 2)
    ops = [O.JosephOperator(N, t) for t in O.node_angles(24, 3)]
    img = O.shepp_logan(N)
    b = [o.forward(img) for o in ops]
    adj = (np.pi / ops[0].nang) * (2.0 / N) / (2.0 / N) ** 2
    nrm = O.pdhg_opnorm(ops[0], adj, N, 40)
    for _ in range(5):
        v = rng.standard_normal(N * N)
        g1, g2 = O.odl_grad(v, N)
        Lv2 = v @ (adj * ops[0].adjoint(ops[0].forward(v)) + O.odl_grad_adjoint(g1, g2, N))   # <v, L* L v>
        assert np.sqrt(Lv2) <= nrm * np.linalg.norm(v) * (1 + 1e-3)

    def objective(xv):
        g1, g2 = O.odl_grad(xv, N)
        return sum(np.sum((o.forward(xv) - bi) ** 2) for o, bi in zip(ops, b)) + 0.005 * np.sqrt(g1 ** 2 + g2 ** 2).sum()
    r5 = O.pdhg_consensus(ops, b, img, N, niter=5)
    r40 = O.pdhg_consensus(ops, b, img, N, niter=40)
    assert objective(r40["x_agg"]) < objective(r5["x_agg"]) < objective(np.zeros(N * N))
    assert r40["mse_agg_list"][-1] < r5["mse_agg_list"][-1]
