"""Drop-in for the reference's block_5_node_problem.py: `build_node_problem` keeps its signature and returns
CVXPY-shaped handles `(xi, prob)`, but `prob.solve()` runs the TV-split + CG x-update on the GPU (C ABI
`admm_rhs0` + `admm_x_update`) instead of canonicalising to SCS.  CUDA-only."""
from __future__ import annotations

import numpy as np

from admm_b200.solver import NodeProblem


class _Variable:
    """Stands for `cp.Variable(n)`: `.value` is the solution vector after `prob.solve()` (block_6_ver2:135)."""

    def __init__(self, n):
        self.shape = (n,)
        self.size = n
        self.value = None


class _SolverStats:
    def __init__(self):
        self.num_iters = None
        self.solve_time = None
        self.solver_name = "ADMM_B200_TVSPLIT_CG"


class _Problem:
    """Stands for `cp.Problem`: `.solve(**kw)`, `.value`, `.status`, `.solver_stats.num_iters`
    (block_6_admm_loop_ver2.py:123-135)."""

    def __init__(self, xi, node):
        self._xi, self._node = xi, node
        self.value = None
        self.status = None
        self.solver_stats = _SolverStats()

    def solve(self, solver=None, eps=1e-4, eps_abs=None, eps_rel=None, max_iters=200, verbose=False, warm_start=True,
              cg_iters=8, **ignored):
        """`eps` -> stationarity target on |g| relative to |A^T b| (the a14 quantity); `max_iters` caps the total CG
        iterations.  SCS-only kwargs (acceleration_lookback, use_indirect, alpha, scale, ...) are ignored."""
        import time
        t0 = time.perf_counter()
        node = self._node
        tol = min(v for v in (eps, eps_abs, eps_rel) if v is not None)
        scale = float(node.atb.norm().item()) or 1.0
        done0 = node.cg_done
        status = "optimal_inaccurate"
        cg = max(1, min(int(cg_iters), int(max_iters)))
        prev = None
        while node.cg_done - done0 + cg <= max(int(max_iters), cg):
            node.sweep(cg)
            obj, gn = node.stats()
            if verbose:
                print(f"[admm_b200] sweep cg={node.cg_done - done0} obj={obj:.6e} |g|={gn:.3e}")
            if prev is not None and abs(prev - obj) <= tol * max(1.0, abs(obj)) and gn <= max(tol, 1e-6) * scale * 10:
                status = "optimal"
                break
            prev = obj
        obj, gn = node.stats()
        self._xi.value = node.x_value()
        self.value = float(obj)
        self.status = status
        self.g_norm = gn
        self.solver_stats.num_iters = node.cg_done - done0
        self.solver_stats.solve_time = time.perf_counter() - t0
        return self.value


def build_node_problem(Ai, bi, rho, neighbor_terms, N, lam_tv, Qij_terms, tv_mu=None):
    """
    Ai: matrix-free operator of shape (m_i, n)  [the reference passes a dense matrix, block_5_node_problem.py:8]
    bi: vector shape (m_i,)
    neighbor_terms: list of vectors v_ij = z_ij - y_ij,i for each neighbor j
    Qij_terms: list of q_ij diagonal vectors shape (n,) to weight the squared norms

    Objective (block_5_node_problem.py:14-16):
        0.5*||Ai xi - bi||_2^2 + lam_tv * TV(xi) + (rho/2) * sum_j || xi - v_ij ||_{Qij}^2
    TV is the canonical isotropic TV of block_4's NumPy helpers.
    """
    n = Ai.shape[1]
    if n != N * N:
        raise ValueError("operator domain size does not match N")
    xi = _Variable(n)
    node = NodeProblem(Ai, np.asarray(bi).reshape(-1), rho, list(neighbor_terms), N, lam_tv, list(Qij_terms), tv_mu=tv_mu)
    return xi, _Problem(xi, node)
