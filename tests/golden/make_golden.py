"""Generate tests/golden/reference_golden.npz by importing the REFERENCE's own Python modules from
/root/reference (read-only, only available in the build container) with their third-party imports stubbed
(odl, cvxpy, matplotlib are not installed).  Run once here; the fixtures are committed, this script documents
how they were made.  Nothing under tests/ reads /root/reference at test time.

    python tests/golden/make_golden.py

What is pinned (reference code executed verbatim):
  * block_4_tv_helpers: _grad_forward_2d_from_vec, _div_backward_2d_to_vec, kt_subgrad_isotropic_tv_from_x
  * block_4_tv_helpers_with_plot: edge_map_from_vector
  * block_3_graph_and_precisions: make_precisions (both q modes), _build_all_pixel_masks (knn / mst / chain)
  * block_2_load_odl_data._build_parallel_beam_operators: the integer angle split and the partitions it requests
    from odl (captured by a recording stub)
  * Gen_Sino_Partitioned: ConstIm, randIm (seeded global RNG)
  * test_final_integration.psnr
  * block_6_admm_loop_ver2.decentralized_admm: the whole outer loop (init, neighbour assembly order, acceptance
    logic, metrics, z / y updates, residuals, history) with block_5's CVXPY problem replaced by a stub whose
    solve() is one solve of the oracle's TV-split + CG x-update on the same dense matrices.  EVERY solve() call does
    that work, warm-started from the previous call's (x, d, w) like SCS's warm_start=True on the same cp.Problem, so
    the reference's accept / tighten-and-retry loop (:100-176) really runs: the fixtures hold its eps_used_history and
    the eps values it handed to solve() (b6_*_solve_eps: [outer k, node, eps] per call)
What is NOT pinned (parity unpinned): odl.tomo.RayTransform's discretisation and the SCS solutions.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


# ---- stubs for absent third parties ---------------------------------------------------------------------
class _Anything:
    def __getattr__(self, k):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


plt = stub("matplotlib.pyplot", **{k: (lambda *a, **kw: _Anything()) for k in
                                   ("figure", "imshow", "title", "axis", "tight_layout", "savefig", "close", "show",
                                    "subplots", "plot", "legend", "xlabel", "ylabel", "grid", "hist", "colorbar")})
stub("matplotlib", pyplot=plt, use=lambda *a, **k: None)
stub("cvxpy", SCS="SCS", Variable=None, Problem=None, Minimize=None)

ODL_CALLS = []


class _Partition:
    def __init__(self, a, b, m):
        self.a, self.b, self.m = a, b, m


def _uniform_partition(a, b, m):
    ODL_CALLS.append(("uniform_partition", float(a), float(b), int(m)))
    return _Partition(a, b, m)


class _Geom:
    def __init__(self, pa, pd):
        self.angles, self.det = pa, pd


class _RT:
    def __init__(self, space, geom, impl=None):
        self.space, self.geom, self.impl = space, geom, impl


odl = stub("odl", uniform_discr=lambda **kw: ("space", kw), uniform_partition=_uniform_partition,
           tomo=types.SimpleNamespace(Parallel2dGeometry=_Geom, RayTransform=_RT))
stub("odl.tomo", Parallel2dGeometry=_Geom, RayTransform=_RT)

sys.path.insert(0, REF)
out = {}
rng = np.random.default_rng(20261018)

# ---- block_4 ----------------------------------------------------------------------------------------------
import block_4_tv_helpers as b4  # noqa: E402
import block_4_tv_helpers_with_plot as b4p  # noqa: E402

for N in (5, 16):
    x = rng.standard_normal(N * N)
    x[rng.integers(0, N * N, 3)] = 0.0
    xf = np.round(x, 1)  # plateaus -> exercises the |g| <= eps mask
    px, py = rng.standard_normal((N, N)), rng.standard_normal((N, N))
    gx, gy = b4._grad_forward_2d_from_vec(x, N)
    out[f"b4_N{N}_x"], out[f"b4_N{N}_xf"], out[f"b4_N{N}_px"], out[f"b4_N{N}_py"] = x, xf, px, py
    out[f"b4_N{N}_gx"], out[f"b4_N{N}_gy"] = gx, gy
    out[f"b4_N{N}_div"] = b4._div_backward_2d_to_vec(px, py, N)
    out[f"b4_N{N}_kt"] = b4.kt_subgrad_isotropic_tv_from_x(x, N)
    out[f"b4_N{N}_ktf"] = b4.kt_subgrad_isotropic_tv_from_x(xf, N)
    out[f"b4_N{N}_edge"] = b4p.edge_map_from_vector(x, N)
    out[f"b4_N{N}_edge_raw"] = b4p.edge_map_from_vector(x, N, normalize=False)

# ---- block_3 ------------------------------------------------------------------------------------------------
import block_3_graph_and_precisions as b3  # noqa: E402

V3, m3, n3 = 5, 7, 12
A3 = [rng.standard_normal((m3, n3)).astype(np.float32) for _ in range(V3)]
A3[1][:, 4] = 0.0  # exercises the 1e-12 floor
out["b3_A"] = np.stack(A3)
for mode in ("arithmetic", "harmonic"):
    Wi, Q = b3.make_precisions(A3, q_mode=mode)
    out[f"b3_W_{mode}"] = np.stack(Wi)
    out[f"b3_Q_{mode}"] = np.stack([np.stack([Q(i, j) if i != j else np.zeros(n3) for j in range(V3)]) for i in range(V3)])
    qc = b3._precompute_q_cache(V3, Q)
    for strat in ("knn", "mst", "chain"):
        out[f"b3_keep_{mode}_{strat}"] = b3._build_all_pixel_masks(qc, V3, n3, strategy=strat, k=2, seed=123)

# ---- block_2 angle split (integers) through the reference function -------------------------------------------
import block_1_env_and_imports  # noqa: E402,F401  (block_2 imports ConstIm/randIm from it: absent there -> inject)
import Gen_Sino_Partitioned as gs  # noqa: E402

sys.modules["block_1_env_and_imports"].ConstIm = gs.ConstIm
sys.modules["block_1_env_and_imports"].randIm = gs.randIm
import block_2_load_odl_data as b2  # noqa: E402

split_cases = [(128, 4, 180), (512, 16, 360), (1024, 32, 720), (2048, 64, 720), (64, 5, None), (33, 7, 100)]
rows = []
for N, V, M in split_cases:
    ODL_CALLS.clear()
    space, rts, agg = b2._build_parallel_beam_operators(N, V, angles_total=M)
    per = [rt.geom.angles.m for rt in rts]
    rows.append([N, V, -1 if M is None else M, agg.geom.angles.m] + per + [-1] * (64 - len(per)))
    assert all(rt.geom.angles.a == 0.0 and abs(rt.geom.angles.b - np.pi) < 1e-15 for rt in rts)  # App. B-1
    assert all(rt.geom.det.m == N and rt.geom.det.a == -1.0 and rt.geom.det.b == 1.0 for rt in rts)
out["b2_split"] = np.array(rows, dtype=np.int64)

# ---- phantoms, psnr ----------------------------------------------------------------------------------------
for N in (32, 64):
    out[f"ConstIm_{N}"] = gs.ConstIm(N)
    np.random.seed(7 + N)
    out[f"randIm_{N}_seed{7 + N}"] = gs.randIm(N)
import test_final_integration as tfi  # noqa: E402

pa, pb = rng.random((9, 9)), rng.random((9, 9))
out["psnr_a"], out["psnr_b"] = pa, pb
out["psnr_val"] = np.array([tfi.psnr(pa, pb), tfi.psnr(pa, pb, data_range=2.5)])

# ---- block_6_ver2 outer loop with a stubbed block_5 ------------------------------------------------------------
import networkx as nx  # noqa: E402


class DenseOp:
    """Adapter giving the oracle's x-update a dense matrix (what the reference holds in A_dense_list)."""

    def __init__(self, A, N):
        self.A, self.N, self.D = A, N, None

    def forward(self, v):
        return self.A @ v

    def adjoint(self, q):
        return self.A.T @ q


LOOP = dict(S=1, C=6, mu=1.5)
_state = {}


class _Var:
    value = None


class _Prob:
    def __init__(self, key, Ai, bi, rho, vs, N, lam, Qs):
        self.key, self.args = key, (Ai, bi, rho, vs, N, lam, Qs)
        self.value, self.status = None, None
        self.solver_stats = types.SimpleNamespace(num_iters=LOOP["C"])

    def solve(self, **kw):
        # every call -- the first and the retries of the acceptance loop (block_6_ver2:115-176) -- is one more solve
        _solves.append((_calls["k"] - 1) // _calls["V"], self.key, float(kw["eps"]))
        assert kw["max_iters"] == 200 and kw["warm_start"] is True
        Ai, bi, rho, vs, N, lam, Qs = self.args
        n = N * N
        op = DenseOp(np.asarray(Ai, dtype=np.float64), N)
        st = _state.setdefault(self.key, dict(x=np.zeros(n), d=np.zeros(2 * n), w=np.zeros(2 * n)))
        cons = np.zeros(n)
        Dv = np.zeros(n)
        for v, q in zip(vs, Qs):
            cons += rho * q * v
            Dv += q
        x = st["x"].copy()
        O.np_x_update(op, 1.0, op.adjoint(bi) + cons, rho * Dv, LOOP["mu"], lam, LOOP["S"], LOOP["C"], x, st["d"], st["w"])
        st["x"] = x
        self.var.value = x
        res = op.forward(x) - bi
        pen = sum(float(np.sum(q * (x - v) ** 2)) for v, q in zip(vs, Qs))
        self.value = 0.5 * float(res @ res) + lam * O.tv_canonical(x, N) + 0.5 * rho * pen
        self.status = "optimal"
        return self.value


_calls = {"k": 0}


class _Solves(list):
    def append(self, k, node, eps):
        super().append((k, node, eps))


_solves = _Solves()


def _build_node_problem(Ai, bi, rho, neighbor_terms, N, lam_tv, Qij_terms):
    key = _calls["k"] % _calls["V"]
    _calls["k"] += 1
    var = _Var()
    prob = _Prob(key, Ai, bi, rho, neighbor_terms, N, lam_tv, Qij_terms)
    prob.var = var
    return var, prob


stub("block_5_node_problem", build_node_problem=_build_node_problem)
import block_6_admm_loop_ver2 as b6  # noqa: E402

import contextlib  # noqa: E402
import io  # noqa: E402

N6, M6 = 12, 24
for tag, G, iters in (("ring4", nx.cycle_graph(4), 16),
                      ("irr5", nx.Graph([(0, 3), (3, 1), (1, 4), (4, 0), (2, 3), (2, 1)]), 14)):
    V6 = G.number_of_nodes()
    thetas = O.node_angles(M6, V6)
    ops = [O.JosephOperator(N6, t) for t in thetas]
    dense = [op.dense().astype(np.float32) for op in ops]            # float32 like the reference's A
    img = O.shepp_logan(N6)
    sinos = [(dense[i].astype(np.float64) @ img.reshape(-1) + 0.01 * np.random.default_rng(50 + i).standard_normal(dense[i].shape[0]))
             .reshape(len(thetas[i]), N6) for i in range(V6)]
    Wi, Q = b3.make_precisions(dense, q_mode="arithmetic")
    _state.clear()
    del _solves[:]
    _calls.update(k=0, V=V6)
    with contextlib.redirect_stdout(io.StringIO()):
        cwd = os.getcwd()
        os.chdir("/tmp")   # the reference writes admm_internal_params.txt into the cwd
        try:
            x, h = b6.decentralized_admm(dense, sinos, G, Wi, Q, N6, lam_tv=0.02, rho=2.0, max_iters=iters,
                                         eps_pri=1e-9, eps_dual=1e-9, verbose=False, phantom_true=img)
        finally:
            os.chdir(cwd)
    out[f"b6_{tag}_edges"] = np.array(list(G.edges()), dtype=np.int64)
    out[f"b6_{tag}_dense"] = np.stack([np.pad(d, ((0, max(dd.shape[0] for dd in dense) - d.shape[0]), (0, 0))) for d in dense])
    out[f"b6_{tag}_rows"] = np.array([d.shape[0] for d in dense])
    out[f"b6_{tag}_sino"] = np.concatenate([s.reshape(-1) for s in sinos])
    out[f"b6_{tag}_x"] = np.stack(x)
    out[f"b6_{tag}_solve_eps"] = np.array(_solves, dtype=np.float64)
    per = np.zeros((iters, V6), dtype=np.int64)
    for k6, node6, _ in _solves:
        per[int(k6), int(node6)] += 1
    print(tag, "solves per (iteration, node):", per.tolist())
    for key in ("primal", "dual", "pri_per_node", "dual_per_node", "obj_per_node", "obj_total", "mse_sino_per_node",
                "mse_sino_total", "img_mse_per_node", "img_mse_total", "g_norm_history", "eps_used_history",
                "eps_target_history"):
        out[f"b6_{tag}_{key}"] = np.array(h[key], dtype=np.float64)
out["b6_params"] = np.array([N6, M6, LOOP["S"], LOOP["C"], LOOP["mu"], 0.02, 2.0])

np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
print("wrote", os.path.join(HERE, "reference_golden.npz"), len(out), "arrays,",
      os.path.getsize(os.path.join(HERE, "reference_golden.npz")), "bytes")

# ---- call-compatibility facts: how the reference's drivers call the boundary functions -------------------------
# (keyword names passed, history keys indexed) extracted from the reference sources with `ast`; the CPU tests check
# that the drop-in modules accept exactly these calls.
import ast  # noqa: E402
import json  # noqa: E402

BOUNDARY = ("load_odl_data", "build_pixel_connected_Q_provider", "decentralized_admm", "build_node_problem",
            "make_precisions", "generate_sinogram", "kt_subgrad_isotropic_tv_from_x", "edge_map_from_vector")
facts = {"calls": [], "history_keys": [], "data_keys": []}
for fn in sorted(os.listdir(REF)):
    if not fn.endswith(".py"):
        continue
    try:
        tree = ast.parse(open(os.path.join(REF, fn)).read())
    except SyntaxError:
        continue
    for node in ast.walk(tree):
        if isinstance(node, ast.Call):
            name = node.func.id if isinstance(node.func, ast.Name) else getattr(node.func, "attr", None)
            if name in BOUNDARY:
                facts["calls"].append({"file": fn, "line": node.lineno, "func": name, "n_positional": len(node.args),
                                       "keywords": sorted(k.arg for k in node.keywords if k.arg)})
        if isinstance(node, ast.Subscript) and isinstance(node.value, ast.Name) and isinstance(node.slice, ast.Constant) \
                and isinstance(node.slice.value, str):
            if node.value.id in ("hist", "history"):
                facts["history_keys"].append(node.slice.value)
            if node.value.id == "data":
                facts["data_keys"].append(node.slice.value)
facts["history_keys"] = sorted(set(facts["history_keys"]))
facts["data_keys"] = sorted(set(facts["data_keys"]))
json.dump(facts, open(os.path.join(HERE, "reference_call_facts.json"), "w"), indent=1)
print("wrote reference_call_facts.json:", len(facts["calls"]), "calls,", facts["history_keys"], facts["data_keys"])
