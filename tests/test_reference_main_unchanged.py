"""`block_7_main_ver3.py` -- the reference's live driver -- runs UNCHANGED against the drop-in modules.

Executed with `runpy` straight from /root/reference (only present in the build container, so this test is skipped on
the GPU box) with `sys.path` pointing at the drop-ins and matplotlib stubbed.  This box has no GPU, so the CUDA layer
under the drop-ins is replaced, FOR THIS TEST ONLY, by the fp64 oracle: `RayTransformCUDA`'s three native calls and
the engine class `block_6_admm_loop_ver2` instantiates.  What is checked is the boundary: every import, call
signature, keyword, returned structure and history key the reference driver relies on, end to end
through `load_odl_data -> build_pixel_connected_Q_provider -> decentralized_admm -> np.save of every history`.
The CUDA path behind the same boundary is covered by tests/test_gpu_boundary.py::test_block7_*.
"""
import glob
import os
import runpy
import sys
import types

import numpy as np
import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "distributed-inverse-problem-admm_b200")

pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "block_7_main_ver3.py")),
                                reason="the reference checkout is only available in the build container")


class _Anything:
    def __getattr__(self, k):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter((_Anything(), _Anything()))


class OracleEngine:
    """Stands where `admm_b200.solver.ADMMEngine` stands (same constructor keywords, same methods), computing with the
    fp64 oracle.  Test infrastructure only."""

    def __init__(self, thetas, sinograms, G, N, D=None, det_w=2.0, lam_tv=0.01, rho=1.0, Q=None, Wi_list=None,
                 node_prec=None, tv_mu=None, tv_sweeps=1, cg_iters=8, phantom_true=None, weighted_z=False,
                 max_iters=200, acceptance=False, max_tighten=2, **_ignored):
        from oracle import oracle as O
        ops = [O.JosephOperator(N, t, D, det_w) for t in thetas]
        # a plumbing test: the numbers are checked elsewhere, so the 200 hard-coded outer iterations of the reference
        # driver run with the cheapest inner solve (2 CG iterations, no retries) to keep the CPU suite short
        cg_iters, acceptance = min(int(cg_iters), 2), False
        self.S, self.C, self.mu = tv_sweeps, cg_iters, (tv_mu if tv_mu is not None else rho)
        self.acceptance, self.max_tighten = acceptance, max_tighten
        self.xs = []
        uq = 1.0 if Q is None else (float(Q) if np.isscalar(Q) else None)
        _, self.h = O.decentralized_admm(ops, sinograms, G, Wi_list, Q, N, lam_tv=lam_tv, rho=rho, max_iters=max_iters,
                                         eps_pri=0.0, eps_dual=0.0, phantom_true=phantom_true, node_prec=node_prec,
                                         tv_mu=tv_mu, tv_sweeps=tv_sweeps, cg_iters=cg_iters, weighted_z=weighted_z,
                                         uniform_q=uq, stop=False, acceptance=acceptance, max_tighten=max_tighten,
                                         on_iteration=lambda k, x: self.xs.append([xi.copy() for xi in x]))
        self.k = 0

    def step(self):
        self.k += 1

    def residuals(self):
        return self.h["primal"][self.k - 1], self.h["dual"][self.k - 1]

    def history(self, iters=None):
        iters = self.k if iters is None else iters
        return {k: list(v[:iters]) for k, v in self.h.items()}

    def x_all(self, gather="all"):
        return [xi.astype(np.float32) for xi in self.xs[self.k - 1]]

    def close(self, sync=True):
        pass


def test_reference_block7_main_ver3_runs_unchanged(tmp_path, monkeypatch):
    from oracle import oracle as O
    import torch
    import admm_b200.operators as ops_mod
    import block_6_admm_loop_ver2 as b6
    assert b6.__file__.startswith(PKG)                     # the drop-in, not the reference's module

    def _ref(op):
        return O.JosephOperator(op.N, op.theta, op.D, op.det_w)

    monkeypatch.setattr(ops_mod.RayTransformCUDA, "_forward_np", lambda self, x: _ref(self).forward(x).astype(np.float32))
    monkeypatch.setattr(ops_mod.RayTransformCUDA, "_adjoint_np", lambda self, y: _ref(self).adjoint(y).astype(np.float32))
    monkeypatch.setattr(ops_mod.RayTransformCUDA, "colnorm2", lambda self: _ref(self).colnorm2())
    monkeypatch.setattr(b6, "ADMMEngine", OracleEngine)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    plt = types.ModuleType("matplotlib.pyplot")
    plt.__getattr__ = lambda name: (lambda *a, **k: _Anything())
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    mpl.use = lambda *a, **k: None
    monkeypatch.setitem(sys.modules, "matplotlib", mpl)
    monkeypatch.setitem(sys.modules, "matplotlib.pyplot", plt)
    monkeypatch.chdir(tmp_path)
    np.random.seed(0)                                      # the reference draws phantom and noise from the global RNG
    ns = runpy.run_path(os.path.join(REF, "block_7_main_ver3.py"), run_name="reference_block_7_main_ver3")
    # the reference's modules must not have been picked up in place of the drop-ins
    for name in ("block_2_load_odl_data", "block_3_graph_and_precisions", "block_6_admm_loop_ver2"):
        assert sys.modules[name].__file__.startswith(PKG), name
    ns["main"]()                                           # block_7_main_ver3.py:332-371, N=64, 5 nodes, knn k=2, 200 its

    out = glob.glob(str(tmp_path / "Recon_Out_ADMM_*" / "knn_k2"))
    assert len(out) == 1
    out = out[0]
    assert "Strategy: knn" in open(os.path.join(out, "run_parameters.txt")).read()          # :38-57
    npy = {os.path.basename(f) for f in glob.glob(os.path.join(out, "*.npy"))}
    want = {f"knn_k2_node_{i}.npy" for i in range(5)}                                       # save_recons :16-27
    want |= {f"knn_k2_{k}.npy" for k in ("primal_hist", "dual_hist", "pri_per_node", "dual_per_node", "obj_per_node",
                                         "obj_total", "sino_mse_per_node", "sino_mse_total", "img_mse_per_node",
                                         "img_mse_total")}                                  # :189-325
    assert want <= npy, want - npy
    assert np.load(os.path.join(out, "knn_k2_primal_hist.npy")).shape == (200,)
    assert np.load(os.path.join(out, "knn_k2_pri_per_node.npy")).shape == (200, 5)
    snaps = glob.glob(os.path.join(out, "snapshots", "iter_*_node_*.npy"))                  # block_6_ver2:269-281
    assert snaps and np.load(snaps[0]).shape == (64, 64)
    assert os.path.isfile(os.path.join(out, "snapshots", "admm_internal_params.txt"))       # block_6_ver2:293-306
    rec = [np.load(f) for f in sorted(glob.glob(os.path.join(out, "*node_0.npy")))]
    assert rec and rec[0].shape == (64, 64) and np.isfinite(rec[0]).all()
