"""PDHG consensus variant on the B200 kernels -- drop-in for the computation of the reference's ADMM_Tomo_Only.py
(:89-148).  The reference file is a script (pickled operators in, matplotlib out); its loop is the
function `pdhg_consensus` here, with the script's constants as defaults and its result lists under the same names.

Per outer iteration k (:89):
  * lambda_tv = lambda_penalty * exp(alpha_tv * k)                                                        (:95)
  * x_a = error-weighted convex combination of the node iterates, eta_i = |A_i(:,p)|_2 / (|x_i - phantom| + 1e-8),
    normalised per pixel (ground-truth dependent, as in the reference)                                     (:100-118)
  * every node: 5 steps of odl.solvers.pdhg from a zero dual for
        gamma |x - x_a|^2 + lambda_tv (|A_i x - b_i|^2 + |grad x|_{2,1}),  tau = sigma = 1 / |(A_i, grad)|  (:121-133)
  * the aggregate problem sum_i |A_i x - b_i|^2 + lambda_agg |grad x|_{2,1}: 15 warm-started steps          (:142-148)
  * metrics: image MSE (mean), sinogram error norm, per node and aggregate                                 (:134-139, :152-159)

All nodes advance together: one batched K1 / K2 launch plus three small kernels per PDHG step (csrc/pdhg.cu), state
resident on the device.  Conventions of the ODL pieces (Gradient, weighted adjoint, prox operators) are stated in
csrc/pdhg.cu; `power_method_opnorm` starts from a fixed vector instead of a random one and is evaluated once per operator
(the reference re-estimates the same norm every outer iteration).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from admm_b200 import _native as nat
from admm_b200.operators import Plan


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _PdhgProblem:
    """Device state of `nodes` simultaneous PDHG problems over one plan."""

    def __init__(self, torch, plan, sinos, adj, dev):
        self.torch, self.plan, self.dev = torch, plan, dev
        self.V, self.n, self.N = plan.V, plan.n, plan.N
        f32 = dict(dtype=torch.float32, device=dev)
        self.b = torch.from_numpy(np.concatenate([np.asarray(s, dtype=np.float32).reshape(-1) for s in sinos])).to(dev)
        self.x = torch.zeros(self.V, self.n, **f32)
        self.xbar = torch.zeros(self.V, self.n, **f32)
        self.y1 = torch.zeros_like(self.b)
        self.q = torch.zeros_like(self.b)
        self.y2 = torch.zeros(self.V, 2, self.n, **f32)
        self.back = torch.zeros(self.V, self.n, **f32)
        self.adj = torch.tensor(adj, **f32)
        self.sums = torch.zeros(self.V, 2, dtype=torch.float64, device=dev)
        self.step = None      # [V] tau = sigma

    def _st(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def opnorm(self, iters):
        """power_method_opnorm of L = (A_i, grad) per node: x <- L* L x / |x|, estimate sqrt(|L* L x| / |x|)."""
        torch, L, h = self.torch, nat.lib(), self.plan.handle
        k = torch.arange(self.n, device=self.dev, dtype=torch.float64)
        v = (1.0 + 0.5 * torch.cos(0.37 * k)).to(torch.float32).repeat(self.V, 1).contiguous()
        out = torch.empty_like(v)
        est = np.zeros(self.V)
        for _ in range(iters):
            nat.check(L.admm_pdhg_sums(h, _ptr(v), self.n, None, None, None, _ptr(self.sums), 0, self.V, self._st()),
                      "admm_pdhg_sums")
            v = (v / torch.sqrt(self.sums[:, 0]).to(torch.float32)[:, None]).contiguous()
            self.plan.forward(v, self.q)
            self.plan.adjoint(self.q, self.back)
            nat.check(L.admm_pdhg_normal(h, _ptr(v), self.n, _ptr(self.back), _ptr(self.adj), _ptr(out), 0, self.V,
                                         self._st()), "admm_pdhg_normal")
            nat.check(L.admm_pdhg_sums(h, _ptr(out), self.n, None, None, None, _ptr(self.sums), 0, self.V, self._st()),
                      "admm_pdhg_sums")
            est = np.sqrt(np.sqrt(self.sums[:, 0].cpu().numpy()))
            v, out = out, v
        self.step = self.torch.tensor(1.0 / est, dtype=self.torch.float32, device=self.dev)
        return est

    def steps(self, niter, gamma, pull, lam_d, lam_t, theta=1.0, cold_dual=False):
        L, h = nat.lib(), self.plan.handle
        if cold_dual:
            self.y1.zero_()
            self.y2.zero_()
        self.xbar.copy_(self.x)
        for _ in range(niter):
            self.plan.forward(self.xbar, self.q)
            nat.check(L.admm_pdhg_dual(h, _ptr(self.xbar), self.n, _ptr(self.y1), _ptr(self.y2), _ptr(self.q), _ptr(self.b),
                                       _ptr(self.step), lam_d, lam_t, 0, self.V, self._st()), "admm_pdhg_dual")
            self.plan.adjoint(self.y1, self.back)
            nat.check(L.admm_pdhg_primal(h, _ptr(self.x), _ptr(self.xbar), self.n, _ptr(self.back), _ptr(self.y2),
                                         _ptr(pull), _ptr(self.step), _ptr(self.adj), gamma, theta, 0, self.V, self._st()),
                      "admm_pdhg_primal")

    def metrics(self, phantom):
        """(sum (x - phantom)^2, sum (A x - b)^2) per node, fp64 on the device."""
        self.plan.forward(self.x, self.q)
        nat.check(nat.lib().admm_pdhg_sums(self.plan.handle, _ptr(self.x), self.n, _ptr(phantom), _ptr(self.q), _ptr(self.b),
                                           _ptr(self.sums), 0, self.V, self._st()), "admm_pdhg_sums")
        return self.sums.cpu().numpy().copy()


def pdhg_consensus(ray_transforms, sinograms, phantom, niter=100, lambda_penalty=0.005, alpha_tv=0.0, lambda_agg=0.005,
                   gamma=2.0, node_niter=5, agg_niter=15, opnorm_iters=30, device=0, verbose=False):
    """The loop of ADMM_Tomo_Only.py:89-159 on the device.  `ray_transforms`: the node operators (RayTransformCUDA, as
    block_2's `ray_transforms`), `sinograms`: their (noisy) data, `phantom`: the (N, N) ground truth the weights use.
    Returns a dict with the script's result names: x_vars, x_agg, mse_lists, mse_sino_lists, mse_agg_list,
    mse_agg_sino_list (plus the operator-norm estimates)."""
    import torch
    nat.require_cuda()
    ops = list(ray_transforms)
    V, N = len(ops), ops[0].N
    n = N * N
    dev = torch.device("cuda", device)
    torch.cuda.set_device(dev)
    hx2 = ops[0].domain.cell_volume
    plan = Plan(N, [o.theta for o in ops], ops[0].D, ops[0].det_w, device, impl=ops[0].impl)
    agg = Plan(N, [np.concatenate([o.theta for o in ops])], ops[0].D, ops[0].det_w, device, impl=ops[0].impl)
    # aggregate range: the reference's agg operator covers [0, pi) with all angles (block_2_load_odl_data.py:58-63)
    agg_cell = math.pi / sum(o.nang for o in ops) * (ops[0].det_w / ops[0].D)
    nodes = _PdhgProblem(torch, plan, sinograms, [o.range.cell_volume / hx2 for o in ops], dev)
    glob = _PdhgProblem(torch, agg, sinograms, [agg_cell / hx2], dev)
    ph = torch.from_numpy(np.asarray(phantom, dtype=np.float32).reshape(-1)).to(dev)
    cn = torch.empty(V, n, dtype=torch.float32, device=dev)
    plan.colnorm2(cn)
    cn.sqrt_()                                                            # :55 np.linalg.norm(A_i_dense, axis=0)
    xa = torch.empty(n, dtype=torch.float32, device=dev)
    out = {"mse_lists": [[] for _ in range(V)], "mse_sino_lists": [[] for _ in range(V)], "mse_agg_list": [],
           "mse_agg_sino_list": []}
    out["op_norms"] = [float(v) for v in nodes.opnorm(opnorm_iters)]
    out["op_norm_agg"] = float(glob.opnorm(opnorm_iters)[0])
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)   # noqa: E731
    for k in range(niter):
        lam = float(lambda_penalty * math.exp(alpha_tv * k))
        nat.check(nat.lib().admm_pdhg_combine(plan.handle, _ptr(nodes.x), n, _ptr(cn), _ptr(ph), _ptr(xa), V, st()),
                  "admm_pdhg_combine")
        nodes.steps(node_niter, float(gamma), xa, lam, lam, cold_dual=True)
        m = nodes.metrics(ph)
        for i in range(V):
            out["mse_lists"][i].append(float(m[i, 0] / n))
            out["mse_sino_lists"][i].append(float(math.sqrt(m[i, 1])))
        glob.steps(agg_niter, 0.0, None, 1.0, float(lambda_agg))
        g = glob.metrics(ph)
        out["mse_agg_list"].append(float(g[0, 0] / n))
        out["mse_agg_sino_list"].append(float(math.sqrt(g[0, 1])))
        if verbose and (k % 20 == 0 or k == niter - 1):
            print(f"Iteration {k + 1:03d}  image MSEs = " + ", ".join(f"{v[-1]:.4f}" for v in out["mse_lists"]) +
                  f"  agg = {out['mse_agg_list'][-1]:.4f}")
    out["x_vars"] = [nodes.x[i].cpu().numpy().reshape(N, N) for i in range(V)]
    out["x_agg"] = glob.x[0].cpu().numpy().reshape(N, N)
    plan.close()
    agg.close()
    return out
