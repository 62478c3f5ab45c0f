"""Tiny solve for compute-sanitizer runs (not a pytest file)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "distributed-inverse-problem-admm_b200")):
    sys.path.insert(0, p)
import numpy as np
from admm_b200 import RayTransformCUDA, node_angles, shepp_logan, make_graph
from block_6_admm_loop_ver2 import decentralized_admm
N, M, V = 48, 36, 3
thetas = node_angles(M, V, sys.argv[1] if len(sys.argv) > 1 else "contiguous")
img = shepp_logan(N)
ops = [RayTransformCUDA(N, t) for t in thetas]
sinos = [op(op.domain.element(img)).asarray() for op in ops]
x, h = decentralized_admm(ops, sinos, make_graph("ring", V), None, None, N, lam_tv=0.02, rho=2.0, max_iters=3, eps_pri=0, eps_dual=0,
                          verbose=False, phantom_true=img, cg_iters=3)
print("ok", h["primal"])
