"""GPU-vs-oracle trace error of the residual-carry modes ("Reductions").

Runs the cfg-1 shape for `iters` outer iterations with carry_residual in {False, 'iteration', 'always'} and prints the
relative trace error at a few checkpoints, so the growth pattern (roundoff floor vs accumulation) is visible.
Usage (GPU box):  python tools/carry_study.py [N] [iters] [default-only] [diag]
Any host:         python tools/carry_study.py N iters default-only oracle-only     (writes the cached fp64 oracle run; the
                  128 / 200 one is committed as tests/golden/cfg1_default_schedule_200.npz)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "distributed-inverse-problem-admm_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402


CARRIES = (False, "first_retry", "iteration", "always")
SCHEDS = ((1, 2, True), (1, 8, False))


def main():
    global CARRIES, SCHEDS
    if "diag" in sys.argv:
        CARRIES = (False,)
    if "default-only" in sys.argv:
        CARRIES, SCHEDS = (False, "first_retry", "iteration"), ((1, 2, True),)
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    from admm_b200 import RayTransformCUDA, node_angles
    from oracle import oracle as O
    M, V = 180, 4
    thetas = node_angles(M, V, "contiguous")
    img = O.shepp_logan(N)
    ops_o = [O.JosephOperator(N, t) for t in thetas]
    sinos = [(op.forward(img) + 0.005 * np.random.default_rng(1234 + i).standard_normal(op.shape[0]))
             .reshape(op.nang, op.D).astype(np.float32) for i, op in enumerate(ops_o)]
    if "oracle-only" not in sys.argv:
        from block_6_admm_loop_ver2 import decentralized_admm
        ops_g = [RayTransformCUDA(N, t) for t in thetas]
    G = O.make_graph("ring", V)
    kw = dict(lam_tv=0.02, rho=2.0, max_iters=iters, eps_pri=0.0, eps_dual=0.0, phantom_true=img)
    out = []
    for (S, C, acc) in SCHEDS:
        # the fp64 oracle run is the expensive part and needs no GPU: cached next to the profiles (not committed), so it can
        # be produced on any host (`oracle-only`) and travel to the GPU box with the snapshot
        cache = os.path.join(ROOT, "profiles", f"_study_ref_carry{N}_{iters}_S{S}C{C}{int(acc)}.npz")
        if os.path.exists(cache):
            zf = np.load(cache)
            xo, ho = list(zf["x"]), {"primal": zf["primal"], "dual": zf["dual"], "tighten_history": zf["tighten"]}
        else:
            xo, ho = O.decentralized_admm(ops_o, sinos, G, None, None, N, uniform_q=1.0, tv_sweeps=S, cg_iters=C,
                                          acceptance=acc, **kw)
            np.savez_compressed(cache, x=np.stack(xo), primal=np.array(ho["primal"]), dual=np.array(ho["dual"]),
                                tighten=np.array(ho["tighten_history"]))
        if "oracle-only" in sys.argv:
            continue
        po, do = np.array(ho["primal"]), np.array(ho["dual"])
        for carry in CARRIES:
            xg, hg = decentralized_admm(ops_g, sinos, G, None, None, N, verbose=False, tv_sweeps=S, cg_iters=C,
                                        acceptance=acc, carry_residual=carry, **kw)
            pg, dg = np.array(hg["primal"]), np.array(hg["dual"])
            ep, ed = np.abs(pg - po) / po, np.abs(dg - do) / do
            xe = max(np.linalg.norm(a - b) / np.linalg.norm(b) for a, b in zip(xg, xo))
            row = {"N": N, "S": S, "C": C, "accept": acc, "carry": carry, "primal_max": float(ep.max()),
                   "primal_argmax": int(ep.argmax()), "dual_max": float(ed.max()), "dual_argmax": int(ed.argmax()),
                   "x": float(xe), "primal_at": {k: float(ep[k - 1]) for k in (10, 50, 100, iters) if k <= iters},
                   "dual_at": {k: float(ed[k - 1]) for k in (10, 50, 100, iters) if k <= iters},
                   "primal_last": float(po[-1]), "dual_last": float(do[-1]),
                   "same_decisions": bool(np.array_equal(np.array(hg["tighten_history"]),
                                                         np.array(ho["tighten_history"])))}
            print(json.dumps(row), flush=True)
            out.append(row)
    return out


if __name__ == "__main__":
    main()
