#!/usr/bin/env python
"""Oracle-only study (CPU, fp64): how far is an inexact x-update (S TV-split sweeps x C CG iterations per solve, with or
without the reference's accept / tighten-and-retry rule, block_6_admm_loop_ver2.py:100-176) from a CONVERGED inner
solve of eq. (1), in the quantities north_star grades: per-iteration primal/dual residual traces, final x, PSNR.

    python tools/inner_schedule_study.py [--configs cfg1,cfg2s] [--iters 200] [--out profiles/r2_inner_schedule_study.json]

cfg1  = BASELINE configs[0] at full size (128^2, 180 angles, ring of 4);
cfg2s = BASELINE configs[1]'s shape (360 angles, 16 nodes, random 4-regular graph) at 128^2 -- the converged inner
        solve of the 512^2 problem is days of CPU time.
The "converged" reference runs REF_S = 50 sweeps x REF_C = 8 CG per solve with mu = 4 rho (the inner fixed point does
not depend on mu; 4 rho converges fastest here: per-sweep change 4e-5 |x| after 50 sweeps from a cold start, far less
when warm-started); `--deeper` re-runs it with 100 sweeps to show what is left.
Uses only oracle/ (test infrastructure); nothing here is on the product path.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

LAM, RHO, SIGMA = 0.02, 2.0, 0.005

CONFIGS = {
    "cfg1": dict(N=128, M=180, V=4, graph=("ring", {})),
    "cfg1s": dict(N=64, M=180, V=4, graph=("ring", {})),
    "cfg2s": dict(N=128, M=360, V=16, graph=("regular", dict(seed=0, degree=4))),
}

# name -> (S, C, acceptance, mu)
REF_S, REF_C, REF_MU = 50, 8, 4 * RHO
SCHEDULES = [
    ("S1C8", 1, 8, False, RHO),          # round-1 bench schedule
    ("S1C4+accept", 1, 4, True, RHO),
    ("S1C8+accept", 1, 8, True, RHO),
    ("S2C4", 2, 4, False, RHO),
    ("S3C3", 3, 3, False, RHO),
    ("S4C2", 4, 2, False, RHO),
    ("S3C8", 3, 8, False, RHO),
    ("S3C30", 3, 30, False, RHO),
    ("S8C4", 8, 4, False, RHO),
    ("S1C8 mu=4rho", 1, 8, False, 4 * RHO),
    ("S2C4 mu=4rho", 2, 4, False, 4 * RHO),
    ("S4C2 mu=4rho", 4, 2, False, 4 * RHO),
    ("S8C4 mu=4rho", 8, 4, False, 4 * RHO),
    # what decides the accuracy is the number of TV sweeps, not the CG count: cheap sweeps
    ("S1C1+accept", 1, 1, True, RHO),
    ("S1C2+accept", 1, 2, True, RHO),
    ("S2C2+accept", 2, 2, True, RHO),
    ("S2C1+accept", 2, 1, True, RHO),
    ("S4C1", 4, 1, False, RHO),
    ("S6C1", 6, 1, False, RHO),
    ("S6C2", 6, 2, False, RHO),
    ("S8C1", 8, 1, False, RHO),
    ("S12C2", 12, 2, False, RHO),
]


def problem(cfg):
    N, M, V = cfg["N"], cfg["M"], cfg["V"]
    thetas = O.node_angles(M, V)
    img = O.shepp_logan(N)
    ops = [O.JosephOperator(N, t) for t in thetas]
    sinos = [op.forward(img) + SIGMA * np.random.default_rng(1234 + i).standard_normal(op.shape[0])
             for i, op in enumerate(ops)]
    G = O.make_graph(cfg["graph"][0], V, **cfg["graph"][1])
    return ops, sinos, G, img


class Counting:
    """x-update wrapper that counts the inner work actually done."""

    def __init__(self):
        self.cg = 0
        self.sweeps = 0

    def __call__(self, op, prec, rhs0, rhoD, mu, lam, S, C, x, d, w):
        self.cg += S * C
        self.sweeps += S
        return O.x_update(op, prec, rhs0, rhoD, mu, lam, S, C, x, d, w)


def run(cfg, iters, S, C, accept, mu):
    ops, sinos, G, img = problem(cfg)
    cnt = Counting()
    t = time.perf_counter()
    x, h = O.decentralized_admm(ops, sinos, G, None, None, cfg["N"], lam_tv=LAM, rho=RHO, max_iters=iters,
                                eps_pri=0.0, eps_dual=0.0, phantom_true=img, tv_mu=mu, tv_sweeps=S, cg_iters=C,
                                uniform_q=1.0, x_update_fn=cnt, acceptance=accept)
    return x, h, cnt, time.perf_counter() - t, img


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="cfg1,cfg2s")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_inner_schedule_study.json"))
    ap.add_argument("--schedules", default="")
    ap.add_argument("--deeper", action="store_true", help="also run the reference with 2x the sweeps")
    args = ap.parse_args()
    want = [s for s in args.schedules.split(",") if s]
    out = {}
    if os.path.exists(args.out):
        out = json.load(open(args.out))
    for cname in args.configs.split(","):
        cfg = CONFIGS[cname]
        V, N = cfg["V"], cfg["N"]
        # the reference run is the expensive part: cached next to the output (not committed)
        cache = os.path.join(os.path.dirname(args.out), f"_study_ref_{cname}_{args.iters}.npz")
        if os.path.exists(cache):
            z = np.load(cache, allow_pickle=True)
            xr, hr, tr, img = list(z["x"]), z["h"].item(), float(z["t"]), z["img"]
            cr = Counting()
            cr.cg, cr.sweeps = REF_S * REF_C * V * args.iters, REF_S * V * args.iters
        else:
            xr, hr, cr, tr, img = run(cfg, args.iters, REF_S, REF_C, False, REF_MU)
            np.savez_compressed(cache, x=np.stack(xr), h=np.array(hr, dtype=object), t=tr, img=img)
        pr, dr = np.array(hr["primal"]), np.array(hr["dual"])
        res = out.setdefault(cname, {})
        res["_config"] = {k: (v if not isinstance(v, tuple) else list(v)) for k, v in cfg.items()}
        res["_config"].update(lam_tv=LAM, rho=RHO, sigma=SIGMA, iters=args.iters)
        res["converged"] = {"S": REF_S, "C": REF_C, "tv_mu": REF_MU, "cg_per_node_iter": cr.cg / (V * args.iters), "sweeps_per_node_iter": cr.sweeps / (V * args.iters),
                            "final_primal": float(pr[-1]), "final_dual": float(dr[-1]),
                            "psnr_node0": float(O.psnr(xr[0].reshape(N, N), img)), "wall_s": round(tr, 1),
                            "g_norm_last": [float(v) for v in hr["g_norm_history"][-1]],
                            "eps_target_last": float(hr["eps_target_history"][-1][0])}
        print(cname, "converged:", res["converged"], flush=True)
        scheds = list(SCHEDULES)
        if args.deeper:
            scheds.insert(0, ("reference x2 sweeps", 2 * REF_S, REF_C, False, REF_MU))
        for name, S, C, acc, mu in scheds:
            if want and name not in want:
                continue
            x, h, cnt, t, _ = run(cfg, args.iters, S, C, acc, mu)
            p, d = np.array(h["primal"]), np.array(h["dual"])
            tight = np.array(h["tighten_history"])
            r = {
                "S": S, "C": C, "acceptance": acc, "tv_mu": mu,
                "cg_per_node_iter": cnt.cg / (V * args.iters), "sweeps_per_node_iter": cnt.sweeps / (V * args.iters),
                "primal_trace_max_rel_err": float(np.max(np.abs(p - pr) / pr)),
                "dual_trace_max_rel_err": float(np.max(np.abs(d - dr) / dr)),
                "primal_trace_rel_err_at_last": float(abs(p[-1] - pr[-1]) / pr[-1]),
                "final_x_rel_l2_max": float(max(np.linalg.norm(a - b) / np.linalg.norm(b) for a, b in zip(x, xr))),
                "psnr_diff_db_max": float(max(abs(O.psnr(a.reshape(N, N), img) - O.psnr(b.reshape(N, N), img))
                                              for a, b in zip(x, xr))),
                "tighten_mean": float(tight.mean()), "accepted_first_try_frac": float((tight == 0).mean()),
                "g_norm_last_max": float(np.max(h["g_norm_history"][-1])), "wall_s": round(t, 1),
            }
            res[name] = r
            print(cname, name, r, flush=True)
            json.dump(out, open(args.out, "w"), indent=1)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
