"""Test-only: a rank of the SHARDED decentralized ADMM with the oracle's CPU kernels standing in for the CUDA ones,
driven by the product's own sharding logic (admm_b200.sharding: ShardPlan, exchange schedule, ownership flags).
Lets the multi-rank path (neighbour exchange + the residual all-reduce) be checked on CPU with gloo."""
import math

import numpy as np


def run_rank(rank, world, port, cfg, ret):
    import torch
    import torch.distributed as dist
    from admm_b200.sharding import build_shard_plan, partition_nodes, phase_bounds, post_exchange
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    try:
        N, M, V, iters, rho, lam = cfg["N"], cfg["M"], cfg["V"], cfg["iters"], cfg["rho"], cfg["lam"]
        n = N * N
        G = O.make_graph(cfg["graph"], V, seed=0, p=0.4, degree=3)
        phases = cfg.get("phases", 1)
        sp = build_shard_plan(G, world, rank, phases, partition_nodes(G, world, cfg.get("partition", "contiguous")))
        thetas = O.node_angles(M, V)
        img = O.shepp_logan(N)
        ops = {g: O.JosephOperator(N, thetas[g]) for g in sp.local_nodes}
        b = {g: ops[g].forward(img) + 0.01 * np.random.default_rng(1234 + g).standard_normal(ops[g].shape[0])
             for g in sp.local_nodes}
        atb = {g: ops[g].adjoint(b[g]) for g in sp.local_nodes}
        x = {g: np.zeros(n) for g in sp.local_nodes}
        d = {g: np.zeros(2 * n) for g in sp.local_nodes}
        w = {g: np.zeros(2 * n) for g in sp.local_nodes}
        E = len(sp.local_edges)
        z = [np.zeros(n) for _ in range(E)]
        y = [[np.zeros(n), np.zeros(n)] for _ in range(E)]
        send = {p: torch.zeros(len(sp.exch[p]), n, dtype=torch.float64) for p in sp.peers}
        recv = {p: torch.zeros(len(sp.exch[p]), n, dtype=torch.float64) for p in sp.peers}
        pri_hist, dual_hist = [], []
        for _ in range(iters):
            reqs = []
            bounds = phase_bounds(len(sp.local_nodes), phases)
            for li, g in enumerate(sp.local_nodes):
                cons = np.zeros(n)
                for kk in range(sp.nbr_ptr[g], sp.nbr_ptr[g + 1]):
                    s, end = sp.eslot[int(sp.nbr_edge[kk])], int(sp.nbr_end[kk])
                    cons += rho * (z[s] - y[s][end])
                deg = int(sp.nbr_ptr[g + 1] - sp.nbr_ptr[g])
                O.x_update(ops[g], 1.0, atb[g] + cons, rho * deg, rho, lam, 1, 6, x[g], d[g], w[g])
                if li + 1 in bounds[1:]:                    # a phase just finished: pack its cut-edge ends and post them
                    for ph in [k for k in range(phases) if bounds[k + 1] == li + 1]:
                        for le in sp.local_edges:           # a = x + y of this rank's end of every cut edge of the phase
                            if le.peer >= 0 and le.sphase == ph:
                                gg, end = (le.gi, 0) if le.i_local else (le.gj, 1)
                                send[le.peer][le.sslot] = torch.from_numpy(x[gg] + y[le.slot][end])
                        reqs += post_exchange(dist, sp, send, recv, phase=ph if phases > 1 else None)
            for r in reqs:
                r.wait()
            row = np.zeros(2 + 2 * V)
            for le in sp.local_edges:
                s = le.slot
                ai = x[le.gi] + y[s][0] if le.i_local else recv[le.peer][le.rslot].numpy()
                aj = x[le.gj] + y[s][1] if le.j_local else recv[le.peer][le.rslot].numpy()
                zn = (ai + aj) / 2.0
                if le.i_local:
                    ri = x[le.gi] - zn
                    y[s][0] = y[s][0] + x[le.gi] - zn
                    row[0] += ri @ ri
                    row[2 + le.gi] += ri @ ri
                if le.j_local:
                    rj = x[le.gj] - zn
                    y[s][1] = y[s][1] + x[le.gj] - zn
                    row[0] += rj @ rj
                    row[2 + le.gj] += rj @ rj
                if le.owns_dual:
                    dz = zn - z[s]
                    row[1] += rho * rho * (dz @ dz)
                    row[2 + V + le.gi] += rho * rho * (dz @ dz)
                    row[2 + V + le.gj] += rho * rho * (dz @ dz)
                z[s] = zn
            t = torch.from_numpy(row)
            dist.all_reduce(t)                              # the only collective
            pri_hist.append(math.sqrt(row[0]))
            dual_hist.append(math.sqrt(row[1]))
        ret[rank] = dict(primal=pri_hist, dual=dual_hist, x={g: x[g].copy() for g in sp.local_nodes},
                         last_row=row.copy(), n_cut=sp.n_cut)
    finally:
        dist.destroy_process_group()
