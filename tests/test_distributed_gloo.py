"""world_size-2 (and 3) gloo runs of the sharded path on CPU: ShardPlan + grouped send/recv exchange + the residual
all-reduce reproduce the single-process loop."""
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,graph,V,phases,partition", [(2, "ring", 4, 1, "contiguous"), (2, "er", 7, 2, "mincut"),
                                                            (3, "regular", 6, 1, "mincut"), (3, "er", 8, 3, "contiguous")])
def test_sharded_equals_single_process(world, graph, V, phases, partition):
    """phases > 1: the exchange is posted block by block as the x-updates finish (separate send / recv row orders)."""
    import torch.multiprocessing as mp
    from dist_helpers import run_rank
    from oracle import oracle as O
    cfg = dict(N=16, M=36, V=V, iters=12, rho=2.0, lam=0.02, graph=graph, phases=phases, partition=partition)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(run_rank, args=(world, _free_port(), cfg, ret), nprocs=world, join=True)
    assert sorted(ret.keys()) == list(range(world))
    # single-process oracle on the same problem
    N, M = cfg["N"], cfg["M"]
    G = O.make_graph(graph, V, seed=0, p=0.4, degree=3)
    thetas = O.node_angles(M, V)
    ops = [O.JosephOperator(N, t) for t in thetas]
    img = O.shepp_logan(N)
    sinos = [ops[g].forward(img) + 0.01 * np.random.default_rng(1234 + g).standard_normal(ops[g].shape[0]) for g in range(V)]
    x, h = O.decentralized_admm(ops, sinos, G, None, None, N, lam_tv=cfg["lam"], rho=cfg["rho"], max_iters=cfg["iters"],
                                eps_pri=0, eps_dual=0, uniform_q=1.0, tv_sweeps=1, cg_iters=6, acceptance=False)   # run_rank: one solve per iteration
    assert sum(ret[r]["n_cut"] for r in range(world)) > 0            # the exchange path was exercised
    for r in range(world):
        assert np.allclose(ret[r]["primal"], h["primal"], rtol=1e-11)
        assert np.allclose(ret[r]["dual"], h["dual"], rtol=1e-11)
        for g, xg in ret[r]["x"].items():
            assert np.allclose(xg, x[g], rtol=1e-11, atol=1e-14)
    rows = [ret[r]["last_row"] for r in range(world)]
    assert all(np.array_equal(rows[0], rr) for rr in rows)            # every rank holds the same reduced row
    assert np.allclose(np.sqrt(rows[0][2:2 + V]), h["pri_per_node"][-1], rtol=1e-11)
    assert np.allclose(np.sqrt(rows[0][2 + V:]), h["dual_per_node"][-1], rtol=1e-11)
