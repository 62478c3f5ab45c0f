"""torchrun worker: the NCCL-sharded solve must reproduce the single-GPU solve (same kernels, same per-node
arithmetic; only the order of the fp64 residual sums changes)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "distributed-inverse-problem-admm_b200")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    from admm_b200 import RayTransformCUDA, make_graph, node_angles, shepp_logan
    from admm_b200.sharding import cut_statistics
    from block_6_admm_loop_ver2 import decentralized_admm
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for (N, M, V, graph, wq) in ((64, 96, 8, "er", False), (48, 60, 5, "ring", True)):
        thetas = node_angles(M, V)
        img = shepp_logan(N)
        ops = [RayTransformCUDA(N, t, device=local) for t in thetas]
        sinos = [np.asarray(op(op.domain.element(img)).asarray()) +
                 0.01 * np.random.default_rng(1234 + i).standard_normal((len(thetas[i]), N)).astype(np.float32)
                 for i, op in enumerate(ops)]
        G = make_graph(graph, V, seed=0, p=0.4)
        Wl, Q = None, None
        if wq:
            Wl = [np.maximum(op.colnorm2(), 1e-12) for op in ops]
            Q = lambda i, j: 0.5 * (Wl[i] + Wl[j])  # noqa: E731
        kw = dict(lam_tv=0.02, rho=2.0, max_iters=25, eps_pri=0.0, eps_dual=0.0, verbose=False, phantom_true=img,
                  cg_iters=6, weighted_z=wq)
        xs, hs = decentralized_admm(ops, sinos, G, Wl, Q, N, exchange="p2p", **kw)       # peer memory over NVLink
        xn, hn = decentralized_admm(ops, sinos, G, Wl, Q, N, exchange="nccl", **kw)      # grouped send/recv
        xp, hp = decentralized_admm(ops, sinos, G, Wl, Q, N, exchange="push", **kw)      # producers store into the peers
        x2, h2 = decentralized_admm(ops, sinos, G, Wl, Q, N, exchange="nccl", exchange_phases=2, **kw)
        xc, hc = decentralized_admm(ops, sinos, G, Wl, Q, N, partition="contiguous", **kw)   # default map: balanced min-cut
        assert all(np.array_equal(a, b) for a, b in zip(xs, xc))          # the map moves nodes, not arithmetic
        assert np.allclose(hs["primal"], hc["primal"], rtol=1e-10)          # (the all-reduce sums ranks' shares in another order)
        assert all(np.array_equal(a, b) for a, b in zip(xs, xp)) and hs["primal"] == hp["primal"]
        assert all(np.array_equal(a, b) for a, b in zip(xs, x2)) and hs["primal"] == h2["primal"]
        x1, h1 = decentralized_admm(ops, sinos, G, Wl, Q, N, distributed=False, **kw)
        assert all(np.array_equal(a, b) for a, b in zip(xs, xn)) and hs["primal"] == hn["primal"]
        cut = cut_statistics(G, world)["cut"]
        same_x = all(np.array_equal(a, b) for a, b in zip(xs, x1))
        tr = max(np.max(np.abs(np.array(hs[k]) - np.array(h1[k])) / np.maximum(np.abs(np.array(h1[k])), 1e-30))
                 for k in ("primal", "dual", "pri_per_node", "dual_per_node", "mse_sino_per_node", "obj_per_node"))
        good = same_x and tr < 1e-10 and (cut > 0 or world == 1)
        ok = ok and good
        if rank == 0:
            print(f"world {world} N {N} V {V} {graph}: cut edges {cut}, x bit-identical {same_x}, max trace rel diff {tr:.2e}")
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("MULTI_GPU_OK")


if __name__ == "__main__":
    main()
