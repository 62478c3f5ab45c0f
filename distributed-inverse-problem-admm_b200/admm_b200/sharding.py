"""Node sharding of the consensus graph over the GPUs of one box -- pure host logic, bit-exact.

Units = graph nodes.  gpu(i) = (i*G)//V (contiguous blocks).  Edge (i, j), i < j, is *local* to a rank that owns
both ends and *cut* otherwise; a cut edge lives on both owners: each keeps a replica of z_ij and its own
y_ij,end, receives the peer's a = x + y once per iteration (NCCL send/recv) and computes the identical z' (the
midpoint / W-weighted fusions are symmetric in (i, j)).  The owner of the min end contributes the edge's dual
residual.  The only collective is the all-reduce of the residual row.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .geometry import graph_csr, node_to_gpu


@dataclass
class LocalEdge:
    e: int            # global edge id (G.edges() order)
    gi: int           # global node ids, gi < gj
    gj: int
    i_local: bool
    j_local: bool
    owns_dual: bool   # this rank adds s2 / dual_node for the edge
    slot: int = 0     # index into this rank's edge arrays
    peer: int = -1    # owner of the remote end (cut edges)
    xslot: int = -1   # index into the exchange buffers of `peer`


@dataclass
class ShardPlan:
    world: int
    rank: int
    V: int
    node_rank: list
    local_nodes: list                 # global ids, ascending
    g2l: dict                         # global node id -> local index
    edges: np.ndarray                 # [E, 2] global
    nbr_ptr: np.ndarray
    nbr_idx: np.ndarray
    nbr_edge: np.ndarray
    nbr_end: np.ndarray
    local_edges: list = field(default_factory=list)
    eslot: dict = field(default_factory=dict)       # global edge id -> local slot
    peers: list = field(default_factory=list)       # ascending peer ranks this rank exchanges with
    exch: dict = field(default_factory=dict)        # peer -> [global edge ids] ascending (same list on both sides)

    @property
    def n_cut(self):
        return sum(len(v) for v in self.exch.values())


def build_shard_plan(G, world: int, rank: int) -> ShardPlan:
    edges, ptr, idx, ned, nend = graph_csr(G)
    V = G.number_of_nodes()
    nr = node_to_gpu(V, world)
    local_nodes = [i for i in range(V) if nr[i] == rank]
    sp = ShardPlan(world, rank, V, nr, local_nodes, {g: l for l, g in enumerate(local_nodes)}, edges, ptr, idx,
                   ned, nend)
    exch = {}
    for e, (i, j) in enumerate(edges):
        i, j = int(i), int(j)
        il, jl = nr[i] == rank, nr[j] == rank
        if not (il or jl):
            continue
        le = LocalEdge(e, i, j, il, jl, owns_dual=il, slot=len(sp.local_edges))
        if il != jl:
            le.peer = nr[j] if il else nr[i]
            exch.setdefault(le.peer, []).append(e)
        sp.eslot[e] = le.slot
        sp.local_edges.append(le)
    sp.peers = sorted(exch)
    sp.exch = {p: sorted(exch[p]) for p in sp.peers}
    for p in sp.peers:
        for k, e in enumerate(sp.exch[p]):
            sp.local_edges[sp.eslot[e]].xslot = k
    return sp


def post_exchange(dist, sp: ShardPlan, send: dict, recv: dict, group=None):
    """Post one neighbour exchange and return the requests: for every peer p, send[p] ([n_cut_p, n], this rank's
    a = x + y of the cut edges shared with p, in `sp.exch[p]` order) goes to p and p's matching buffer lands in
    recv[p].  Grouped P2P (ncclSend/ncclRecv inside one group on NCCL; works unchanged on gloo with CPU tensors)."""
    if not sp.peers:
        return []
    ops = []
    for p in sp.peers:
        ops.append(dist.P2POp(dist.isend, send[p], p, group=group))
        ops.append(dist.P2POp(dist.irecv, recv[p], p, group=group))
    return dist.batch_isend_irecv(ops)


def exchange(dist, sp: ShardPlan, send: dict, recv: dict, group=None):
    """Blocking form of post_exchange."""
    for req in post_exchange(dist, sp, send, recv, group):
        req.wait()


def cut_statistics(G, world: int) -> dict:
    """Edges cut by the contiguous map and the per-rank exchange volume in units of n floats."""
    edges = graph_csr(G)[0]
    nr = node_to_gpu(G.number_of_nodes(), world)
    cut = [(int(i), int(j)) for i, j in edges if nr[int(i)] != nr[int(j)]]
    per_rank = [0] * world
    for i, j in cut:
        per_rank[nr[i]] += 1
        per_rank[nr[j]] += 1
    return {"edges": len(edges), "cut": len(cut), "per_rank_ends": per_rank}
