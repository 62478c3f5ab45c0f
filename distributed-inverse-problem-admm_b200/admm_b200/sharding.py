"""Node sharding of the consensus graph over the GPUs of one box -- pure host logic, bit-exact.

Units = graph nodes.  gpu(i) = (i*G)//V (contiguous blocks) or a balanced min-cut map (partition_nodes).  Edge (i, j), i < j, is *local* to a rank that owns
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
    xslot: int = -1   # index into the exchange buffers of `peer` (phases == 1: same row both ways)
    sslot: int = -1   # row of this rank's a in send[peer]  (ordered by the phase of the LOCAL end's node)
    rslot: int = -1   # row of the peer's a in recv[peer]     (ordered by the phase of the REMOTE end's node)
    sphase: int = 0   # exchange phase in which this rank's end is sent
    owner: int = -1   # rank that computes the edge update in the single-owner exchange (local edges: this rank)


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
    phases: int = 1                                 # the exchange is posted in this many pieces per iteration
    node_phase: dict = field(default_factory=dict)  # global node id -> phase of its owner's x-update schedule
    send_rows: dict = field(default_factory=dict)   # peer -> [phases+1] row offsets into send[peer]
    recv_rows: dict = field(default_factory=dict)   # peer -> [phases+1] row offsets into recv[peer]

    @property
    def n_cut(self):
        return sum(len(v) for v in self.exch.values())


def _map_score(edges, nr, world):
    """(largest number of cut-edge ends on one rank, cut edges): the slowest rank's exchange volume comes first."""
    per = [0] * world
    cut = 0
    for i, j in edges:
        a, b = nr[int(i)], nr[int(j)]
        if a != b:
            per[a] += 1
            per[b] += 1
            cut += 1
    return (max(per) if per else 0, cut)


def partition_nodes(G, world: int, method: str = "auto", trials: int = 16, seed: int = 1234, weights=None) -> list:
    """Node -> GPU map with V//world or V//world+1 nodes per rank (the x-update work is per node, so the map stays
    balanced).  "contiguous": gpu(i) = (i*G)//V.  "mincut": deterministic multi-start pairwise-swap
    refinement (Kernighan-Lin moves on the cut size, vectorised: gain(i,j) = P[i,part j] + P[j,part i] - 2 A[i,j]
    with P[i,w] = #nbrs of i in w - #nbrs of i in its own part), best of `trials` starts by (max cut-edge ends on a
    rank, cut edges); never worse than contiguous, which is start 0.  "auto": mincut for V <= 256, else contiguous
    (the search is O(V^2) per move; slice-batched graphs are disjoint unions that the contiguous map already keeps
    whole).  Every rank calls this with the same arguments and gets the same list.

    `weights` (one integer per node, e.g. its number of projection angles -- the projector kernels' time is
    proportional to the angle rows a rank holds): the map is then also balanced by weight.  Definition, so that it can
    be re-derived: nodes are sorted by (-weight, id) and dealt to the ranks in snake order (0..G-1, G-1..0, ...):
    every rank gets V//G or V//G+1 nodes and the rank weights differ by at most max(weights) - min(weights); that is
    start 0.  Start t > 0 permutes, with np.random.RandomState(seed) drawn once for all starts in order, the nodes
    WITHIN each weight class before dealing.  Each start is refined by the swap moves above restricted to pairs of
    EQUAL weight (rank weights never change), ties in the gain broken by the smallest flat index i*V + j; the best
    start by (max cut-edge ends on a rank, cut edges, start index) wins; ranks are relabelled in order of their
    smallest node id."""
    V = G.number_of_nodes()
    cont = node_to_gpu(V, world)
    if method not in ("auto", "mincut", "contiguous"):
        raise ValueError(f"unknown partition method {method!r}")
    if world <= 1 or method == "contiguous" or (method == "auto" and V > 256) or V <= world:
        return cont
    edges = graph_csr(G)[0]
    if len(edges) == 0:
        return cont
    w = None
    if weights is not None:
        w = np.asarray([int(v) for v in weights], dtype=np.int64)
        if len(w) != V:
            raise ValueError("weights must hold one value per node")
        if np.all(w == w[0]):
            w = None                                  # uniform: the plain balanced min-cut
    key = (V, world, trials, seed, edges.tobytes(), None if w is None else w.tobytes())
    if key in _PARTITION_CACHE:                       # repeated solves on one graph (e.g. bench.py's two legs)
        return list(_PARTITION_CACHE[key])
    result = _mincut_partition(edges, V, world, cont, trials, seed, w)
    if len(_PARTITION_CACHE) < 64:
        _PARTITION_CACHE[key] = tuple(result)
    return result


_PARTITION_CACHE: dict = {}


def _snake_deal(order, V, world):
    part = np.zeros(V, dtype=np.int64)
    for k, g in enumerate(order):
        r, c = divmod(k, world)
        part[g] = c if r % 2 == 0 else world - 1 - c
    return part


def _mincut_partition(edges, V, world, cont, trials, seed, w=None) -> list:
    A = np.zeros((V, V), dtype=np.int32)
    for i, j in edges:
        A[int(i), int(j)] = A[int(j), int(i)] = 1
    rng = np.random.RandomState(seed)
    if w is None:
        best, best_score = cont, _map_score(edges, cont, world)
        same_w = None
    else:
        best, best_score = None, None
        same_w = w[:, None] == w[None, :]
    rows = np.arange(V)
    for t in range(max(1, trials)):
        if w is None:
            part = np.asarray(cont, dtype=np.int64)
            if t > 0:
                part = part[rng.permutation(V)]
        else:
            order = sorted(range(V), key=lambda g: (-int(w[g]), g))
            if t > 0:                                  # shuffle inside each weight class
                keyed = {}
                for g in order:
                    keyed.setdefault(int(w[g]), []).append(g)
                order = []
                for wt in sorted(keyed, reverse=True):
                    cls = keyed[wt]
                    order += [cls[k] for k in rng.permutation(len(cls))]
            part = _snake_deal(order, V, world)
        for _ in range(8 * V):
            X = np.zeros((V, world), dtype=np.int32)
            X[rows, part] = 1
            C = A @ X                                   # C[i, w]: neighbours of i in part w
            P = C - C[rows, part][:, None]
            M = P[:, part]                              # M[i, j] = P[i, part(j)]
            gain = M + M.T - 2 * A
            gain[part[:, None] == part[None, :]] = -1
            if same_w is not None:
                gain[~same_w] = -1                      # only swaps that leave every rank's weight unchanged
            k = int(np.argmax(gain))
            i, j = divmod(k, V)
            if gain[i, j] <= 0:
                break
            part[i], part[j] = part[j], part[i]
        # canonical labels: ranks in order of their smallest node id (keeps rank 0 holding node 0, deterministic)
        order = {}
        for g in range(V):
            order.setdefault(int(part[g]), len(order))
        cand = [order[int(part[g])] for g in range(V)]
        sc = _map_score(edges, cand, world)
        if best_score is None or sc < best_score:
            best, best_score = cand, sc
    return best


def phase_bounds(count: int, phases: int) -> list:
    """Local-node index ranges of the x-update phases: phase k = [b[k], b[k+1]) (contiguous, sizes differ by <= 1)."""
    return [-((-k * count) // phases) for k in range(phases)] + [count]


def build_shard_plan(G, world: int, rank: int, phases: int = 1, node_rank=None) -> ShardPlan:
    """phases > 1: every rank runs its x-updates in `phases` contiguous node blocks and posts the a = x + y of a
    block's cut-edge ends as soon as the block is done, so the transfer hides behind the next block's x-update.
    The rows of send[p] are therefore ordered by the phase of the local end, the rows of recv[p] by the phase of
    the remote end (= the peer's send order); both sides derive the same orders from (G, world, phases)."""
    edges, ptr, idx, ned, nend = graph_csr(G)
    V = G.number_of_nodes()
    nr = node_to_gpu(V, world) if node_rank is None else [int(r) for r in node_rank]
    if len(nr) != V or any(r < 0 or r >= world for r in nr):
        raise ValueError("node_rank must give a rank in [0, world) for every node")
    phases = max(1, int(phases))
    node_phase = {}
    for r in range(world):
        mine = [i for i in range(V) if nr[i] == r]
        b = phase_bounds(len(mine), phases)
        for li, g in enumerate(mine):
            node_phase[g] = max(k for k in range(phases) if b[k] <= li)
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
    # single-owner exchange: every cut edge is updated by ONE of its two ranks.  Deterministic greedy balance, the same
    # on every rank: starting from each rank's count of local edges, cut edges in G.edges() order go to the end whose
    # rank has fewer edges to update so far (tie: the min end's rank)
    owned = [0] * world
    owner_of = {}
    for e, (i, j) in enumerate(edges):
        if nr[int(i)] == nr[int(j)]:
            owner_of[e] = nr[int(i)]
            owned[nr[int(i)]] += 1
    for e, (i, j) in enumerate(edges):
        ri, rj = nr[int(i)], nr[int(j)]
        if ri != rj:
            o = ri if owned[ri] <= owned[rj] else rj
            owner_of[e] = o
            owned[o] += 1
    for le in sp.local_edges:
        le.owner = owner_of[le.e]
    sp.peers = sorted(exch)
    sp.exch = {p: sorted(exch[p]) for p in sp.peers}
    sp.phases, sp.node_phase = phases, node_phase
    for p in sp.peers:
        mine = lambda e: int(edges[e][0]) if nr[int(edges[e][0])] == rank else int(edges[e][1])      # noqa: E731
        theirs = lambda e: int(edges[e][1]) if nr[int(edges[e][0])] == rank else int(edges[e][0])    # noqa: E731
        sorder = sorted(sp.exch[p], key=lambda e: (node_phase[mine(e)], e))
        rorder = sorted(sp.exch[p], key=lambda e: (node_phase[theirs(e)], e))
        for k, e in enumerate(sp.exch[p]):
            sp.local_edges[sp.eslot[e]].xslot = k
        for k, e in enumerate(sorder):
            le = sp.local_edges[sp.eslot[e]]
            le.sslot, le.sphase = k, node_phase[mine(e)]
        for k, e in enumerate(rorder):
            sp.local_edges[sp.eslot[e]].rslot = k
        sp.send_rows[p] = [sum(1 for e in sorder if node_phase[mine(e)] < k) for k in range(phases + 1)]
        sp.recv_rows[p] = [sum(1 for e in rorder if node_phase[theirs(e)] < k) for k in range(phases + 1)]
    return sp


def post_exchange(dist, sp: ShardPlan, send: dict, recv: dict, group=None, phase=None):
    """Post one neighbour exchange and return the requests: for every peer p, send[p] ([n_cut_p, n], this rank's
    a = x + y of the cut edges shared with p, row `LocalEdge.sslot`) goes to p and p's matching buffer lands in
    recv[p] (row `LocalEdge.rslot`).  `phase=k` posts only the rows of exchange phase k (see build_shard_plan).
    Grouped P2P (ncclSend/ncclRecv inside one group on NCCL; works unchanged on gloo with CPU tensors)."""
    if not sp.peers:
        return []
    ops = []
    for p in sp.peers:
        s0, s1 = (0, len(sp.exch[p])) if phase is None else sp.send_rows[p][phase:phase + 2]
        r0, r1 = (0, len(sp.exch[p])) if phase is None else sp.recv_rows[p][phase:phase + 2]
        if s1 > s0:
            ops.append(dist.P2POp(dist.isend, send[p][s0:s1], p, group=group))
        if r1 > r0:
            ops.append(dist.P2POp(dist.irecv, recv[p][r0:r1], p, group=group))
    return dist.batch_isend_irecv(ops) if ops else []


def exchange(dist, sp: ShardPlan, send: dict, recv: dict, group=None):
    """Blocking form of post_exchange (all phases at once)."""
    for req in post_exchange(dist, sp, send, recv, group):
        req.wait()


def cut_statistics(G, world: int, node_rank=None) -> dict:
    """Edges cut by the node map (default: contiguous) and the per-rank exchange volume in units of n floats."""
    edges = graph_csr(G)[0]
    nr = node_to_gpu(G.number_of_nodes(), world) if node_rank is None else list(node_rank)
    cut = [(int(i), int(j)) for i, j in edges if nr[int(i)] != nr[int(j)]]
    per_rank = [0] * world
    for i, j in cut:
        per_rank[nr[i]] += 1
        per_rank[nr[j]] += 1
    return {"edges": len(edges), "cut": len(cut), "per_rank_ends": per_rank}
