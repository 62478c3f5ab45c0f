"""Device-resident decentralized TV-ADMM engine (block_6_admm_loop_ver2.py:15-326 on the GPU).

All state (x_i, z_ij, y_ij,*, CG vectors, TV multipliers, sinograms) lives in HBM for the whole solve; Python only
sequences C-ABI calls into libadmm_b200.so and, when the graph is sharded, the NCCL send/recv of cut-edge iterates
and the one all-reduce of the residual row.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import _native as nat
from .operators import DenseOperatorCUDA, make_plan
from .sharding import ShardPlan, build_shard_plan, partition_nodes, phase_bounds, post_exchange


def _torch():
    import torch
    return torch


def _rows(t):
    """Work weight of a node's operator: its number of projection angles, or matrix rows for a dense operator."""
    return int(t.shape[0]) if hasattr(t, "shape") and len(getattr(t, "shape", ())) == 2 else len(t)


HIST_KEYS = ("primal", "dual", "pri_per_node", "dual_per_node", "obj_per_node", "obj_total", "mse_sino_per_node",
             "mse_sino_total", "img_mse_per_node", "img_mse_total", "g_norm_history", "eps_used_history",
             "eps_target_history")


class ADMMEngine:
    """One rank's share of the solve.

    thetas      per global node angle arrays (radians) -- or dense operators (DenseOperatorCUDA / 2-D ndarrays: the
                reference's literal A_dense_list, tiny problems only)
    sinograms   per global node (M_i, D) arrays (only the local ones are uploaded)
    G           networkx graph on 0..V-1
    Q           None / float -> uniform Q_ij = q (D_i = deg_i q);  callable (i, j) -> n-vector otherwise
    Wi_list     per-pixel precisions W_i, used only when weighted_z (PDF eq. (2))
    """

    def __init__(self, thetas, sinograms, G, N, D=None, det_w=2.0, lam_tv=0.01, rho=1.0, Q=None, Wi_list=None,
                 node_prec=None, tv_mu=None, tv_sweeps=1, cg_iters=2, phantom_true=None, weighted_z=False,
                 device=0, dist=None, rank=0, world=1, group=None, node_group=None, fuse_pupdate=True,
                 max_iters=200, ax_refresh_every=10, exchange="auto", exchange_phases=None, partition="auto",
                 acceptance=False, max_tighten=2, carry_residual=False, cuda_graph="auto", impl="joseph"):
        torch = _torch()
        nat.require_cuda()
        self.torch = torch
        self.dev = torch.device(f"cuda:{device}")
        torch.cuda.set_device(self.dev)
        self.N, self.n = int(N), int(N) * int(N)
        self.D = int(D if D is not None else N)
        self.rho, self.lam = float(rho), float(lam_tv)
        self.mu = float(tv_mu if tv_mu is not None else rho)
        self.S, self.C = int(tv_sweeps), int(cg_iters)
        # a14 accept / tighten-and-retry rule (block_6_admm_loop_ver2.py:100-176), decided on the device
        self.acceptance, self.max_tighten = bool(acceptance), int(max_tighten)
        # CG residual between solves.  False (default): every solve rebuilds r = rhs0 + tvterm - H x with a
        # back-projection.  "iteration": only the first solve of an outer iteration does, later sweeps / a14 retry solves
        # take the r the TV pass carried along (r += tvterm' - tvterm); "always": also across iterations (the rhs0
        # assembly carries it; rebuilt every ax_refresh_every iterations).  Carrying saves 1.7 ms per solve at cfg 4 but
        # keeps the fp32 CG-recurrence residual instead of the true one: measured 200-iteration trace error vs
        # the fp64 oracle 1e-3 ("iteration") and 2e-4..6e-3 ("always") instead of 4e-6 -- at north_star's tolerance, so off
        # "first_retry": only the a14 rule's FIRST retry solve takes the carried residual; the last solve of an iteration
        # always rebuilds it, so whatever the carried one missed is seen (and corrected) before x leaves the iteration
        if carry_residual not in ("first_retry", "iteration", "always", False, None):
            raise ValueError("carry_residual must be 'first_retry', 'iteration', 'always' or False, "
                             f"not {carry_residual!r}")
        self.carry_r = carry_residual or False
        self.dist, self.rank, self.world, self.group = dist, int(rank), int(world), group
        self._G = G
        # NCCL exchange: posted in pieces, each right after the x-update of its block of nodes (hidden behind the next
        # block's x-update).  Peer-memory exchange: nothing to post early -- the consumer reads the remote buffer.
        # (measured on 2 and 8 B200s: splitting the x-updates into two blocks costs more in short-kernel tails than the
        # earlier transfer hides, so the default is one phase -- profiles/README.md)
        if exchange not in ("auto", "owner", "p2p", "push", "nccl"):
            raise ValueError(f"unknown exchange mode {exchange!r}")
        # single-owner exchange ("owner", the default when sharded): every cut edge is updated by ONE of its two ranks
        # (balanced, sharding.build_shard_plan); the other rank stores x of its end into the owner's buffer before the
        # edge pass and gets v = z' - y' back, written by the owner's edge kernel straight into its buffer
        self._owner = (self.world > 1 and exchange in ("auto", "owner"))
        self.phases = 1 if (self.world == 1 or exchange != "nccl") else max(1, int(exchange_phases or 1))
        self._peer_mem = exchange  # "p2p": consumers pull from the producer's buffer; "push": producers store into the consumer's
        # node -> GPU map: balanced min-cut by default (every rank computes the same map), "contiguous", or a list
        if self.world > G.number_of_nodes():
            # every rank sees the same graph, so every rank raises here -- before any collective can hang
            raise ValueError(f"{self.world} ranks but only {G.number_of_nodes()} graph nodes: a rank would own no node")
        # (balanced by node count AND by angle rows: the projector kernels' time follows the angles a rank holds)
        self.node_rank = (partition_nodes(G, self.world, partition, weights=[_rows(t) for t in thetas])
                          if isinstance(partition, str) else [int(r) for r in partition])
        self.sp: ShardPlan = build_shard_plan(G, self.world, self.rank, self.phases, self.node_rank)
        sp = self.sp
        self.Vg = sp.V
        self.loc = sp.local_nodes
        V = self.V = len(self.loc)
        if V == 0:
            raise ValueError("rank owns no node (more ranks than nodes)")
        n = self.n
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.plan = make_plan(N, [thetas[g] for g in self.loc], self.D, det_w, device, impl=impl)
        self.dense = bool(getattr(self.plan, "thetas", None) is None)
        if self.dense:      # explicit matrices: "sinograms" are plain vectors, and the fused CG staging does not apply
            self.D, fuse_pupdate = 1, 0
        elif self.plan.impl != "joseph":    # rotate-and-sum variant: plain kernels, no fused CG staging
            fuse_pupdate = 0
        A = self.A = self.plan.A
        self.node_group = int(node_group) if node_group else V
        self.ax_refresh_every = max(1, int(ax_refresh_every))

        # ---- data ------------------------------------------------------------------------------------
        b = np.concatenate([np.asarray(sinograms[g], dtype=np.float32).reshape(-1, self.D) for g in self.loc], axis=0)
        if b.shape != (A, self.D):
            raise ValueError(f"sinograms of the local nodes have shape {b.shape}, expected {(A, self.D)}")
        self.b = torch.from_numpy(np.ascontiguousarray(b)).to(self.dev)
        self.h2d_bytes = b.nbytes
        self.prec = None
        if node_prec is not None:
            self.prec = torch.tensor([float(node_prec[g]) for g in self.loc], **f32)
        self.atb = torch.empty(V, n, **f32)
        self.plan.adjoint(self.b, self.atb, prec=self.prec)
        self.xtrue = None
        if phantom_true is not None:
            self.xtrue = torch.from_numpy(np.ascontiguousarray(np.asarray(phantom_true, dtype=np.float32).reshape(-1))).to(self.dev)
            self.h2d_bytes += self.xtrue.numel() * 4

        # ---- Q: uniform scalar or per directed edge vectors -----------------------------------------------
        deg = [int(sp.nbr_ptr[g + 1] - sp.nbr_ptr[g]) for g in self.loc]
        self.q_uniform = 1.0
        self.Qdir = None
        # directed pairs (i, j) whose Q_ij this rank needs: every local node's neighbours in G.neighbors order, then --
        # single-owner exchange -- the peer's end of the cut edges this rank updates (for that end's penalty value)
        dirs = [(g, int(sp.nbr_idx[k])) for g in self.loc for k in range(sp.nbr_ptr[g], sp.nbr_ptr[g + 1])]
        if self._owner:
            for le in sp.local_edges:
                if le.peer >= 0 and le.owner == self.rank:
                    dirs.append((le.gj, le.gi) if le.i_local else (le.gi, le.gj))
        self.dirslot = {d: k for k, d in enumerate(dirs)}
        if Q is None or np.isscalar(Q):
            self.q_uniform = 1.0 if Q is None else float(Q)
        elif getattr(Q, "_admm_b200_spec", None) is not None:
            # block_3.make_precisions provider: upload the W vectors once, form Q_ij on the device
            mode, Wl = Q._admm_b200_spec
            need = sorted({g for d in dirs for g in d} | set(self.loc))
            wmap = {g: k for k, g in enumerate(need)}
            Wd = torch.from_numpy(np.stack([np.asarray(Wl[g], dtype=np.float32).reshape(-1) for g in need])).to(self.dev)
            self.h2d_bytes += Wd.numel() * 4
            ii = torch.tensor([wmap[a] for a, _ in dirs] or [0], device=self.dev)
            jj = torch.tensor([wmap[b] for _, b in dirs] or [0], device=self.dev)
            Wi_, Wj_ = Wd[ii], Wd[jj]
            qd = 0.5 * (Wi_ + Wj_) if mode == "arithmetic" else (Wi_ * Wj_) / (Wi_ + Wj_)
            self.Qdir = torch.clamp_min(qd, 1e-12).contiguous()
            del Wi_, Wj_, qd
        else:
            qv = [np.asarray(Q(a, b), dtype=np.float32).reshape(-1) for a, b in dirs]
            first = qv[0][0] if qv else 1.0
            if all(np.all(q == first) for q in qv):
                self.q_uniform = float(first)
            else:
                self.Qdir = torch.from_numpy(np.stack(qv)).to(self.dev)
                self.h2d_bytes += self.Qdir.numel() * 4
        self.rhoD_s = torch.tensor([self.rho * d * self.q_uniform for d in deg], **f32)
        self.rhoD_vec = None
        if self.Qdir is not None:
            self.rhoD_vec = torch.zeros(V, n, **f32)
            k = 0
            for li, d in enumerate(deg):      # neighbour order, like sum(np.stack(neighbor_Qs)) (block_6_ver2:139)
                for _ in range(d):
                    self.rhoD_vec[li] += self.Qdir[k]
                    k += 1
            self.rhoD_vec *= self.rho

        # ---- state ---------------------------------------------------------------------------------------
        z = lambda *s: torch.zeros(*s, **f32)  # noqa: E731
        self.x, self.r, self.p0, self.p1, self.hp = z(V, n), z(V, n), z(V, n), z(V, n), z(V, n)
        self.r1 = z(V, n)
        self.rhs0, self.tvterm = z(V, n), z(V, n)
        self.w0, self.w1 = z(V, 2, n), z(V, 2, n)
        self.q, self.ax = z(A, self.D), z(A, self.D)
        self.scal = torch.zeros(V, nat.NSCAL, dtype=torch.float64, device=self.dev)
        E = self.E = len(sp.local_edges)
        units = max(V, E, 1)
        self.part = z(units * self.plan.part_floats)
        self.counter = torch.zeros(units, dtype=torch.int32, device=self.dev)
        self.z = z(max(E, 1), n)
        self.y = z(max(E, 1), 2, n)
        self.sums = torch.zeros(max(E, 1), 5, dtype=torch.float64, device=self.dev)
        self.ROW = 2 + 8 * self.Vg   # r2, s2 and 8 per-node blocks (admm_finalize)
        self.row = torch.zeros(self.ROW, dtype=torch.float64, device=self.dev)
        self.max_iters = max(1, int(max_iters))
        self.hist = torch.zeros(self.max_iters, self.ROW, dtype=torch.float64, device=self.dev)
        self.ctl = torch.zeros(V, 4, dtype=torch.int32, device=self.dev)   # admm_node_ctl per local node
        # device-resident iteration counter: admm_accept derives eps_target from it, admm_finalize picks the history row
        # with it and increments it -- nothing host-side changes between iterations, so one iteration can be replayed
        # as a CUDA graph
        self.iter_dev = torch.zeros(1, dtype=torch.int32, device=self.dev)
        # the outer iteration as a CUDA graph (single GPU): ~15 C-ABI calls / 20-60 launches per iteration replay as
        # one submission -- what makes the small, launch-bound configs (cfg 1, 2, 5) run at kernel speed
        self.use_graph = ((self.world == 1) if cuda_graph == "auto" else bool(cuda_graph)) and not self.dense
        import os
        if os.environ.get("ADMM_B200_NOGRAPH") == "1":     # diagnostics: always launch eagerly
            self.use_graph = False
        self._graphs = {}
        self.replayed_launches = 0     # kernels launched through graph replays (the library's own counter sees only eager launches)
        self.W = None
        if weighted_z:
            if Wi_list is None:
                raise ValueError("weighted_z needs Wi_list")
            need = sorted({g for le in sp.local_edges for g in (le.gi, le.gj)})
            self.Wmap = {g: k for k, g in enumerate(need)}
            self.W = torch.from_numpy(np.stack([np.asarray(Wi_list[g], dtype=np.float32).reshape(-1) for g in need])).to(self.dev)

        # ---- cut-edge exchange: peer memory over NVLink (CUDA IPC) when available, NCCL send/recv otherwise -------
        self.exchange_mode = "none" if self.world == 1 else self._setup_peer_memory(exchange)
        self.send, self.recv = {}, {}
        if self.exchange_mode == "nccl":   # one contiguous block per peer
            self.send = {p: z(len(sp.exch[p]), n) for p in sp.peers}
            self.recv = {p: z(len(sp.exch[p]), n) for p in sp.peers}

        # the (pinned) host landing buffer of the final x download is page-locked in the background while the
        # GPU iterates (cudaHostAlloc of GBs takes longer than the copy itself)
        self._host_x = None
        self._host_thread = None
        import threading
        counts = [sum(1 for r in sp.node_rank if r == k) for k in range(self.world)]
        self._gather_rows = self.world * max(counts) if self.world > 1 else V

        def _alloc():
            try:
                self._host_x = torch.empty((self._gather_rows, n), dtype=torch.float32, pin_memory=True)
            except Exception:
                self._host_x = None
        self._host_thread = threading.Thread(target=_alloc, daemon=True)
        self._host_thread.start()

        if self._owner and self.exchange_mode == "p2p":
            self._build_tables_owner()
        else:
            self._owner = False
            self._build_tables()
        self.st = nat.State()
        self._fill_state(fuse_pupdate)
        self.k = 0
        torch.cuda.synchronize(self.dev)

    # ----------------------------------------------------------------------------------------------------
    def _addr(self, t, *idx):
        off = 0
        for i, s in zip(idx, t.stride()):
            off += i * s
        return t.data_ptr() + off * t.element_size()

    def _setup_peer_memory(self, exchange):
        """Allocate this rank's double-buffered send buffer through the C ABI (cudaMalloc + IPC handle), ship the
        handle to the peers and map theirs.  Returns "p2p" on success on every rank, else "nccl"."""
        import ctypes as C
        L, sp, dist = nat.lib(), self.sp, self.dist
        cut_sorted = sorted(le.e for le in sp.local_edges if le.peer >= 0)
        self._my_slot = {e: k for k, e in enumerate(cut_sorted)}
        self._ncut = len(cut_sorted)
        self._peer_base, self._peer_slot, self._peer_ncut = {}, {}, {}
        self._ipc_mine, self._ipc_opened = None, []
        # measured on B200s (profiles/README.md): the producer-side push on a side stream wins at every rank count once
        # the node map is the balanced min-cut one (2 GPUs: 22.6 vs 22.9 ms pull; 8 GPUs: 8.35 vs 9.0 NCCL, 9.2 pull)
        self._push = exchange in ("push", "auto", "owner")
        if exchange == "nccl":
            return "nccl"
        ok, handle = 1, b""
        try:
            ptr = C.c_void_p()
            buf = (C.c_ubyte * 64)()
            nbytes = 2 * max(self._ncut, 1) * self.n * 4
            nat.check(L.admm_ipc_alloc(nbytes, C.byref(ptr), buf), "admm_ipc_alloc")
            self._ipc_mine = ptr.value
            handle = bytes(buf)
        except Exception:
            ok = 0
        infos = [None] * self.world
        dist.all_gather_object(infos, (ok, handle, self._ncut), group=self.group)
        if all(i[0] for i in infos):
            try:
                for p in sp.peers:
                    ptr = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(infos[p][1])
                    nat.check(L.admm_ipc_open(hb, C.byref(ptr)), "admm_ipc_open")
                    self._peer_base[p] = ptr.value
                    self._ipc_opened.append(ptr.value)
                    self._peer_ncut[p] = infos[p][2]
                    sp_p = build_shard_plan(self._G, self.world, p, 1, self.node_rank)
                    self._peer_slot[p] = {e: k for k, e in enumerate(sorted(le.e for le in sp_p.local_edges if le.peer >= 0))}
            except Exception:
                ok = 0
        else:
            ok = 0
        flag = self.torch.tensor([ok], dtype=self.torch.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            self._bar = self.torch.zeros(1, dtype=self.torch.float32, device=self.dev)
            self._side = None
            import os
            if self._push and os.environ.get("ADMM_B200_PUSH_SIDE", "1") == "1":
                # the pack kernel IS the transfer (posted stores into the peers' buffers): it runs on a side stream, one
                # block per SM, beside the HBM-bound TV pass and local-edge update
                self._side = self.torch.cuda.Stream(device=self.dev)
                nsm = self.torch.cuda.get_device_properties(self.dev).multi_processor_count
                per_sm = int(os.environ.get("ADMM_B200_PUSH_BLOCKS_PER_SM", "2"))
                nat.check(L.admm_plan_set(self.plan.handle, nat.OPT_PACK_BLOCKS, nsm * per_sm), "admm_plan_set")
            return "p2p"
        self._release_ipc(sync=False)     # falling back to NCCL: give the IPC buffer and any mapped peers back now
        if exchange in ("p2p", "push", "owner"):
            raise RuntimeError("peer-memory exchange requested but CUDA IPC setup failed on some rank")
        return "nccl"

    def _remote_a(self, le, parity):
        """Device address (on THIS GPU's address space) of the remote end's a = x + y of cut edge `le`."""
        if self.exchange_mode == "p2p":
            if self._push:     # the peer stored it into MY buffer
                return self._ipc_mine + (parity * max(self._ncut, 1) + self._my_slot[le.e]) * self.n * 4
            p = le.peer
            return self._peer_base[p] + (parity * max(self._peer_ncut[p], 1) + self._peer_slot[p][le.e]) * self.n * 4
        return self._addr(self.recv[le.peer], le.rslot)

    def _pack_out(self, le, parity):
        if self.exchange_mode == "p2p":
            if self._push:     # posted NVLink stores straight into the consumer's buffer
                p = le.peer
                return self._peer_base[p] + (parity * max(self._peer_ncut[p], 1) + self._peer_slot[p][le.e]) * self.n * 4
            return self._ipc_mine + (parity * max(self._ncut, 1) + self._my_slot[le.e]) * self.n * 4
        return self._addr(self.send[le.peer], le.sslot)

    def _build_tables(self):
        torch, sp = self.torch, self.sp
        # K6 neighbour tables (G.neighbors(i) order of every local node)
        ptr, za, ya, qa = [0], [], [], []
        k = 0
        for g in self.loc:
            for kk in range(sp.nbr_ptr[g], sp.nbr_ptr[g + 1]):
                slot = sp.eslot[int(sp.nbr_edge[kk])]
                za.append(self._addr(self.z, slot))
                ya.append(self._addr(self.y, slot, int(sp.nbr_end[kk])))
                qa.append(self._addr(self.Qdir, k) if self.Qdir is not None else 0)
                k += 1
            ptr.append(len(za))
        i64 = lambda a: torch.tensor(a if len(a) else [0], dtype=torch.int64, device=self.dev)  # noqa: E731
        self.nbr_ptr = torch.tensor(ptr, dtype=torch.int32, device=self.dev)
        self.nbr_z, self.nbr_y, self.nbr_q = i64(za), i64(ya), i64(qa)
        # K5 edge descriptors, pack items, finalize tables
        ed, gi, gj, fl, packs = [], [], [], [], []
        ordered = [le for le in sp.local_edges if le.peer < 0] + [le for le in sp.local_edges if le.peer >= 0]
        self.n_edges_local = sum(1 for le in sp.local_edges if le.peer < 0)
        for le in ordered:
            s = le.slot
            xi = self._addr(self.x, sp.g2l[le.gi]) if le.i_local else 0
            xj = self._addr(self.x, sp.g2l[le.gj]) if le.j_local else 0
            yi = self._addr(self.y, s, 0) if le.i_local else 0
            yj = self._addr(self.y, s, 1) if le.j_local else 0
            ai = self._remote_a(le, 0) if not le.i_local else 0
            aj = self._remote_a(le, 0) if not le.j_local else 0
            Wi = self._addr(self.W, self.Wmap[le.gi]) if self.W is not None else 0
            Wj = self._addr(self.W, self.Wmap[le.gj]) if self.W is not None else 0
            qij = self._addr(self.Qdir, self.dirslot[(le.gi, le.gj)]) if (self.Qdir is not None and le.i_local) else 0
            qji = self._addr(self.Qdir, self.dirslot[(le.gj, le.gi)]) if (self.Qdir is not None and le.j_local) else 0
            ed.append([xi, xj, yi, yj, self._addr(self.z, s), ai, aj, Wi, Wj, qij, qji, 0, 0])
            gi.append(le.gi)
            gj.append(le.gj)
            fl.append((1 if le.i_local else 0) | (2 if le.j_local else 0) | (4 if le.owns_dual else 0))
            if le.peer >= 0:
                packs.append((le.sphase, [xi, yi, self._pack_out(le, 0)] if le.i_local else [xj, yj, self._pack_out(le, 0)],
                              le))
        # phase-major; inside a phase rank r serves peer r+1 first, then r+2, ... so that at any moment every GPU is the
        # destination of one producer (the narrow push kernel walks the items in this order)
        packs.sort(key=lambda t: (t[0], (t[2].peer - self.rank) % self.world))
        self.pack_rows = [sum(1 for t in packs if t[0] < k) for k in range(self.phases + 1)]
        pack_les = [t[2] for t in packs]
        packs = [t[1] for t in packs]
        self.edge_desc = torch.tensor(ed if ed else [[0] * 13], dtype=torch.int64, device=self.dev)
        self.E_run = len(ordered)
        i32 = lambda a: torch.tensor(a if len(a) else [0], dtype=torch.int32, device=self.dev)  # noqa: E731
        self.edge_gi, self.edge_gj, self.edge_fl = i32(gi), i32(gj), i32(fl)
        epos_of = {le.e: k for k, le in enumerate(ordered)}       # global edge id -> position in the edge arrays
        fptr, fepos, fend = [0], [], []
        for g in self.loc:
            for kk in range(sp.nbr_ptr[g], sp.nbr_ptr[g + 1]):
                fepos.append(epos_of[int(sp.nbr_edge[kk])])
                fend.append(int(sp.nbr_end[kk]))
            fptr.append(len(fepos))
        self.fin_ptr, self.fin_epos, self.fin_end = i32(fptr), i32(fepos), i32(fend)
        self.node_gid = i32(self.loc)
        self.pack_desc = torch.tensor(packs if packs else [[0, 0, 0]], dtype=torch.int64, device=self.dev)
        self.n_pack = len(packs)
        self.edge_desc_par = [self.edge_desc, self.edge_desc]
        self.pack_desc_par = [self.pack_desc, self.pack_desc]
        if self.exchange_mode == "p2p":   # second parity of the double-buffered peer-memory exchange
            ed1 = self.edge_desc.clone()
            pk1 = self.pack_desc.clone()
            for pos, le in enumerate(ordered):
                if le.peer >= 0:
                    col = 5 if not le.i_local else 6           # admm_edge.ai / .aj
                    ed1[pos, col] = self._remote_a(le, 1)
            for k, le in enumerate(pack_les):
                pk1[k, 2] = self._pack_out(le, 1)
            self.edge_desc_par = [self.edge_desc, ed1]
            self.pack_desc_par = [self.pack_desc, pk1]

    def _build_tables_owner(self):
        """Tables of the single-owner exchange.  This rank's edge pass covers its local edges and the cut edges it owns;
        for an owned cut edge the peer's x arrives in this rank's buffer (slot of the edge) and v = z' - y' of the
        peer's end is stored into the PEER's buffer by the edge kernel.  For a cut edge owned by the peer this rank only
        sends x of its end and its rhs0 reads the returned v (as "z", with a zero "y")."""
        torch, sp, n = self.torch, self.sp, self.n
        mine = lambda le: le.peer < 0 or le.owner == self.rank      # noqa: E731
        inbox = lambda e: self._ipc_mine + self._my_slot[e] * n * 4    # noqa: E731
        peer_inbox = lambda le: self._peer_base[le.peer] + self._peer_slot[le.peer][le.e] * n * 4   # noqa: E731
        self._zero = torch.zeros(n, dtype=torch.float32, device=self.dev)
        # K6 neighbour tables
        ptr, za, ya, qa = [0], [], [], []
        for g in self.loc:
            for kk in range(sp.nbr_ptr[g], sp.nbr_ptr[g + 1]):
                le = sp.local_edges[sp.eslot[int(sp.nbr_edge[kk])]]
                if mine(le):
                    za.append(self._addr(self.z, le.slot))
                    ya.append(self._addr(self.y, le.slot, int(sp.nbr_end[kk])))
                else:
                    za.append(inbox(le.e))
                    ya.append(self._zero.data_ptr())
                qa.append(self._addr(self.Qdir, self.dirslot[(g, int(sp.nbr_idx[kk]))]) if self.Qdir is not None else 0)
            ptr.append(len(za))
        i64 = lambda a: torch.tensor(a if len(a) else [0], dtype=torch.int64, device=self.dev)  # noqa: E731
        i32 = lambda a: torch.tensor(a if len(a) else [0], dtype=torch.int32, device=self.dev)  # noqa: E731
        self.nbr_ptr = torch.tensor(ptr, dtype=torch.int32, device=self.dev)
        self.nbr_z, self.nbr_y, self.nbr_q = i64(za), i64(ya), i64(qa)
        # K5 descriptors: local edges first (they run under the push), then the owned cut edges
        ordered = [le for le in sp.local_edges if le.peer < 0] + [le for le in sp.local_edges if le.peer >= 0 and mine(le)]
        self.n_edges_local = sum(1 for le in sp.local_edges if le.peer < 0)
        ed, gi, gj, fl = [], [], [], []
        qaddr = lambda a, b: self._addr(self.Qdir, self.dirslot[(a, b)]) if self.Qdir is not None else 0   # noqa: E731
        for le in ordered:
            s = le.slot
            xi = self._addr(self.x, sp.g2l[le.gi]) if le.i_local else inbox(le.e)
            xj = self._addr(self.x, sp.g2l[le.gj]) if le.j_local else inbox(le.e)
            Wi = self._addr(self.W, self.Wmap[le.gi]) if self.W is not None else 0
            Wj = self._addr(self.W, self.Wmap[le.gj]) if self.W is not None else 0
            vi = peer_inbox(le) if not le.i_local else 0
            vj = peer_inbox(le) if not le.j_local else 0
            ed.append([xi, xj, self._addr(self.y, s, 0), self._addr(self.y, s, 1), self._addr(self.z, s), 0, 0, Wi, Wj,
                       qaddr(le.gi, le.gj), qaddr(le.gj, le.gi), vi, vj])
            gi.append(le.gi)
            gj.append(le.gj)
            fl.append(1 | 2 | 4 | (8 if not le.i_local else 0) | (16 if not le.j_local else 0))
        self.edge_desc = torch.tensor(ed if ed else [[0] * 13], dtype=torch.int64, device=self.dev)
        self.E_run = len(ordered)
        self.edge_gi, self.edge_gj, self.edge_fl = i32(gi), i32(gj), i32(fl)
        epos_of = {le.e: k for k, le in enumerate(ordered)}
        fptr, fepos, fend = [0], [], []
        for g in self.loc:
            for kk in range(sp.nbr_ptr[g], sp.nbr_ptr[g + 1]):
                fepos.append(epos_of.get(int(sp.nbr_edge[kk]), -1))     # -1: the owning peer adds this node's pieces
                fend.append(int(sp.nbr_end[kk]))
            fptr.append(len(fepos))
        self.fin_ptr, self.fin_epos, self.fin_end = i32(fptr), i32(fepos), i32(fend)
        self.node_gid = i32(self.loc)
        # push items: x of this rank's end of every cut edge the peer owns (y = 0: plain copy), peers served in rotation
        sends = [le for le in sp.local_edges if le.peer >= 0 and not mine(le)]
        sends.sort(key=lambda le: ((le.peer - self.rank) % self.world, le.e))
        packs = [[self._addr(self.x, sp.g2l[le.gi if le.i_local else le.gj]), 0, peer_inbox(le)] for le in sends]
        self.pack_rows = [0, len(packs)]
        # the x push is a set of plain copies: issued through the copy engines (one cudaMemcpyAsync per item on the side
        # stream; measured at 8 GPUs: 6.85 vs 7.01 ms per iteration with the SM copy kernel, which competes with the TV
        # and edge kernels for the SMs).  ADMM_B200_PUSH_CE=0 selects the kernel
        import os
        self._push_ce = os.environ.get("ADMM_B200_PUSH_CE", "1") == "1"
        self._pack_host = np.ascontiguousarray(np.array(packs if packs else [[0, 0, 0]], dtype=np.uint64))
        self.pack_desc = torch.tensor(packs if packs else [[0, 0, 0]], dtype=torch.int64, device=self.dev)
        self.n_pack = len(packs)
        self.edge_desc_par = [self.edge_desc, self.edge_desc]      # single-buffered: two barriers per iteration order it
        self.pack_desc_par = [self.pack_desc, self.pack_desc]

    def _fill_state(self, fuse):
        st = self.st
        for name in ("x", "r", "r1", "p0", "p1", "hp", "rhs0", "tvterm", "atb", "w0", "w1", "q", "ax", "b", "scal", "part",
                     "counter", "rhoD_s"):
            setattr(st, name, getattr(self, name).data_ptr())
        st.ctl = self.ctl.data_ptr()
        st.iter_dev = self.iter_dev.data_ptr()
        st.hist_stride = self.ROW
        st.masked = 0
        st.rhoD_vec = self.rhoD_vec.data_ptr() if self.rhoD_vec is not None else None
        st.prec = self.prec.data_ptr() if self.prec is not None else None
        st.xtrue = self.xtrue.data_ptr() if self.xtrue is not None else None
        st.stride = self.n
        st.rho, st.lam, st.mu, st.q_uniform = self.rho, self.lam, self.mu, self.q_uniform
        st.w_parity = 0
        st.fuse_pupdate = int(fuse) if not isinstance(fuse, bool) else (2 if fuse else 0)
        # sharded runs hold the last TV pass back so that the cut-edge exchange starts as soon as x is final.  With the
        # a14 rule only the LAST retry solve's TV pass can be held back (the earlier ones feed the decisions)
        st.defer_tv = 1 if (self.world > 1 and not self.acceptance) else 0
        self._defer_last_tv = (self.world > 1 and self.acceptance and self.max_tighten >= 1 and self.phases == 1
                               and self.node_group >= self.V)

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    # ---- one outer iteration ------------------------------------------------------------------------------
    def nodes_phase(self):
        """rhs0 assembly (K6) + x-updates (K1-K4) of every local node, node group by node group.  When the graph is
        sharded the last TV pass is deferred (see step) so the exchange can start as soon as x is final."""
        L, h, st = nat.lib(), self.plan.handle, self.st
        sref = ctypes.byref(st)
        # A x is carried by the CG recurrence (ax += alpha A p); it is re-projected every `ax_refresh_every` iterations
        st.reuse_ax = 0 if (self.k % self.ax_refresh_every == 0) else 1
        # likewise the CG residual: the TV pass and the rhs0 assembly carry r = rhs0 + tvterm - H x along, so a solve
        # starts without a back-projection; rebuilt from scratch on the refresh iterations
        # bit 0: a solve may take the carried residual; bit 1: its TV pass hands the residual on to the next solve
        st.carry_r = 3 if self.carry_r else 0
        st.reuse_r = st.reuse_ax if self.carry_r == "always" else 0
        nat.check(L.admm_rhs0(h, sref, self.nbr_ptr.data_ptr(), self.nbr_z.data_ptr(), self.nbr_y.data_ptr(),
                              self.nbr_q.data_ptr(), 0, self.V, self._stream()), "admm_rhs0")
        reqs = []
        bounds = phase_bounds(self.V, self.phases)
        eps_target = 2.0 / ((self.k + 1) ** 1.005)          # block_6_admm_loop_ver2.py:101-103
        for ph in range(self.phases):
            for n0 in range(bounds[ph], bounds[ph + 1], self.node_group):
                nn = min(self.node_group, bounds[ph + 1] - n0)
                if not self.acceptance:
                    nat.check(L.admm_x_update(h, sref, n0, nn, self.S, self.C, self._stream()), "admm_x_update")
                    continue
                # :155-176 on the device: nodes whose |g| misses the target are solved again (warm start, masked launches),
                # at most max_tighten times; no host round trip.  The decision itself is taken inside the TV pass that
                # ends each solve (by the last block of every node), |A x - b|^2 is refreshed by the last solve only
                st.max_tighten, st.eps_target = self.max_tighten, eps_target
                st.accept_mode, st.skip_mse = 1, (1 if self.max_tighten > 0 else 0)
                nat.check(L.admm_x_update(h, sref, n0, nn, self.S, self.C, self._stream()), "admm_x_update")
                keep = (st.reuse_ax, st.reuse_r, st.carry_r)
                st.masked, st.reuse_ax, st.reuse_r, st.accept_mode = 1, 1, (1 if st.carry_r else 0), 2
                for t in range(self.max_tighten):
                    last = (t == self.max_tighten - 1)
                    if self.carry_r == "first_retry":
                        st.reuse_r = 1 if (t == 0 and not last) else 0
                    if last and self.carry_r and self.carry_r != "always":
                        st.carry_r = 1      # the next iteration's first solve rebuilds r: this solve's TV pass need not hand it on
                    # sharded: x is final after the LAST retry's CG, so its TV pass (w, tvterm, |g| only) is held back and
                    # runs under the cut-edge exchange (tv_phase); needs the whole rank in one node group
                    st.defer_tv = 1 if (self._defer_last_tv and last) else 0
                    st.skip_mse = 0 if last else 1
                    nat.check(L.admm_x_update(h, sref, n0, nn, self.S, self.C, self._stream()), "admm_x_update")
                if not self._defer_last_tv:
                    st.masked, st.accept_mode = 0, 0
                st.reuse_ax, st.reuse_r, st.skip_mse = keep[0], keep[1], 0
                if not self._defer_last_tv:
                    st.carry_r = keep[2]      # (a deferred last TV pass still needs this solve's setting: reset next iteration)
            if self.phases > 1:
                if ph == self.phases - 1 and getattr(self, "time_exchange", False):
                    self._ex_t0 = self.torch.cuda.Event(enable_timing=True)
                    self._ex_t0.record()
                reqs += self.exchange_start(ph)   # this block's x is final: its transfer hides behind the next block
        return reqs

    def tv_phase(self):
        """The deferred last TV pass (K3) of every local node (with the a14 rule: of the nodes still active in the last
        retry solve, followed by the final acceptance bookkeeping)."""
        st = self.st
        L, h, sref = nat.lib(), self.plan.handle, ctypes.byref(self.st)
        # flags: 1 diagnostics; 2 the solve's last CG update left r <- r - alpha Hp to this pass; 4 ... and r sits in r1
        # (the fully fused CG ping-pongs r / r1 once per iteration after the first)
        flags = 1
        if self.C > 0:
            flags |= 2
            if int(st.fuse_pupdate) == 2 and (self.C - 1) % 2 == 1:
                flags |= 4
        if self.acceptance:     # masked pass of the last retry solve; carries that solve's a14 bookkeeping (accept_mode 2)
            nat.check(L.admm_tv_pass(h, sref, 0, self.V, flags, self._stream()), "admm_tv_pass")
            st.masked, st.defer_tv, st.accept_mode = 0, 0, 0
            return
        nat.check(L.admm_tv_pass(h, sref, 0, self.V, flags, self._stream()), "admm_tv_pass")

    def exchange_start(self, phase=None):
        """Pack a = x + y of this rank's cut-edge ends (of exchange phase `phase`, or all).  NCCL mode: post the grouped
        send/recv and return the requests.  Peer-memory mode: the pack kernel writes straight into the IPC-shared
        buffer the peers read from."""
        if self.world == 1:
            return []
        par = self.k & 1 if self.exchange_mode == "p2p" else 0
        k0, k1 = (0, self.n_pack) if phase is None else self.pack_rows[phase:phase + 2]
        if self.exchange_mode == "p2p" and self._side is not None:
            torch = self.torch
            main = torch.cuda.current_stream()
            self._side.wait_stream(main)          # x is final
            if k1 > k0 and self._owner and getattr(self, "_push_ce", False):
                nat.check(nat.lib().admm_push_copy(self.plan.handle, self._pack_host.ctypes.data + k0 * 24, k1 - k0,
                                                   ctypes.c_void_p(self._side.cuda_stream)), "admm_push_copy")
            elif k1 > k0:
                nat.check(nat.lib().admm_pack(self.plan.handle, self.pack_desc_par[par].data_ptr() + k0 * 24, k1 - k0,
                                              ctypes.c_void_p(self._side.cuda_stream)), "admm_pack")
            self._pushed = torch.cuda.Event()
            self._pushed.record(self._side)
            return []
        if k1 > k0:
            nat.check(nat.lib().admm_pack(self.plan.handle, self.pack_desc_par[par].data_ptr() + k0 * 24, k1 - k0,
                                          self._stream()), "admm_pack")
        if self.exchange_mode == "p2p":
            return []
        return post_exchange(self.dist, self.sp, self.send, self.recv, self.group, phase=phase)

    def edges_phase(self, reqs=()):
        """K5 on the local edges (overlaps the exchange), then on the cut edges, then the residual row."""
        L, h = nat.lib(), self.plan.handle
        sref = ctypes.byref(self.st)
        nl, E = self.n_edges_local, self.E_run
        esz = 13 * 8
        par = self.k & 1 if self.exchange_mode == "p2p" else 0
        desc = self.edge_desc_par[par]
        if nl:
            nat.check(L.admm_edge_update(h, sref, desc.data_ptr(), nl, self.sums.data_ptr(), self._stream()),
                      "admm_edge_update")
        if self.exchange_mode == "p2p":
            if self._side is not None:
                self.torch.cuda.current_stream().wait_event(self._pushed)
            # device-side barrier: every rank's pack of this iteration has completed before anyone reads it (the double
            # buffer makes this the only synchronisation the exchange needs).  The barrier IS the all-reduce of the
            # previous iteration's residual row, which nothing on the device needs any earlier: ONE collective per
            # iteration instead of two (a dedicated one-float all-reduce only when no row is pending)
            if not self._flush_row():
                self.dist.all_reduce(self._bar, group=self.group)
            # (single-owner exchange: the row was reduced right after the edge pass -- that collective is the barrier
            #  before anyone's next rhs0 reads the returned v -- so this is the dedicated one)
        for r in reqs:
            r.wait()          # stream-level wait: the compute stream now depends on the received buffers
        if getattr(self, "_edges_timed", None):
            # pack + TV + local edges + whatever of the exchange they did not hide (bench diagnostics)
            ea, eb = self._edges_timed
            eb.record()
            self.exchange_events = getattr(self, "exchange_events", []) + [(ea, eb)]
            self._edges_timed = None
        if E - nl:
            nat.check(L.admm_edge_update(h, sref, desc.data_ptr() + nl * esz, E - nl,
                                         self.sums.data_ptr() + nl * 5 * 8, self._stream()), "admm_edge_update")
        # history row k is written in place (device-side row index = the device's iteration counter)
        nat.check(L.admm_finalize(h, sref, self.sums.data_ptr(), self.edge_gi.data_ptr(), self.edge_gj.data_ptr(),
                                  self.edge_fl.data_ptr(), E, self.n_edges_local, self.node_gid.data_ptr(),
                                  self.fin_ptr.data_ptr(), self.fin_epos.data_ptr(), self.fin_end.data_ptr(), self.Vg,
                                  self.hist.data_ptr(), self._stream()), "admm_finalize")
        if self.world > 1:
            # the only data collective: the sum of the ranks' rows.  Peer-memory exchange: deferred to the
            # next iteration's barrier (or to whoever reads the residuals first); NCCL exchange: right away
            self._pending_row = self.k
            if self.exchange_mode != "p2p" or self._owner:
                self._flush_row()

    def _flush_row(self):
        """All-reduce the residual row of the last finished iteration in place in the history (collective: every rank
        reaches it at the same point of the same loop).  Returns False when nothing was pending."""
        k = getattr(self, "_pending_row", None)
        if k is None or self.world == 1:
            return False
        self.dist.all_reduce(self.hist[k], group=self.group)
        self._pending_row = None
        return True

    def _grow_history(self):
        if self.k >= self.hist.shape[0]:      # engine re-used past its first max_iters: grow the history buffer
            self._flush_row()
            grown = self.torch.zeros(2 * self.hist.shape[0], self.ROW, dtype=self.torch.float64, device=self.dev)
            grown[: self.hist.shape[0]] = self.hist
            self.hist = grown
            self._graphs = {}                 # captured graphs hold the old buffer's address

    def _step_body(self):
        reqs = self.nodes_phase()
        if self.world > 1:
            timed = getattr(self, "time_exchange", False)
            if self.phases == 1:
                if timed:
                    self._ex_t0 = self.torch.cuda.Event(enable_timing=True)
                    self._ex_t0.record()
                reqs = self.exchange_start()  # x is final: the exchange runs under the TV pass and the local edges
            if self.st.defer_tv or self._defer_last_tv:
                self.tv_phase()
            if timed:
                self._edges_timed = (self._ex_t0, self.torch.cuda.Event(enable_timing=True))
        self.edges_phase(reqs)

    def step(self):
        self._grow_history()
        if self.use_graph and self.world == 1 and not nat.profiling():
            # two variants of the iteration exist: with and without the periodic re-projection of A x
            key = (self.k % self.ax_refresh_every == 0)
            g = self._graphs.get(key)
            if g is None and self.k >= 1:     # iteration 0 runs eagerly (lazy one-time setup inside the library)
                g = self._capture()
                self._graphs[key] = g
            if g:
                g[0].replay()
                self.replayed_launches += g[1]
                self.k += 1
                return
        self._step_body()
        self.k += 1

    def _capture(self):
        """Record one outer iteration (every C-ABI call launches on torch's current stream, which is the capturing
        stream here).  Returns the graph, or False (and graphs stay off) if the capture fails."""
        torch = self.torch
        try:
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            l0 = nat.launch_count()
            with torch.cuda.graph(g):
                self._step_body()
            return (g, nat.launch_count() - l0)   # the capture pass counted the graph's kernel nodes, launched nothing
        except Exception as e:               # pragma: no cover
            import warnings
            warnings.warn(f"admm_b200: CUDA-graph capture of the outer iteration failed ({e}); running eagerly")
            self.use_graph = False
            torch.cuda.synchronize(self.dev)
            return False

    # ---- results --------------------------------------------------------------------------------------------
    def residuals(self):
        """(primal, dual) norms of the last completed iteration -- forces a device sync."""
        self._flush_row()
        r = self.hist[self.k - 1, :2].cpu().numpy() if self.k > 0 else np.zeros(2)
        return math.sqrt(r[0]), math.sqrt(r[1])

    def history(self, iters=None):
        """History dict with the keys of block_6_admm_loop_ver2.py:310-326 (one entry per iteration)."""
        self._flush_row()
        iters = self.k if iters is None else min(int(iters), self.k)
        H = self.hist[:iters].cpu().numpy()
        Vg = self.Vg
        sl = lambda k: H[:, 2 + k * Vg: 2 + (k + 1) * Vg]  # noqa: E731
        pri, dual, pen, mse, tv, gn2, img, tries = (sl(k) for k in range(8))
        out = {k: [] for k in HIST_KEYS}
        out["tighten_history"] = []
        obj = 0.5 * self._prec_global()[None, :] * mse + self.lam * tv + 0.5 * self.rho * pen
        for k in range(iters):
            eps_target = 2.0 / ((k + 1) ** 1.005)  # block_6_admm_loop_ver2.py:101-103
            out["primal"].append(math.sqrt(H[k, 0]))
            out["dual"].append(math.sqrt(H[k, 1]))
            out["pri_per_node"].append(np.sqrt(pri[k]))
            out["dual_per_node"].append(np.sqrt(dual[k]))
            out["obj_per_node"].append(obj[k].copy())
            out["obj_total"].append(float(np.sum(obj[k])))
            out["mse_sino_per_node"].append(mse[k].copy())
            out["mse_sino_total"].append(float(np.sum(mse[k])))
            out["img_mse_per_node"].append(img[k].copy())
            out["img_mse_total"].append(float(np.sum(img[k])))
            out["g_norm_history"].append(np.sqrt(gn2[k]))
            # :106-108,161,170,176: eps of the accepted try = min(1e-2, eps_target) / 5^(tries the device really spent)
            out["eps_used_history"].append(min(1e-2, eps_target) / 5.0 ** tries[k])
            out["eps_target_history"].append(np.full(Vg, eps_target))
            out["tighten_history"].append(tries[k].astype(np.int64))
        return out

    def _prec_global(self):
        p = np.ones(self.Vg)
        if getattr(self, "_node_prec_all", None) is not None:
            p = np.asarray(self._node_prec_all, dtype=np.float64)
        return p

    def x_local(self):
        return self.x.cpu().numpy()

    def x_all(self, gather="all"):
        """x of every node on every rank (gather="all") or on rank 0 only (gather="rank0": the other ranks get their
        own nodes' x and None elsewhere, and skip the full device->host copy): list of V float32 arrays of length n (views of one pinned host buffer;
        the reference's are float64 -- `np.stack`, `.reshape(N, N)` and arithmetic behave the same).  When the graph
        is sharded this is a collective (all_gather over NVLink) and every rank returns the full list."""
        torch = self.torch
        if self._host_thread is not None:
            self._host_thread.join()
            self._host_thread = None
        host = self._host_x
        if host is None:
            host = torch.empty((self._gather_rows, self.n), dtype=torch.float32, pin_memory=True)
        self._host_x = None          # the caller owns the views from here on
        if self.world == 1:
            host.copy_(self.x, non_blocking=True)
            torch.cuda.synchronize(self.dev)
            arr = host.numpy()
            return [arr[i] for i in range(self.V)]
        counts = [sum(1 for r in self.sp.node_rank if r == k) for k in range(self.world)]
        mx = max(counts)
        pad = torch.zeros(mx, self.n, dtype=torch.float32, device=self.dev)
        pad[: self.V] = self.x
        gathered = torch.empty(self.world * mx, self.n, dtype=torch.float32, device=self.dev)
        self.dist.all_gather_into_tensor(gathered, pad, group=self.group)
        if gather == "rank0" and self.rank != 0:
            lo = self.rank * mx
            host[lo:lo + self.V].copy_(gathered[lo:lo + self.V], non_blocking=True)
            torch.cuda.synchronize(self.dev)
            arr = host.numpy()
            out = [None] * self.Vg
            for li, g in enumerate(self.loc):
                out[g] = arr[lo + li]
            return out
        host.copy_(gathered, non_blocking=True)
        torch.cuda.synchronize(self.dev)
        arr = host.numpy()
        out = [None] * self.Vg
        seen = [0] * self.world
        for g in range(self.Vg):          # rank k's rows are its nodes in ascending global id
            k = self.sp.node_rank[g]
            out[g] = arr[k * mx + seen[k]]
            seen[k] += 1
        return out

    def _release_ipc(self, sync):
        """Unmap the peers' buffers and free this rank's (`sync`: barrier so nobody unmaps while a peer still reads;
        off on the error path, where the peers may never arrive)."""
        L = nat.lib()
        if sync:
            self.torch.cuda.synchronize(self.dev)
            self.dist.barrier(group=self.group)
        for ptr in getattr(self, "_ipc_opened", []):
            L.admm_ipc_close(ctypes.c_void_p(ptr))
        self._ipc_opened = []
        if sync:
            self.dist.barrier(group=self.group)
        if getattr(self, "_ipc_mine", None):
            L.admm_ipc_free(ctypes.c_void_p(self._ipc_mine))
            self._ipc_mine = None

    def close(self, sync=True):
        if self.world > 1 and (getattr(self, "_ipc_mine", None) or getattr(self, "_ipc_opened", None)):
            self._release_ipc(sync and getattr(self, "exchange_mode", "") == "p2p")
        self.plan.close()


def solve(engine: ADMMEngine, max_iters, eps_pri, eps_dual, verbose=False, stop=True, check_every=1,
          snapshot=None):
    """Run the outer loop with the reference's stop test (block_6_admm_loop_ver2.py:286-289)."""
    done = 0
    for k in range(max_iters):
        engine.step()
        done = k + 1
        need = (stop and ((k + 1) % check_every == 0)) or (verbose and k % 10 == 0)
        if need:
            pri, dual = engine.residuals()
            if verbose and k % 10 == 0:
                print(f"iter {k}, primal {pri:.3e}, dual {dual:.3e}")
            if stop and pri < eps_pri and dual < eps_dual:
                if verbose:
                    print(f"stopped at iter {k}, primal {pri:.3e}, dual {dual:.3e}")
                break
        if snapshot is not None:
            snapshot(k, engine)
    return done


class NodeProblem:
    """One node's subproblem, eq. (1) (block_5_node_problem.py:21-29), on the GPU:
        min_x 1/2 |A x - b|^2 + lam TV(x) + rho/2 sum_j |x - v_j|^2_{Q_j}
    solved by TV-split sweeps + CG through the same C-ABI x-update the outer loop uses (a graph with one node)."""

    def __init__(self, op, b, rho, neighbor_terms, N, lam_tv, Qij_terms, tv_mu=None, device=None, prec=1.0):
        torch = _torch()
        nat.require_cuda()
        if isinstance(op, np.ndarray) and op.ndim == 2:
            op = DenseOperatorCUDA(op, N)          # the reference's literal dense Ai (block_5_node_problem.py:8)
        if not (hasattr(op, "angles") or isinstance(op, DenseOperatorCUDA)):
            raise TypeError("Ai must be a RayTransformCUDA operator or a dense (m, n) matrix")
        device = op.device if device is None else device
        self.torch, self.dev = torch, torch.device(f"cuda:{device}")
        torch.cuda.set_device(self.dev)
        self.N, self.n, self.D = int(N), int(N) * int(N), op.D
        self.rho, self.lam = float(rho), float(lam_tv)
        self.mu = float(tv_mu) if tv_mu is not None else (self.rho if self.rho > 0 else 1.0)
        self.plan = make_plan(N, [op], op.D, op.det_w, device)
        plain = isinstance(op, DenseOperatorCUDA) or getattr(self.plan, "impl", "joseph") != "joseph"
        n, f32 = self.n, dict(dtype=torch.float32, device=self.dev)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))).to(self.dev)  # noqa: E731
        self.b = up(b).reshape(self.plan.A, self.D)
        self.prec = torch.tensor([float(prec)], **f32)
        self.atb = torch.empty(1, n, **f32)
        self.plan.adjoint(self.b, self.atb, prec=self.prec)
        deg = len(neighbor_terms)
        if len(Qij_terms) != deg:
            raise ValueError("neighbor_terms and Qij_terms must have the same length")
        self.v = torch.stack([up(v) for v in neighbor_terms]) if deg else torch.zeros(1, n, **f32)
        self.qv = torch.stack([up(q) for q in Qij_terms]) if deg else torch.zeros(1, n, **f32)
        self.zero = torch.zeros(n, **f32)
        self.deg = deg
        z = lambda *s: torch.zeros(*s, **f32)  # noqa: E731
        self.x, self.r, self.p0, self.p1, self.hp = z(1, n), z(1, n), z(1, n), z(1, n), z(1, n)
        self.r1 = z(1, n)
        self.rhs0, self.tvterm = z(1, n), z(1, n)
        self.w0, self.w1 = z(1, 2, n), z(1, 2, n)
        self.q, self.ax = z(self.plan.A, self.D), z(self.plan.A, self.D)
        self.scal = torch.zeros(1, nat.NSCAL, dtype=torch.float64, device=self.dev)
        self.part = z(self.plan.part_floats)
        self.counter = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.rhoD_vec = (self.rho * self.qv[:deg].sum(dim=0)).reshape(1, n).contiguous() if deg else z(1, n)
        self.rhoD_s = z(1)
        st = self.st = nat.State()
        for name in ("x", "r", "r1", "p0", "p1", "hp", "rhs0", "tvterm", "atb", "w0", "w1", "q", "ax", "b", "scal", "part",
                     "counter", "rhoD_s", "rhoD_vec", "prec"):
            setattr(st, name, getattr(self, name).data_ptr())
        st.xtrue = None
        st.stride, st.rho, st.lam, st.mu, st.q_uniform = n, self.rho, self.lam, self.mu, 1.0
        st.w_parity, st.fuse_pupdate = 0, (0 if plain else 2)
        ptr = torch.tensor([0, deg], dtype=torch.int32, device=self.dev)
        a = lambda t, k: t.data_ptr() + k * n * 4  # noqa: E731
        i64 = lambda v: torch.tensor(v if v else [0], dtype=torch.int64, device=self.dev)  # noqa: E731
        self._tabs = (ptr, i64([a(self.v, k) for k in range(deg)]), i64([self.zero.data_ptr()] * deg),
                      i64([a(self.qv, k) for k in range(deg)]))
        s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        nat.check(nat.lib().admm_rhs0(self.plan.handle, ctypes.byref(st), ptr.data_ptr(), self._tabs[1].data_ptr(),
                                      self._tabs[2].data_ptr(), self._tabs[3].data_ptr(), 0, 1, s), "admm_rhs0")
        self.cg_done = 0

    def sweep(self, cg_iters, sweeps=1):
        s = ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)
        nat.check(nat.lib().admm_x_update(self.plan.handle, ctypes.byref(self.st), 0, 1, sweeps, cg_iters, s),
                  "admm_x_update")
        self.st.w_parity ^= (sweeps & 1)
        self.cg_done += sweeps * cg_iters

    def stats(self):
        """(objective of eq. (1) with canonical TV, stationarity norm |g|) of the current x."""
        sc = self.scal[0].cpu().numpy()
        x = self.x[0]
        pen = 0.0
        if self.deg:
            pen = float(((x[None, :] - self.v[: self.deg]) ** 2 * self.qv[: self.deg]).sum().item())
        obj = 0.5 * float(self.prec.item()) * sc[nat.S_MSE] + self.lam * sc[nat.S_TV] + 0.5 * self.rho * pen
        return obj, math.sqrt(max(sc[nat.S_GN2], 0.0))

    def x_value(self):
        return self.x[0].cpu().numpy().astype(np.float64)

    def close(self):
        self.plan.close()
