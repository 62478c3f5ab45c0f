#!/usr/bin/env python
"""bench.py -- ADMM iterations/sec of the decentralized TV-ADMM tomography solve on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg4] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one outer ADMM iteration: all V node x-updates (tv_sweeps x cg_iters CG iterations on
A^T P A + rho D + mu K^T K with the Joseph projector pair) + all E edge z / y updates + the global residuals
(SURVEY.md 8(d)).  Default workload = BASELINE.json configs[3], the one the north-star target is quoted on
(2048^2, 720 angles, 64 nodes, Erdos-Renyi graph): it fits one GPU; with N > 1 the same problem is sharded
(strong scaling).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "distributed-inverse-problem-admm_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

CONFIGS = {
    # BASELINE.json configs[0..3]; SURVEY 8(d) fixes graphs / noise / Q
    "cfg1": dict(N=128, M=180, V=4, graph="ring", hetero=False, wq=False,
                 desc="Shepp-Logan 128x128, 180 angles over 4 ring nodes"),
    "cfg2": dict(N=512, M=360, V=16, graph="regular", hetero=False, wq=False,
                 desc="Shepp-Logan 512x512, 360 angles, 16 nodes, random 4-regular graph, uniform precisions"),
    "cfg3": dict(N=1024, M=720, V=32, graph="regular", hetero=True, wq=True,
                 desc="1024x1024, 720 angles, 32 nodes, heterogeneous noise precisions (weighted LS), block_3 arithmetic Q"),
    "cfg4": dict(N=2048, M=720, V=64, graph="er", hetero=False, wq=False,
                 desc="Shepp-Logan 2048x2048, 720 angles, 64 nodes, connected Erdos-Renyi(p=0.1) graph, uniform precisions"),
    # configs[4]: slice-parallel batch = disjoint union of per-slice graphs; batch dimension = slice x node
    "cfg5": dict(N=512, M=360, V=16, graph="regular", hetero=False, wq=False, slices=256,
                 desc="batched 256 slices of 512x512 (slice-parallel), 360 angles, 16 nodes each on a 4-regular graph"),
}
LAM, RHO, SIGMA = 0.02, 2.0, 0.005  # block_7_main_ver3.py:336-337,342
METRIC = "ADMM iters/sec (all nodes)"


def make_graph(cfg):
    from admm_b200 import make_graph as mg
    if cfg["graph"] == "ring":
        G = mg("ring", cfg["V"])
    elif cfg["graph"] == "regular":
        G = mg("regular", cfg["V"], seed=0, degree=4)
    else:
        G = mg("er", cfg["V"], seed=0, p=0.1)
    S = cfg.get("slices", 1)
    if S > 1:   # slice-parallel: S independent copies, node ids slice*V + i
        import networkx as nx
        H = nx.Graph()
        V = cfg["V"]
        for s in range(S):
            H.add_nodes_from(range(s * V, (s + 1) * V))
        for s in range(S):
            for i in range(V):                       # keep each copy's adjacency (neighbour) order
                for j in G.neighbors(i):
                    H.add_edge(s * V + i, s * V + j)
        return H
    return G


def total_nodes(cfg):
    return cfg["V"] * cfg.get("slices", 1)


def node_sigma(cfg, i):
    return SIGMA * (2.0 ** ((i % 4) - 1)) if cfg["hetero"] else SIGMA


def node_prec(cfg):
    if not cfg["hetero"]:
        return None
    s = np.array([node_sigma(cfg, i) for i in range(total_nodes(cfg))])
    return (s ** -2) / np.max(s ** -2)


def noise(cfg, i, shape):
    return (node_sigma(cfg, i) * np.random.default_rng(1234 + i).standard_normal(shape)).astype(np.float32)


def synth_gpu(cfg, device):
    """Synthetic inputs: b_i = A_i x_true + sigma_i eps_i (eps from default_rng(1234+i)), built with the CUDA
    operator (input synthesis only)."""
    import torch
    from admm_b200 import Plan, node_angles, shepp_logan
    N, M = cfg["N"], cfg["M"]
    thetas = node_angles(M, cfg["V"]) * cfg.get("slices", 1)
    V = len(thetas)
    img = shepp_logan(N).astype(np.float32)
    plan = Plan(N, thetas, device=device)
    d_img = torch.from_numpy(np.ascontiguousarray(img.reshape(1, -1))).to(f"cuda:{device}").repeat(V, 1)
    d_s = torch.zeros(plan.A, N, device=f"cuda:{device}")
    plan.forward(d_img, d_s)
    Wl = None
    if cfg["wq"]:
        d_w = torch.zeros(V, N * N, device=f"cuda:{device}")
        plan.colnorm2(d_w)
        Wl = [np.maximum(w, 1e-12) for w in d_w.cpu().numpy().astype(np.float64)]
    s = d_s.cpu().numpy()
    plan.close()
    sinos = []
    for i in range(V):
        a0, a1 = plan.ang_ptr[i], plan.ang_ptr[i + 1]
        sinos.append(s[a0:a1] + noise(cfg, i, (a1 - a0, N)))
    del d_img, d_s
    torch.cuda.empty_cache()
    return thetas, img, sinos, Wl


def q_provider(Wl):
    """block_3_graph_and_precisions.make_precisions' arithmetic-mean provider (:34-39) over the given W vectors."""
    import block_3_graph_and_precisions as b3
    return b3.make_precisions(Wl, q_mode="arithmetic")[1]


def algorithmic_bytes(cfg, G, S, C, nonuniform_q, solves=None):
    """BASELINE.md section 5 per outer iteration (fp32), with the inner work AS EXECUTED: `solves[i]` = average number
    of (S sweeps x C CG iterations) solves node i really ran per iteration (1 + the a14 rule's retries, read back from
    the device's tighten history)."""
    N, M, V = cfg["N"], cfg["M"], total_nodes(cfg)
    from admm_b200 import angle_split
    n = N * N
    per = angle_split(M, cfg["V"]) * cfg.get("slices", 1)
    E = G.number_of_edges()
    tot = 0.0
    for i in range(V):
        deg = G.degree(i)
        m_i = per[i] * N
        k = 1.0 if solves is None else float(solves[i])
        tot += (2 * deg + 6) * n + k * S * (C * (12 * n + 2 * m_i) + 7 * n)
        if nonuniform_q:
            tot += deg * n + k * S * C * n
    tot += 8 * n * E
    return int(4 * tot)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for nme, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the same path on the host cores (bounded sample, scaled to one iteration)
# ------------------------------------------------------------------------------------------------------
def cpu_sample(cfg, S, C, budget_s, steps=1, warmup=0, solves=1):
    from oracle import oracle as O
    N, M = cfg["N"], cfg["M"]
    n = N * N
    G = make_graph(cfg)
    V = G.number_of_nodes()
    edges, ptr, nidx, nedge, nend = O.graph_csr(G)
    E = len(edges)
    thetas = O.node_angles(M, cfg["V"]) * cfg.get("slices", 1)
    img = O.shepp_logan(N)
    cores = O.num_threads()
    prec = node_prec(cfg)
    rng = np.random.default_rng(7)
    L = O.lib()

    def node_work(i, state):
        op, b, x, d, w, atb, zs, ys = state
        cons = np.zeros(n)
        for kk in range(ptr[i], ptr[i + 1]):
            L.orc_accum_cons(n, RHO, None, 1.0, O._p(zs), O._p(ys), O._p(cons))
        deg = int(ptr[i + 1] - ptr[i])
        for _ in range(solves):     # steady state of the a14 rule: every node spends all its solves
            O.x_update(op, 1.0 if prec is None else prec[i], atb + cons, RHO * deg, RHO, LAM, S, C, x, d, w)

    def make_state(i):
        op = O.JosephOperator(N, thetas[i])
        b = op.forward(img) + node_sigma(cfg, i) * np.random.default_rng(1234 + i).standard_normal(op.shape[0])
        return (op, b, np.zeros(n), np.zeros(2 * n), np.zeros(2 * n), op.adjoint(b), 0.01 * rng.standard_normal(n),
                0.01 * rng.standard_normal(n))

    st0 = make_state(0)
    t = time.perf_counter()
    node_work(0, st0)
    t_one = time.perf_counter() - t
    ns = int(max(1, min(V, budget_s * 0.8 / max(t_one, 1e-6))))
    states = [st0] + [make_state(i) for i in range(1, ns)]
    es = int(max(1, min(E, 4)))
    ev = [rng.standard_normal(n) for _ in range(5)]
    sums = np.zeros(5)
    per_step = []
    for it in range(warmup + steps):
        t = time.perf_counter()
        for i in range(ns):
            node_work(i, states[i])
        t_nodes = time.perf_counter() - t
        t = time.perf_counter()
        for _ in range(es):
            L.orc_edge_update(n, O._p(ev[0]), O._p(ev[1]), O._p(ev[2]), O._p(ev[3]), O._p(ev[4]), None, None, None,
                              None, 1.0, O._p(sums))
        t_edges = time.perf_counter() - t
        if it >= warmup:
            per_step.append(t_nodes / ns * V + t_edges / es * E)
    full = float(np.median(per_step))
    sample = (f"ESTIMATE: x-update (rhs assembly + {solves} solve(s) of {S} sweep(s) x {C} CG its, fp64, OpenMP) of {ns} of "
              f"{V} nodes and {es} of {E} edge updates per step, timed and scaled by V/{ns} and E/{es} to one full outer "
              f"iteration")
    return {"value": 1.0 / full, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample,
            "extrapolated": not (ns == V and es == E), "s_per_iteration_est": full}


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    per_step_budget = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    t0 = time.perf_counter()
    res = cpu_sample(cfg, args.tv_sweeps, args.cg_iters, per_step_budget, steps=args.steps, warmup=args.warmup,
                     solves=3 if args.acceptance else 1)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": "iters/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / res["value"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, cfg, make_graph(cfg)),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated")},
            "extrapolated": res["extrapolated"],
            "e2e": {"value": res["value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference CVXPY/SCS/ODL stack is not installable here; this is the oracle port of the same "
                    "path (oracle/admm_oracle.c) on the host cores",
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def inner_accuracy(S, C, acceptance):
    """How far this inner schedule is from a converged inner solve of eq. (1), from the committed oracle-only study
    (tools/inner_schedule_study.py -> profiles/r2_inner_schedule_study.json): max relative error of the primal / dual
    residual traces over 200 iterations, final-x relative L2, PSNR difference."""
    name = f"S{S}C{C}" + ("+accept" if acceptance else "")
    out = {"schedule": name, "source": "profiles/r2_inner_schedule_study.json (oracle, fp64, vs 50 sweeps x 8 CG per solve)"}
    try:
        st = json.load(open(os.path.join(ROOT, "profiles", "r2_inner_schedule_study.json")))
        for c in ("cfg1", "cfg2s", "cfg1s"):
            r = st.get(c, {}).get(name)
            if r:
                out[c] = {"primal_trace_max_rel_err": round(r["primal_trace_max_rel_err"], 4),
                          "dual_trace_max_rel_err": round(r["dual_trace_max_rel_err"], 4),
                          "final_x_rel_l2": round(r["final_x_rel_l2_max"], 5), "psnr_diff_db": round(r["psnr_diff_db_max"], 4),
                          "sweeps_per_node_iter": r["sweeps_per_node_iter"], "cg_per_node_iter": r["cg_per_node_iter"]}
    except Exception:
        pass
    return out


def workload_config(args, cfg, G):
    """The declared workload -- identical in the `ours` and `reference` arms (what was asked for, not what a run found:
    the node map / exchange path a sharded run ended up with are reported beside it under `parallelism`)."""
    return {"workload": f"{args.config}: {cfg['desc']}", "N": cfg["N"], "angles": cfg["M"], "nodes": total_nodes(cfg),
            "graph": cfg["graph"], "edges": G.number_of_edges(), "lam_tv": LAM, "rho": RHO, "tv_mu": RHO,
            "tv_sweeps": args.tv_sweeps, "cg_iters": args.cg_iters, "acceptance": bool(args.acceptance),
            "residual_carry": args.carry,
            "noise_sigma": SIGMA, "gpus": args.gpus, "partition": args.partition, "exchange": args.exchange,
            "inputs_larger_than_L2": cfg["N"] >= 1024 or total_nodes(cfg) * cfg["N"] ** 2 * 4 * 10 > 126e6,
            "stop_test": "disabled in the timed region",
            "inner_accuracy": inner_accuracy(args.tv_sweeps, args.cg_iters, bool(args.acceptance))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg4", choices=sorted(CONFIGS))
    ap.add_argument("--cg-iters", type=int, default=2)
    ap.add_argument("--tv-sweeps", type=int, default=1)
    ap.add_argument("--node-group", type=int, default=0)
    ap.add_argument("--slices", type=int, default=0, help="override the slice count of cfg5")
    ap.add_argument("--no-fuse", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "owner", "p2p", "push", "nccl"])
    ap.add_argument("--partition", default="auto", choices=["auto", "mincut", "contiguous"],
                    help="node -> GPU map when sharded (auto: balanced min-cut for <= 256 nodes)")
    ap.add_argument("--exchange-phases", type=int, default=None,
                    help="NCCL exchange: post the cut-edge transfers in this many pieces per iteration (default 2)")
    ap.add_argument("--acceptance", type=int, default=1, choices=[0, 1],
                    help="1: the reference's accept / tighten-and-retry rule (block_6_ver2:100-176) on the device: up to "
                         "3 solves of tv_sweeps x cg_iters per node and iteration")
    ap.add_argument("--carry", default="iteration", choices=["first_retry", "iteration", "always", "off"],
                    help="CG residual between the solves of one outer iteration: 'iteration' (default) -- the first solve "
                         "rebuilds r = rhs0 + tvterm - Hx with a back-projection, the a14 retry solves take the r the TV pass "
                         "carried along; 'first_retry' -- only the first retry does; 'off' -- every solve rebuilds it; "
                         "'always' -- carried across iterations too.  GPU-vs-oracle trace error at cfg 1 (128^2, 200 its): "
                         "off 0.9e-4, first_retry 1.2e-4, iteration 1.7e-4 (tolerance 1e-3)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="no per-kernel CUDA events in the timed region")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.slices and "slices" in cfg:
        cfg["slices"] = args.slices
        cfg["desc"] = cfg["desc"].replace("256 slices", f"{args.slices} slices")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from admm_b200 import _native as nat
    from admm_b200.solver import ADMMEngine
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    G = make_graph(cfg)
    thetas, img, sinos, Wl = synth_gpu(cfg, local)
    Q = q_provider(Wl) if cfg["wq"] else None
    S, C = args.tv_sweeps, args.cg_iters
    total = args.warmup + 2 * args.steps      # room for the separate profiled pass of the small configs
    eng = ADMMEngine(thetas, sinos, G, cfg["N"], lam_tv=LAM, rho=RHO, Q=Q, Wi_list=Wl, node_prec=node_prec(cfg),
                     tv_sweeps=S, cg_iters=C, phantom_true=img, device=local, dist=dist if world > 1 else None,
                     rank=rank, world=world, node_group=args.node_group or None, fuse_pupdate=not args.no_fuse,
                     max_iters=total, exchange=args.exchange, exchange_phases=args.exchange_phases, partition=args.partition,
                     acceptance=bool(args.acceptance), carry_residual=(False if args.carry == 'off' else args.carry))

    args.exchange_used = ("single-owner (peer memory: x pushed to the edge's owner, v = z' - y' stored back by its edge kernel)"
                          if getattr(eng, "_owner", False) else
                          "push" if (eng.exchange_mode == "p2p" and getattr(eng, "_push", False)) else eng.exchange_mode)
    args.phases_used = eng.phases
    if world > 1:
        from admm_b200.sharding import cut_statistics
        cs = cut_statistics(G, world, eng.node_rank)
        contiguous = eng.node_rank == [(i * world) // len(eng.node_rank) for i in range(len(eng.node_rank))]
        args.partition_used = ("contiguous" if contiguous else "balanced min-cut, angle-balanced") + f" ({cs['cut']} of {cs['edges']} edges cut, max {max(cs['per_rank_ends'])} ends on a rank)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        eng.step()
    barrier()
    prof = not args.no_profile
    # per-kernel CUDA events ride in the timed region when launches are long (>= 1024^2 images: < 0.1 % overhead);
    # for the small, launch-bound configs they would inflate the step, so those get a separate profiled pass
    prof_inline = prof and cfg["N"] >= 1024
    nat.profile_enable(prof_inline)
    if prof_inline:
        nat.profile_read()
    eng.time_exchange = world > 1
    clocks = ClockSampler(local) if rank == 0 else None
    l0 = nat.launch_count() + eng.replayed_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    e0.record()
    for _ in range(args.steps):
        eng.step()
    e1.record()
    barrier()
    tw1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = nat.launch_count() + eng.replayed_launches - l0
    graph_replay = bool(eng.use_graph and eng._graphs and not prof_inline)
    kprof = nat.profile_read() if prof_inline else {}
    nat.profile_enable(False)
    ms_prof = ms
    if prof and not prof_inline:
        # eng.hist holds warmup + steps rows only: the extra pass must not advance the history index past it
        barrier()
        nat.profile_enable(True)
        nat.profile_read()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(args.steps):
            eng.step()
        p1.record()
        barrier()
        ms_prof = p0.elapsed_time(p1)
        kprof = nat.profile_read()
        nat.profile_enable(False)
    clk = clocks.stop(tw0, tw1) if clocks else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=f"cuda:{local}")
        dist.all_reduce(lt)
        launches = int(lt.item())
    ms_per_step = ms / args.steps
    value = 1e3 / ms_per_step
    exch = None
    if world > 1 and getattr(eng, "exchange_events", None):
        tot = sum(a.elapsed_time(b) for a, b in eng.exchange_events[-args.steps:])
        exch = {"pack_tv_localedges_and_exposed_exchange_ms_per_step": round(tot / args.steps, 4), "mode": eng.exchange_mode,
                "cut_edge_ends_this_rank": eng.n_pack, "bytes_out_per_step": eng.n_pack * eng.n * 4}
        if "pack" in kprof and kprof["pack"][1] > 0:   # the push kernel IS the NVLink transfer (posted peer stores)
            gbps = eng.n_pack * eng.n * 4 * kprof["pack"][0] / (kprof["pack"][1] * 1e-3) / 1e9
            exch["nvlink_push_GBps_rank0"] = round(gbps, 1)
            exch["nvlink_peak_GBps_per_direction"] = {"nominal": 900, "measured_peer_copy": 770}
            exch["nvlink_frac_of_measured"] = round(gbps / 770.0, 3)
    pri, dual = eng.residuals()
    eng_iters = eng.k

    # ---- roofline of the dominant kernel (rank 0's launches) -----------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    n = cfg["N"] * cfg["N"]
    Vl, El = eng.V, eng.E
    m_loc = eng.A * cfg["N"]
    nz = 1 if eng.rhoD_vec is not None else 0
    deg_sum = sum(G.degree(g) for g in eng.loc)
    alg = {  # algorithmic bytes per launch (this rank), DESIGN.md "Kernels"
        "fwd": 4 * (Vl * n + m_loc), "fwd_fused": 4 * ((7 if not args.no_fuse else 3) * Vl * n + m_loc),
        # back_resid0: reads x, rhs0, tvterm, writes r (p0 = r is not materialised in the fused CG); cg_update: the x half of
        # the solve's last update (reads x, p, writes x); tv: reads x, w (2), tvterm, r, Hp, writes w' (2), tvterm' and -- in
        # the passes that hand the residual on (2 of 3 with the a14 rule and the carried residual) -- r
        "back_hp": 4 * (m_loc + (2 + nz) * Vl * n),
        "back_resid0": 4 * (m_loc + ((4 if not args.no_fuse else 5) + nz) * Vl * n),
        "cg_update": 4 * (3 if not args.no_fuse else 6) * Vl * n, "p_update": 4 * 3 * Vl * n,
        "tv": int(4 * (9 + (2.0 / 3.0 if (args.carry != "off" and args.acceptance) else 0.0)) * Vl * n),
        "rhs0": 4 * ((2 + nz) * deg_sum + 2 * Vl) * n, "edge": 4 * 8 * n * max(El, 1),
    }
    kernels = []
    for name, (cnt, tms) in sorted(kprof.items(), key=lambda kv: -kv[1][1]):
        k = {"name": name, "launches": cnt, "ms_total": round(tms, 3), "share": round(tms / ms_prof, 4)}
        if name in alg:
            k["alg_GBps"] = round(alg[name] * cnt / (tms * 1e-3) / 1e9, 1)
        kernels.append(k)
    roof = None
    if kernels:
        top = next((k for k in kernels if "alg_GBps" in k), None)
        if top:
            traffic = None
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                traffic = tr.get(args.config, {}).get(top["name"])
            except Exception:
                pass
            ncu = None
            try:   # ncu counters of the same kernel class from the committed capture (profiles/make_profiles.py)
                import glob
                files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_cfg4_*_kernels.json")), reverse=True)
                for f in files:
                    for kk in json.load(open(f)):
                        if kk.get("bench_class") == top["name"] and args.config == "cfg4" and ncu is None:
                            ncu = {"issue_active_pct": kk.get("issue_active_pct"), "dram_pct": kk.get("dram_pct"),
                                   "fma_pipe_pct": kk.get("fma_pipe_pct"), "lsu_pipe_pct": kk.get("lsu_pipe_pct"),
                                   "source": "profiles/" + os.path.basename(f)}
            except Exception:
                pass
            roof = {"kernel": top["name"], "bound": "hbm", "achieved": top["alg_GBps"], "peak": peak, "unit": "GB/s",
                    "ncu": ncu, "kernel_timing": "CUDA events in the timed region" if prof_inline else "CUDA events in a separate profiled pass of the same steps",
                    "frac": round(top["alg_GBps"] / peak, 4), "traffic": traffic, "peak_source": peak_src,
                    "alg_bytes_per_launch": alg[top["name"]], "ms_per_launch": round(top["ms_total"] / top["launches"], 4),
                    "note": "the projector kernels are bound by the FP32 / shared-memory pipes, not by HBM (DESIGN.md section 3): "
                            "the fraction is of the HBM roof as the contract asks, the ncu pipe figures say what binds"}
    solves = None
    if args.acceptance:     # inner work as executed: 1 + retries per node, averaged over the timed iterations
        hh = eng.history()
        th = np.array(hh["tighten_history"][args.warmup:args.warmup + args.steps], dtype=np.float64)
        solves = 1.0 + th.mean(axis=0)
    it_bytes = algorithmic_bytes(cfg, G, S, C, eng.rhoD_vec is not None, solves)
    it_roof = {"alg_bytes_per_iteration": it_bytes, "achieved_GBps": round(it_bytes / (ms_per_step * 1e-3) / 1e9, 1),
               "peak_GBps_all_gpus": peak * world, "frac": round(it_bytes / (ms_per_step * 1e-3) / 1e9 / (peak * world), 4),
               "solves_per_node_per_iteration": None if solves is None else round(float(np.mean(solves)), 3)}

    # ---- e2e: the same solve through the reference-facing block_6 call with HOST inputs / outputs ----------
    e2e = None
    if not args.no_e2e:
        from admm_b200 import RayTransformCUDA
        from block_6_admm_loop_ver2 import decentralized_admm
        del eng
        torch.cuda.empty_cache()
        ops = [RayTransformCUDA(cfg["N"], t, device=local) for t in thetas]
        barrier()
        t0 = time.perf_counter()
        xs, hist, e2 = decentralized_admm(ops, sinos, G, Wl, Q, cfg["N"], lam_tv=LAM, rho=RHO, max_iters=args.steps,
                                          eps_pri=0.0, eps_dual=0.0, verbose=False, phantom_true=img,
                                          cg_iters=C, tv_sweeps=S, node_prec=node_prec(cfg), device=local,
                                          node_group=args.node_group or None, fuse_pupdate=not args.no_fuse,
                                          return_engine=True, exchange=args.exchange, gather="rank0",
                                          exchange_phases=args.exchange_phases, partition=args.partition,
                                          acceptance=bool(args.acceptance), carry_residual=(False if args.carry == 'off' else args.carry))
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = e2.h2d_bytes
        d2h = sum(x.nbytes for x in xs if x is not None) + args.steps * 16 + hist_bytes(hist)
        e2e = {"value": args.steps / dt, "unit": "iters/s", "h2d_bytes_per_step": int(h2d / args.steps),
               "d2h_bytes_per_step": int(d2h / args.steps), "call": "block_6_admm_loop_ver2.decentralized_admm(host numpy "
               "sinograms) -> (host x list, history); includes plan build, uploads, A^T b, per-iteration residual "
               "read-back for the stop test, final x download (to rank 0 when sharded)", "wall_s": round(dt, 3), "iters": len(hist["primal"]),
               "timing_s": {k: round(v, 4) for k, v in hist.get("timing_s", {}).items()}}
        e2.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_sample(cfg, S, C, budget_s=20.0, solves=3 if args.acceptance else 1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, cfg, G),
                "parallelism": (None if world == 1 else {"node_map": getattr(args, "partition_used", None),
                                                        "exchange": getattr(args, "exchange_used", None),
                                                        "phases": getattr(args, "phases_used", 1)}),
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
                "launch_mode": ("CUDA-graph replay of the outer iteration (launch count = kernel nodes replayed)" if graph_replay
                                else "eager launches (per-kernel CUDA events ride in the timed region)"),
                "roofline": roof,
                "iteration_roofline": it_roof, "exchange": exch, "cpu_baseline": cpu, "kernels": kernels,
                "residuals_after": {"primal": pri, "dual": dual, "iterations": eng_iters}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def hist_bytes(hist):
    tot = 0
    for v in hist.values():
        if isinstance(v, list):
            for e in v:
                tot += getattr(e, "nbytes", 8)
    return tot


if __name__ == "__main__":
    main()
