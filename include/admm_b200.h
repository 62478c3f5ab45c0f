/*
 * admm_b200.h -- C ABI of libadmm_b200.so: the B200-native (sm_100a) hot path of the decentralized
 * TV-ADMM tomography solve of prsinha1/Distributed-Inverse-Problem-Admm.
 *
 * The reference has no FFI: its boundary is Python module + function signatures.
 * These entry points are what a ctypes binding inside the reference's block_2/3/5/6 modules binds (see
 * INTEGRATION.md); each cites the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void*; all `d_` pointers are
 *     device pointers owned by the caller (PyTorch in the shipped host code); the library allocates only
 *     inside admm_plan_create (geometry tables + projector workspace) and frees in admm_plan_destroy;
 *   - every function returns 0 on success, a negative code otherwise; admm_last_error() gives the text;
 *   - images are float32 [node][ix*N + iy] (axis 0 = x), sinograms float32 [angle row][D] angle-major,
 *     exactly the layouts of `A_dense_list[i] @ x` / `sinograms[i].reshape(-1)`
 *     (block_6_admm_loop_ver2.py:46,145,193);
 *   - no CPU fallback exists: without a CUDA device every compute entry point fails with ADMM_ERR_CUDA.
 */
#ifndef ADMM_B200_H
#define ADMM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ADMM_OK 0
#define ADMM_ERR_ARG (-1)
#define ADMM_ERR_CUDA (-2)

#define ADMM_NSCAL 16 /* doubles per node in the scalar table */
/* scalar-table slots (per node) */
#define ADMM_S_RR0 0
#define ADMM_S_RR1 1
#define ADMM_S_PHP 2
#define ADMM_S_RHP 3
#define ADMM_S_HPHP 4
#define ADMM_S_TV 5
#define ADMM_S_GN2 6
#define ADMM_S_IMG 7
#define ADMM_S_MSE 8

typedef struct admm_plan admm_plan; /* opaque: geometry of the nodes resident on one GPU */

/* Device-resident state of the nodes / edges owned by one rank.  Arrays are [V][...] over the plan's
 * local nodes with `stride` floats between node images. */
typedef struct admm_state {
    float* x;                 /* [V][n]  iterates x_i                      (block_6_ver2:36)            */
    float* r;                 /* [V][n]  CG residual (the final residual of a solve always lands here)   */
    float* r1;                /* [V][n]  CG residual, second buffer of the fully fused CG step (fuse 2)  */
    float* p0;                /* [V][n]  CG direction (ping)                                            */
    float* p1;                /* [V][n]  CG direction (pong)                                            */
    float* hp;                /* [V][n]  H p                                                            */
    float* rhs0;              /* [V][n]  A^T P b + rho sum_j Q_ij (z_ij - y_ij,i)                       */
    float* tvterm;            /* [V][n]  mu K^T (d - w)                                                 */
    const float* atb;         /* [V][n]  A^T P b                                                        */
    float* w0;                /* [V][2][n] TV multiplier (ping)                                         */
    float* w1;                /* [V][2][n] TV multiplier (pong)                                         */
    const float* rhoD_vec;    /* [V][n] rho * sum_j Q_ij, or NULL when Q is uniform                     */
    const float* rhoD_s;      /* [V] rho * deg_i * q (uniform Q)                                        */
    const float* prec;        /* [V] node measurement precisions P_i, or NULL (= 1)                     */
    const float* xtrue;       /* [n] ground truth for img_mse (block_6_ver2:199-206), or NULL           */
    float* q;                 /* [A][D] scratch sinogram A p                                            */
    float* ax;                /* [A][D] A x (kept current by recurrence)                                */
    const float* b;           /* [A][D] measured sinograms b_i (block_6_ver2:46)                        */
    double* scal;             /* [V][ADMM_NSCAL] node scalars                                           */
    float* part;              /* reduction workspace, admm_plan_info(plan, ADMM_INFO_PART_FLOATS) floats*/
    unsigned* counter;        /* reduction workspace, max(V, E) zero-initialised counters               */
    long long stride;         /* floats between node images (>= n)                                      */
    float rho, lam, mu, q_uniform;
    int w_parity;             /* 0: w0 current, 1: w1 current (caller flips it by sweeps&1 after x-updates) */
    int fuse_pupdate;         /* 0: separate kernels; 1: p = r + beta p fused into the forward projector;
                                 2: the whole CG vector update (x, r, p) fused into the forward projector  */
    int defer_tv;             /* 1: admm_x_update skips the LAST sweep's TV pass; the caller runs admm_tv_pass
                                 itself (lets the cut-edge exchange start before the TV kernel)         */
    int reuse_ax;             /* 1: admm_x_update takes A x from `ax` (kept current by the CG recurrence) instead of
                                 re-projecting x for the warm-start residual; set 0 periodically to refresh   */
    struct admm_node_ctl* ctl; /* [V] per-node control words of the a14 accept / tighten rule, or NULL.  When set, the
                                 TV-multiplier parity is per node and device-resident (w_parity is ignored)   */
    int masked;               /* 1: admm_x_update / admm_tv_pass skip the nodes whose ctl[].active is 0 (retry pass) */
    int carry_r;              /* bit mask.  bit 1 (2): the TV pass that ends a solve also hands the CG residual to the next
                                 solve (r += tvterm' - tvterm, <r,r>; p0 = r too unless fuse_pupdate == 2, where the CG takes
                                 the r buffer itself as its first direction).  bit 0 (1): a solve with reuse_r set (and every
                                 sweep after the first) takes that residual instead of rebuilding it with the A^T(P A x)
                                 back-projection; admm_rhs0 then keeps it current too (r += rhs0' - rhs0)              */
    int reuse_r;              /* 1: r, p0 and <r,r> are current for the solve that follows (set 0 periodically, and for
                                 the first solve, to rebuild r = rhs0 + tvterm - H x and stop fp32 drift)             */
    int* iter_dev;            /* device-resident outer iteration counter k, or NULL.  When set (CUDA-graph replay of the
                                 outer iteration: nothing host-side may change between replays), admm_accept derives
                                 eps_target = 2/(k+1)^1.005 from it on the device, and admm_finalize treats d_row as the
                                 base of the history, writes row k at d_row + k*hist_stride and increments k          */
    long long hist_stride;    /* doubles between history rows (with iter_dev)                                          */
    int accept_mode;          /* a14 decision folded into the TV pass that ends a solve (no separate launch): 0 none,
                                 1 decision after the iteration's first solve, 2 after a retry solve (see admm_accept)  */
    int max_tighten;          /* retry cap of that decision (block_6_admm_loop_ver2.py:113: 2)                           */
    double eps_target;        /* its target when iter_dev is NULL                                                       */
    int skip_mse;             /* 1: admm_x_update does not refresh |A x - b|^2 (only the iteration's last solve needs to) */
} admm_state;

/* Per-node control word (block_6_admm_loop_ver2.py:110-113 `accepted`, `tighten_tries`): zero-initialised by the
 * caller, written by admm_accept and by the TV kernel (wpar). */
typedef struct admm_node_ctl { int active, tries, wpar, pad; } admm_node_ctl;

/* One undirected edge (i<j) as seen by this rank; addresses are device pointers as integers. */
typedef struct admm_edge {
    unsigned long long xi, xj, yi, yj, z, ai, aj, Wi, Wj, qij, qji;
    unsigned long long vi, vj; /* single-owner exchange: where v = z' - y' of an end whose node lives on a peer is stored
                                  (the peer's mapped buffer; its next rhs0 reads it as "z" with "y" = 0), or 0 */
} admm_edge;

typedef struct admm_pack_item { unsigned long long x, y, out; } admm_pack_item;

int admm_version(void);            /* 200: round-2 ABI (admm_state grew ctl / masked; admm_accept; history row + tries) */
long long admm_abi_sizeof(int what); /* 0 admm_state, 1 admm_edge, 2 admm_pack_item, 3 admm_node_ctl: lets a binding check
                                        its struct mirrors against the library it loaded */
const char* admm_last_error(void);
int admm_device_count(void); /* 0 when no CUDA device is visible (never throws) */

/* ---- geometry ---------------------------------------------------------------------------------------
 * Replaces odl.uniform_discr / uniform_partition / Parallel2dGeometry / RayTransform construction in
 * _build_parallel_beam_operators (block_2_load_odl_data.py:16-65) and generate_sinogram
 * (Gen_Sino_Partitioned.py:124-134).  `ang_ptr[V+1]` gives each node's angle-row range, `cos32/sin32`
 * are the fp32-rounded trig values of every angle row (computed in fp64 on the host). */
admm_plan* admm_plan_create(int N, int D, double det_w, int V, const int* ang_ptr, const float* cos32,
                            const float* sin32, int device);
/* Dense-matrix plan: the reference's literal `A_dense_list[i]` ndarrays (block_2_load_odl_data.py:68-96; `Ai @ x`,
 * `Ai.T @ r`, block_6_admm_loop_ver2.py:145,193), for tiny problems only (4 m_i n bytes per node).  `row_ptr[V+1]`:
 * matrix-row range of every node; each matrix is (rows x N*N) float32 row-major, uploaded from the host once.  Every
 * entry point below then works unchanged with D = 1 (a "sinogram" is the vector of a node's rows); admm_x_update
 * needs fuse_pupdate = 0. */
admm_plan* admm_plan_create_dense(int N, int V, const int* row_ptr, int device);
int admm_plan_upload_dense(admm_plan* plan, int node, const float* h_A);
void admm_plan_destroy(admm_plan* plan);
#define ADMM_INFO_N 0
#define ADMM_INFO_D 1
#define ADMM_INFO_V 2
#define ADMM_INFO_A 3
#define ADMM_INFO_PART_FLOATS 4   /* floats of `part` needed per unit (node or edge) */
#define ADMM_INFO_FWD_SPAN 5
#define ADMM_INFO_FWD_NREC 6
#define ADMM_INFO_BACK_SPAN 7
#define ADMM_INFO_WS_BYTES 8
long long admm_plan_info(const admm_plan* plan, int what);
/* Tunables.  ADMM_OPT_PACK_BLOCKS = B > 0: admm_pack runs as B blocks that walk all items (for `out` pointers in a
 * peer GPU's memory: a narrow grid keeps NVLink stores in flight beside HBM-bound kernels on another stream);
 * 0 (default): one block row per item. */
#define ADMM_OPT_PACK_BLOCKS 0
/* ADMM_OPT_IMPL: projector discretisation of the plan.  0 (default): Joseph (the hot path).  1: the
 * "skimage-flavoured" rotate-and-sum variant of Gen_Sino_Partitioned.py:133 (`impl='skimage'`): bilinear ray marching on
 * the rotated pixel grid with its exact transpose; plain kernels, admm_x_update then needs fuse_pupdate = 0.  Set it
 * before asking ADMM_INFO_PART_FLOATS. */
#define ADMM_OPT_IMPL 1
int admm_plan_set(admm_plan* plan, int what, long long value);

/* ---- K1 / K2 / K2b: the operator  (device pointers) ---------------------------------------------------
 * admm_forward  : `Ai @ x` / op(x)            block_6_admm_loop_ver2.py:145,193; block_2_load_odl_data.py:149
 * admm_adjoint  : `Ai.T @ r`                  block_6_admm_loop_ver2.py:145   (plain transpose; prec may be NULL)
 * admm_colnorm2 : np.sum(A_i*A_i, axis=0)     block_3_graph_and_precisions.py:22 */
int admm_forward(admm_plan* plan, const float* d_img, long long stride, int node0, int nodes,
                 float* d_sino, void* stream);
int admm_adjoint(admm_plan* plan, const float* d_sino, const float* d_prec, float* d_img, long long stride,
                 int node0, int nodes, void* stream);
int admm_colnorm2(admm_plan* plan, float* d_img, long long stride, int node0, int nodes, void* stream);

/* same, HOST buffers (copies inside; one node) -- what `op(x).asarray()` costs a NumPy caller */
int admm_forward_host(admm_plan* plan, int node, const float* h_img, float* h_sino);
int admm_adjoint_host(admm_plan* plan, int node, const float* h_sino, float* h_img);
int admm_colnorm2_host(admm_plan* plan, int node, float* h_img);

/* ---- K6: rhs0_i = A^T P b_i + rho sum_j Q_ij (z_ij - y_ij,i)   block_6_admm_loop_ver2.py:87-95 ----------
 * nbr_* are device arrays over the CSR neighbour lists (G.neighbors(i) order): addresses of z_ij, of
 * y_ij,i and of Q_ij (0 = uniform). */
int admm_rhs0(admm_plan* plan, const admm_state* st, const int* d_nbr_ptr, const unsigned long long* d_nbr_z,
              const unsigned long long* d_nbr_y, const unsigned long long* d_nbr_q, int node0, int nodes,
              void* stream);

/* ---- K1+K2+K3+K4: node x-update  (replaces build_node_problem + prob.solve, block_5_node_problem.py:6-32,
 * block_6_admm_loop_ver2.py:97-176): `sweeps` x [ `cg_iters` CG iterations on
 * (A^T P A + rho D + mu K^T K) x = rhs0 + mu K^T(d - w) ; d = shrink2(Kx + w, lam/mu) ; w += Kx - d ].
 * Leaves per-node TV(x), |g|^2, |x-x_true|^2, |Ax-b|^2 in the scalar table. */
int admm_x_update(admm_plan* plan, admm_state* st, int node0, int nodes, int sweeps, int cg_iters,
                  void* stream);

/* ---- a14 acceptance   block_6_admm_loop_ver2.py:100-108,155-176 ----------------------------------------------
 * After a solve (admm_x_update incl. its TV pass, which leaves |g_x,i|^2 of :137-149 in the scalar table) decide per
 * node ON THE DEVICE: accepted if |g| <= eps_target or `max_tighten` retries were spent, else tries += 1 and the node
 * stays active for the next admm_x_update with st->masked = 1.  `first` = 1 for the decision after the iteration's
 * first solve (resets active / tries of every node).  No host synchronisation. */
int admm_accept(admm_plan* plan, admm_state* st, int node0, int nodes, double eps_target, int max_tighten,
                int first, void* stream);

/* ---- K5: edges   block_6_admm_loop_ver2.py:210-264 -------------------------------------------------------
 * d_sums[E][5] = |x_i-z'|^2, |x_j-z'|^2, |z'-z|^2, pen_i, pen_j per edge. */
int admm_edge_update(admm_plan* plan, const admm_state* st, const admm_edge* d_edges, int nedges,
                     double* d_sums, void* stream);
/* out = x + y per item; y == 0: out = x (single-owner exchange: the peer that updates the edge gets x itself) */
int admm_pack(admm_plan* plan, const admm_pack_item* d_items, int nitems, void* stream);
/* the same plain copies (y == 0 items) through the copy engines: one cudaMemcpyAsync per item, `h_items` on the HOST */
int admm_push_copy(admm_plan* plan, const admm_pack_item* h_items, int nitems, void* stream);
/* history row [r2, s2, pri_node[Vg], dual_node[Vg], pen[Vg], mse[Vg], tv[Vg], gn2[Vg], img[Vg], tries[Vg]] (doubles;
 * tries = extra solves the a14 rule spent on the node, 0 without a control table).
 * Edge arrays list the rank's local edges first (nedges_local), then its cut edges; flags: bit0 i local, bit1 j
 * local, bit2 this rank owns the edge's dual residual, bit3 / bit4 end i / j belongs to a peer's node whose edge THIS rank
 * updates (its per-node pieces are added to the row here).  nbr_* = incident-edge CSR of the local nodes
 * (G.neighbors order): position of the edge in the edge arrays and which end the node is. */
int admm_finalize(admm_plan* plan, const admm_state* st, const double* d_sums, const int* d_edge_gi,
                  const int* d_edge_gj, const int* d_edge_flags, int nedges, int nedges_local,
                  const int* d_node_gid, const int* d_nbr_ptr, const int* d_nbr_epos, const int* d_nbr_end,
                  int Vg, double* d_row, void* stream);

/* ---- block_4 helpers on device (block_4_tv_helpers.py:17-46): one TV pass without a CG solve ------------
 * used by the drop-in block_4 module; outputs w', tvterm' and TV(x) like the fused K3.  Also the deferred last pass of a
 * solve (st->defer_tv).  with_diag is a bit mask: 1 stationarity / metrics diagnostics; 2 the solve's last CG update left
 * its residual half, r <- r - alpha Hp, to this pass (admm_x_update with cg_iters > 0 always does; alpha sits in
 * scal[S_ALPHA]); 4 that residual is in st->r1, not st->r (fuse_pupdate == 2 and an even cg_iters). */
int admm_tv_pass(admm_plan* plan, admm_state* st, int node0, int nodes, int with_diag, void* stream);

/* ---- peer-memory exchange buffers (CUDA IPC; one process per GPU) -----------------------------------------------
 * admm_ipc_alloc: cudaMalloc + zero + cudaIpcGetMemHandle (64-byte handle, to be shipped to the peers);
 * admm_ipc_open : map a peer's buffer on the current device (lazy peer access over NVLink).  The mapped address
 * goes into admm_edge.ai / .aj so admm_edge_update reads the remote a = x + y in place. */
int admm_ipc_alloc(long long bytes, void** d_ptr, unsigned char* handle64);
int admm_ipc_open(const unsigned char* handle64, void** d_ptr);
int admm_ipc_close(void* d_ptr);
int admm_ipc_free(void* d_ptr);

/* ---- block_4 NumPy helpers on device, HOST fp64 buffers (bit-identical to the reference's float64 NumPy) ----
 * admm_grad2d_host     : _grad_forward_2d_from_vec        block_4_tv_helpers.py:17-23
 * admm_div2d_host      : _div_backward_2d_to_vec          block_4_tv_helpers.py:25-35 (exact_adjoint=0: as shipped,
 *                        border rows/columns sign-flipped; 1: the exact K^T the solver uses)
 * admm_kt_subgrad_host : kt_subgrad_isotropic_tv_from_x   block_4_tv_helpers.py:37-46; h_mag (optional) receives
 *                        |grad x| = edge_map_from_vector  block_4_tv_helpers_with_plot.py:23-46 */
int admm_grad2d_host(int N, const double* h_x, double* h_gx, double* h_gy);
int admm_div2d_host(int N, const double* h_px, const double* h_py, int exact_adjoint, double* h_out);
int admm_kt_subgrad_host(int N, const double* h_x, double eps, int exact_adjoint, double* h_out, double* h_mag);

/* ---- per-pixel graph masks   block_3_graph_and_precisions.py:62-187 (_pixel_mask_knn_then_connect, _pixel_mask_mst,
 * _pixel_mask_chain, _build_all_pixel_masks) ---------------------------------------------------------------------------
 * One thread per pixel.  d_W[V][n]: make_precisions' W vectors (float32); q_ij[p] = 0.5 (W_i + W_j) or, `harmonic`,
 * W_i W_j / (W_i + W_j), floored at 1e-12, in float32 like the reference (:27-39).  strategy 0 knn (k neighbours, plus the
 * maximum-spanning-tree edges when not connected), 1 mst (Kruskal in networkx's edge order), 2 chain (d_perm[n][V]
 * uint8 permutations drawn by the host).  d_keep_bits[V][n]: bit j of word [i][p] = keep[i, j, p]; V <= 32. */
int admm_pixel_masks(int V, long long n, int strategy, int k, int harmonic, const float* d_W,
                     const unsigned char* d_perm, unsigned* d_keep_bits, void* stream);

/* ---- PDHG consensus variant   ADMM_Tomo_Only.py:89-148 ------------------------------------------------
 * The element-wise / stencil pieces of odl.solvers.pdhg (:132-133, :148) for
 *   min_x gamma |x - x_a|^2 + lam_data |A_i x - b_i|^2 + lam_tv |G x|_{2,1},   L = (A_i, G), batched over nodes;
 * A xbar and A^T y1 are admm_forward / admm_adjoint calls.  All buffers are DEVICE fp32; d_sigma, d_tau, d_adj are [V]
 * per-node step sizes and the adjoint factor w_Y / w_X (A* = adj * A^T).  G = forward differences / h, zero padding.
 * admm_pdhg_dual   : y1 <- (y1 + sigma (q - b)) / (1 + sigma / (2 lam_data)),  q = A xbar  ([A][D], global angle rows);
 *                    y2 <- proj_{|.|_2 <= lam_tv}(y2 + sigma G xbar)                         ([V][2][n])
 * admm_pdhg_primal : x <- (x - tau (adj * back + G^T y2) + 2 tau gamma pull) / (1 + 2 tau gamma), back = A^T y1;
 *                    xbar <- x' + theta (x' - x);  d_pull = the shared n-vector x_a (:117-118) or NULL (f = 0, :142)
 * admm_pdhg_normal : out = adj * back + G^T G x, back = A^T (A x): one step of power_method_opnorm (:128, :145)
 * admm_pdhg_combine: x_a = sum_i eta_i x_i / (sum_i eta_i + 1e-8), eta_i = colnorm_i / (|x_i - phantom| + 1e-8) (:100-118)
 * admm_pdhg_sums   : d_out[node][2] (fp64) = { sum (x - phantom)^2 (sum x^2 if phantom NULL), sum (q - b)^2 over the
 *                    node's sinogram rows (0 if q NULL; b NULL: sum q^2) }                    (:134-139 metrics, norms) */
int admm_pdhg_dual(admm_plan* plan, const float* d_xbar, long long stride, float* d_y1, float* d_y2, const float* d_q,
                   const float* d_b, const float* d_sigma, float lam_data, float lam_tv, int node0, int nodes, void* stream);
int admm_pdhg_primal(admm_plan* plan, float* d_x, float* d_xbar, long long stride, const float* d_back, const float* d_y2,
                     const float* d_pull, const float* d_tau, const float* d_adj, float gamma, float theta, int node0,
                     int nodes, void* stream);
int admm_pdhg_normal(admm_plan* plan, const float* d_x, long long stride, const float* d_back, const float* d_adj,
                     float* d_out, int node0, int nodes, void* stream);
int admm_pdhg_combine(admm_plan* plan, const float* d_x, long long stride, const float* d_colnorm, const float* d_phantom,
                      float* d_xa, int nodes, void* stream);
int admm_pdhg_sums(admm_plan* plan, const float* d_x, long long stride, const float* d_phantom, const float* d_q,
                   const float* d_b, double* d_out, int node0, int nodes, void* stream);

/* launches issued by this library since load (the bench's gpu_launches evidence) */
long long admm_launch_count(void);

/* optional CUDA-event profiler: one event pair per launch on the launching stream, summed per kernel class:
 * 0 fwd, 1 fwd_reduce, 2 back_plain, 3 back_hp, 4 back_resid0, 5 colnorm2, 6 tv, 7 cg_update, 8 p_update,
 * 9 sino_axpy, 10 sino_resid, 11 rhs0, 12 edge, 13 pack, 14 finalize, 15 fwd_fused(p-update), 16 accept */
#define ADMM_KC_COUNT 17
int admm_profile_enable(int on);
int admm_profile_read(double* ms, long long* cnt);

#ifdef __cplusplus
}
#endif
#endif /* ADMM_B200_H */
