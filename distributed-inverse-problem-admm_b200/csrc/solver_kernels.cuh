// solver_kernels.cuh -- parameter blocks of K3-K6 and the bookkeeping kernels.
#pragma once
#include "projector.cuh"

namespace admm {

struct TvParams {
    const float* x;        // [nodes][n]
    const float* w_in;     // [nodes][2][n] split-Bregman multiplier (old)
    float* w_out;          // [nodes][2][n] (new; must not alias w_in)
    float* tvterm;         // [nodes][n] in: mu K^T(d-w) used by the last solve; out: the new one
    const float* r;        // [nodes][n] final CG residual (nullptr: skip the stationarity diagnostic)
    float* r_upd;          // [nodes][n] or nullptr: carry the residual to the next solve without a back-projection --
                           //   r += tvterm' - tvterm ; p_out = r ; <r,r> -> scal[S_RR0]   (r = rhs0 + tvterm - H x stays true)
    float* p_out;          // [nodes][n] CG start direction (with r_upd); nullptr: not materialised (the CG starts from r itself)
    const float* xtrue;    // [n] or nullptr
    long long stride;
    int node0, N;
    float lam, mu;
    float* part; unsigned* counter; double* scal;
    NodeCtl* ctl;          // per-node w parity (flipped here) / active mask; nullptr: host parity, all nodes
    int masked;            // 1: skip nodes with ctl[node].active == 0
    // a14 decision folded into this pass (taken by the last block of each node, right after |g|^2 is reduced):
    // 0: none ; 1: decision after the iteration's first solve ; 2: after a retry solve
    int accept, max_tighten;
    double eps_target2;
    const int* iter_dev;
    int strip;             // 1: interior blocks take the row-marching path (ADMM_B200_TVSTRIP=0 keeps the per-row form)
    // the solve's last CG update is split: cg_update_kernel (x_only) did x += alpha p and left alpha in scal[S_ALPHA];
    // this pass applies the other half, r <- r - alpha Hp, on the fly (hp != nullptr), so r is streamed once, not twice
    const float* hp;
};

struct CgParams {
    float* x; float* r; const float* p; const float* hp; float* p_out;
    const float* r_in;     // residual to read (nullptr: r, in place)
    long long stride, n;
    int node0, rr_in, rr_out;
    float* part; unsigned* counter; double* scal;
    const NodeCtl* ctl;    // masked launches: skip inactive nodes (nullptr: all nodes)
    int x_only;            // 1: x += alpha p only; alpha -> scal[S_ALPHA] (the TV pass that follows updates r, see TvParams::hp)
};

struct SinoParams {
    const float* q; float* ax; const float* b;
    const int* anode;      // [A] node of each angle row
    const int* aptr;       // [V+1]
    int A0, A1, D, node0, mode, rr_in;   // mode 0: ax = q ; 1: ax += alpha q
    double* scal;
    const NodeCtl* ctl;    // masked launches: skip inactive nodes (nullptr: all nodes)
};

struct AcceptParams {      // a14 acceptance decision (block_6_admm_loop_ver2.py:155-176), one thread per node
    NodeCtl* ctl; const double* scal;
    int node0, nodes, first, max_tighten;
    double eps_target2;    // eps_target^2, eps_target = 2 / (k+1)^1.005 (:101-103)
    const int* iter_dev;   // device-resident outer iteration counter k (CUDA-graph replay): overrides eps_target2
};

struct RhsParams {
    const float* atb; float* rhs0;
    const int* nbr_ptr;                 // [V+1] CSR over G.neighbors(i) order
    const unsigned long long* nbr_z;    // [nnz] device address of z_ij
    const unsigned long long* nbr_y;    // [nnz] device address of y_ij,i
    const unsigned long long* nbr_q;    // [nnz] device address of Q_ij or 0 (uniform)
    long long stride, n;
    int node0;
    float rho, q_uniform;
    // optional (r_upd != nullptr): carry the CG residual across the outer iteration, r += rhs0' - rhs0 ; p_out = r ;
    // <r,r> -> scal[S_RR0], so the next solve starts without the A^T(P A x) back-projection of its residual
    float* r_upd; float* p_out;
    float* part; unsigned* counter; double* scal;
};

struct EdgeDesc {
    unsigned long long xi, xj;   // x of each end (0: remote end)
    unsigned long long yi, yj;   // scaled duals of the local ends
    unsigned long long z;        // consensus variable (replicated on both owners of a cut edge)
    unsigned long long ai, aj;   // received a = x + y of a remote end
    unsigned long long Wi, Wj;   // per-pixel precisions for the W-weighted fusion (0: midpoint)
    unsigned long long qij, qji; // Q vectors for the penalty value (0: uniform)
    unsigned long long vi, vj;   // single-owner exchange: where to store v = z' - y' of an end whose node lives on a peer
                                 //   (that node's next rhs0 needs exactly q .* v); 0: nothing to store
};

struct EdgeParams {
    const EdgeDesc* edges;
    long long n;
    float q_uniform;
    float* part; unsigned* counter; double* sums;   // sums: [E][5]
};

struct PackDesc { unsigned long long x, y, out; };
struct PackParams { const PackDesc* items; long long n; int nitems; };

struct FinalizeParams {
    const double* sums;       // [E][5]
    const int* edge_gi;       // [E] global node ids
    const int* edge_gj;
    const int* edge_flags;    // bit0: count end i in r2, bit1: end j, bit2: owns the dual residual, bit3 / bit4: end i / j
                              //   belongs to a node of another rank and THIS rank updates the edge (single-owner exchange):
                              //   its per-node pieces are added to the row here
    const double* scal;       // [V][NSCAL]
    const int* node_gid;      // [V]
    const int* nbr_ptr;       // [V+1] incident-edge lists of the local nodes (G.neighbors order)
    const int* nbr_epos;      // [nnz] position of the edge in the sums / flags arrays
    const int* nbr_end;       // [nnz] 0: the node is the edge's min end, 1: max end
    const NodeCtl* ctl;       // [V] or nullptr: tries of the a14 rule -> row block 7
    double* row;              // [2 + 8*Vg]  (iter_dev set: base of the history, row k = row + k * hist_stride)
    int* iter_dev;            // device-resident outer iteration counter, incremented here (or nullptr)
    long long hist_stride;
    int E, E_local, V, Vg;    // edges [E_local, E) are the cut edges
    float rho;
};

cudaError_t launch_forward(const FwdParams& P, int nodes, int max_chunks, const FwdReduceParams& R, cudaStream_t st);
cudaError_t launch_back(int mode, const BackParams& P, int nodes, cudaStream_t st);
cudaError_t launch_tv(const TvParams& P, int nodes, cudaStream_t st);
cudaError_t launch_cg_update(const CgParams& P, int nodes, int nblk, cudaStream_t st);
cudaError_t launch_p_update(const CgParams& P, int nodes, cudaStream_t st);
cudaError_t launch_sino_axpy(const SinoParams& P, cudaStream_t st);
cudaError_t launch_sino_resid(const SinoParams& P, int nodes, cudaStream_t st);
cudaError_t launch_rhs0(const RhsParams& P, int nodes, cudaStream_t st);
cudaError_t launch_edges(const EdgeParams& P, int nedges, int nblk, cudaStream_t st);
cudaError_t launch_pack(const PackParams& P, int nitems, int narrow_blocks, cudaStream_t st);
cudaError_t launch_finalize(const FinalizeParams& P, cudaStream_t st);
cudaError_t launch_accept(const AcceptParams& P, cudaStream_t st);
// dense-matrix operator (dense.cu): same contracts as launch_forward (mode 0 only) / launch_back
cudaError_t launch_dense_forward(const float* A, const int* anode, const FwdParams& P, int nodes, const FwdReduceParams& R,
                                 cudaStream_t st);
cudaError_t launch_dense_back(const float* A, int mode, const BackParams& P, int nodes, cudaStream_t st);
// rotate-and-sum ("skimage-flavoured") projector variant (rotsum.cu): cs = (cos, sin) per angle row
cudaError_t launch_rs_forward(const float2* cs, const int* anode, double det_w, const FwdParams& P, int nodes,
                              const FwdReduceParams& R, cudaStream_t st);
cudaError_t launch_rs_back(const float2* cs, double det_w, int mode, const BackParams& P, int nodes, cudaStream_t st);
// PDHG consensus variant (pdhg.cu): element-wise / stencil pieces of one PDHG step, batched over nodes
cudaError_t launch_pdhg_dual(float* y1, float* y2, const float* xbar, long long stride, const float* q, const float* b,
                             const int* anode, const float* sigma, float lam_d, float lam_t, int N, int D, int A0, int A1,
                             int node0, int nodes, cudaStream_t st);
cudaError_t launch_pdhg_primal(float* x, float* xbar, long long stride, const float* back, const float* y2,
                               const float* pull, const float* tau, const float* adj, float gamma, float theta, int N,
                               int node0, int nodes, cudaStream_t st);
cudaError_t launch_pdhg_normal(float* out, const float* x, long long stride, const float* back, const float* adj, int N,
                               int node0, int nodes, cudaStream_t st);
cudaError_t launch_pdhg_combine(float* xa, const float* x, long long stride, const float* cn, const float* phantom,
                                long long n, int nodes, cudaStream_t st);
cudaError_t launch_pdhg_sums(double* out, const float* x, long long stride, const float* phantom, const float* q,
                             const float* b, const int* aptr, long long n, int D, int node0, int nodes, cudaStream_t st);

}  // namespace admm
