// projector.cuh -- launch parameter blocks of the Joseph projector kernels (K1, K2, K2b).
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace admm {

// ---- K1 forward: strip/segment decomposition ----------------------------------------------------
constexpr int FW = 108;       // interpolation-axis extent of a strip: sqrt((FW+2)^2 + (FL-1)^2) + 1 <= FTPA
constexpr int FL = 64;        // steps per staged slab
constexpr int FSEG_MAX = 256; // max steps per block segment (the plan picks seg <= FSEG_MAX, a multiple of FL)
constexpr int FTPA = 128;     // threads per angle slot (bins handled per round)
constexpr int FAC = 16;       // angles per block chunk
constexpr int FTHREADS = 256; // 2 angle slots
constexpr int FPITCH = FW + 5;  // smem row pitch (odd): [2 zero columns][0..FW)[3 zero columns]
constexpr int FHALO = 2;        // leading zero columns
constexpr int FPITCH_T = 116;   // row pitch of a y-dominant tile landed by TMA: the box (FPITCH_T, FL) at column U0 - FHALO_T
constexpr int FHALO_T = 4;      // leading columns of that tile: a box must start on a 16-byte boundary of the image row
static_assert(FW % 4 == 0 && FHALO_T % 4 == 0 && FHALO_T + FW + 3 <= FPITCH_T, "TMA tile layout");

struct FwdParams {
    const float* img;        // [nodes][N*N] images to project (p or x)
    long long img_stride;    // floats between node images
    const AngleRec* ang;     // [A] angle records of this rank (sinogram row order)
    const int* optr;         // [2][V+1] ranges into oidx per orientation (0: xdom, 1: ydom)
    const int* oidx;         // angle row ids grouped by (orientation, node)
    float* recs;             // [A][nRec][span] partial line integrals
    int* jstart;             // [A][nRec] first bin of each record
    int V;                   // nodes in optr tables
    int node0;               // first node of this launch
    int N, D;
    int nTi, nSeg, span, seg;  // seg = steps per segment
    // mode 1: fused CG direction update p_new = r + beta * p_old (img = p_old), beta = scal[beta_num]/scal[beta_den]
    // mode 2: fully fused CG step   alpha = rr/<p,Hp>; x += alpha p; r' = r - alpha Hp; p' = r' + beta p with
    //         beta = max(0, rr'/rr), rr' = alpha^2 <Hp,Hp> - rr (CG conjugacy); exact <r',r'> -> scal[rr_out]
    const float* r;          // [nodes][N*N] or nullptr (mode 0)
    float* p_out;            // [nodes][N*N]
    double* scal;            // [V][NSCAL] node scalars
    int beta_num, beta_den;  // mode 1 scalar slots; mode 2: beta_den = slot of rr (in), beta_num unused
    int mode;
    const float* hp;         // mode 2: H p_old
    float* x_io;             // mode 2: x (in place)
    float* r_out;            // mode 2: r' (must not alias r)
    float* part;             // mode 2 reduction workspace [nodes][nTi*nSeg]
    unsigned* counter;       // [nodes]
    int rr_out;              // mode 2: slot receiving <r', r'>
    const NodeCtl* ctl;      // masked launches (a14 retry passes): blocks of nodes with ctl[node].active == 0 exit
    // TMA (tma.cuh).  Operand streams as 3-D tensors [nodes][N][N]: 0 img, 1 r, 2 hp, 3 x.  mapY: box (FPITCH_T, FL, 1)
    // = a y-dominant slab tile incl. halo columns; mapX: box (FL, FW, 1) = an x-dominant slab tile.
    // use_tma bit 0: the next slab's boxes are prefetched to L2 while the current slab is sampled;
    //         bit 1: mode-0 y-dominant tiles land in shared memory by cp.async.bulk.tensor (no staging loop).
    int use_tma;
    CUtensorMap mapY[4], mapX[4];
};

struct FwdReduceParams {
    const float* recs;
    const int* jstart;
    const AngleRec* ang;
    float* out;              // [A][D] sinogram rows
    int A0, A1;              // angle row range
    int D, nRec, span;
    // optional fused axpy: acc_out[a][j] (+)= coef(node) * out  (Ax recurrence), node_of_angle lookup
};

// ---- K2 back-projection (gather) with fused epilogues ---------------------------------------------
constexpr int BTX = 32;        // tile rows (ix)
constexpr int BPG = 4;         // float4 pixel groups per thread (32 columns apart)
constexpr int BTY = 32 * BPG;  // tile cols (iy)
constexpr int BTHREADS = 256;  // 4*BPG pixels per thread
constexpr int BAC = 16;        // angles staged per chunk

enum BackMode : int {
    BACK_PLAIN = 0,    // out = A^T (prec * q)
    BACK_HP = 1,       // out = A^T(prec q) + rhoD*v + mu*K^T K v ; scal[dot_slot] = <v, out>
    BACK_RESID0 = 2,   // r = rhs0 + tvterm - H v ; p = r ; scal[dot_slot] = <r, r>
    BACK_COLNORM2 = 3  // out = sum_rays A[r,p]^2  (q unused)
};

struct BackParams {
    const float* q;          // [A][D] sinogram rows (angle-major)
    const AngleRec* ang;
    const int* aptr;         // [V+1] angle row ranges per node
    const float* prec;       // [V] per-node measurement precision P_i (nullptr = 1)
    float* out;              // [nodes][N*N]  (Hp / r / plain)
    long long stride;        // floats between node images (all image arrays share it)
    int node0, N, D, bspan;
    // epilogue inputs
    const float* v;          // p (BACK_HP) or x (BACK_RESID0)
    const float* rvec;       // (unused)
    const float* rhoD_vec;   // [nodes][N*N] or nullptr
    const float* rhoD_s;     // [V] scalar rho*D_i (used when rhoD_vec == nullptr)
    float mu;
    const float* rhs0;       // BACK_RESID0
    const float* tvterm;     // BACK_RESID0
    float* p_out;            // BACK_RESID0: p = r  (nullptr: not materialised -- the CG takes r itself as its first direction)
    // reduction workspace
    float* part;             // [V][nblk] partials
    unsigned* counter;       // [V]
    double* scal;            // [V][NSCAL]
    int dot_slot;
    const NodeCtl* ctl;      // masked launches: skip inactive nodes (nullptr: all nodes)
    // detector windows by TMA (tma.cuh): q as a 2-D tensor [A][D], box = (bpitch, 1); use_tma = 0 keeps the staging loop
    int A_rows;              // angle rows behind q (set by the plan; bounds the tensor map)
    int bpitch;              // shared-memory row pitch of a window (bspan, or bspan rounded up to 32 floats with TMA)
    int use_tma;
    CUtensorMap qmap;
};

constexpr int NSCAL = 16;  // doubles per node in the scalar table
// scalar slots
enum ScalSlot : int {
    S_RR0 = 0, S_RR1 = 1,   // <r,r> ping-pong by CG iteration parity
    S_PHP = 2,              // <p, Hp>
    S_RHP = 3,              // (reserved: <r, Hp>, equal to <p, Hp> for CG directions)
    S_HPHP = 4,             // <Hp, Hp>
    S_TV = 5,               // canonical TV(x)
    S_GN2 = 6,              // |g|^2 stationarity
    S_IMG = 7,              // |x - x_true|^2
    S_MSE = 8,              // |Ax - b|^2
    S_ALPHA = 9,            // CG step length of the solve's last iteration, pending for r (cg_update x_only -> TV pass)
    S_SCRATCH = 15          // sink for reductions whose result is not wanted
};

}  // namespace admm
