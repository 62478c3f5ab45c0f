// api.cu -- extern "C" boundary of libadmm_b200.so (see include/admm_b200.h) and the native host-side
// drivers that sequence the kernels of one node x-update / one edge sweep.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/admm_b200.h"
#include "solver_kernels.cuh"

using namespace admm;

static_assert(ADMM_NSCAL == NSCAL, "scalar table width");
static_assert(ADMM_KC_COUNT == admm::KC_COUNT, "kernel class count");
static_assert(sizeof(admm_edge) == sizeof(EdgeDesc), "edge descriptor layout");
static_assert(sizeof(admm_pack_item) == sizeof(PackDesc), "pack descriptor layout");
static_assert(sizeof(admm_node_ctl) == sizeof(NodeCtl), "node control layout");

namespace admm {
std::atomic<long long> g_launch_count{0};
std::atomic<bool> g_prof_on{false};
struct ProfRec { int kc; cudaEvent_t e0, e1; };
static std::mutex g_prof_mu;              // the record list and the event pool may be touched from several host threads
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_pool;
static cudaEvent_t prof_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
void prof_mark(int kc, cudaStream_t st, bool begin) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (begin) {
        ProfRec r{kc, prof_event(), prof_event()};
        cudaEventRecord(r.e0, st);
        g_prof.push_back(r);
    } else if (!g_prof.empty()) {
        cudaEventRecord(g_prof.back().e1, st);
    }
}
}  // namespace admm

extern "C" int admm_profile_enable(int on) {
    admm::g_prof_on = (on != 0);
    return ADMM_OK;
}
// Synchronises the device, adds the elapsed ms and launch counts of every recorded launch to ms[kc] / cnt[kc]
// (arrays of ADMM_KC_COUNT) and clears the record list.
extern "C" int admm_profile_read(double* ms, long long* cnt) {
    if (!ms || !cnt) return ADMM_ERR_ARG;
    if (cudaDeviceSynchronize() != cudaSuccess) return ADMM_ERR_CUDA;
    std::lock_guard<std::mutex> lk(admm::g_prof_mu);
    for (auto& r : admm::g_prof) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) { ms[r.kc] += t; cnt[r.kc] += 1; }
        admm::g_pool.push_back(r.e0);
        admm::g_pool.push_back(r.e1);
    }
    admm::g_prof.clear();
    return ADMM_OK;
}

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(ADMM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
    } while (0)

struct admm_plan {
    int N = 0, D = 0, V = 0, A = 0, device = 0;
    double det_w = 2.0;
    std::vector<int> aptr;           // [V+1]
    std::vector<AngleRec> recs_h;    // [A]
    AngleRec* d_ang = nullptr;
    int* d_optr = nullptr;           // [2][V+1]
    int* d_oidx = nullptr;           // [A]
    int* d_aptr = nullptr;           // [V+1]
    int* d_anode = nullptr;          // [A]
    float* d_recs = nullptr;         // [A][nRec][span]
    int* d_jstart = nullptr;         // [A][nRec]
    int nTi = 0, nSeg = 0, span = 0, bspan = 0, max_chunks = 1, seg = FSEG_MAX;
    long long ws_bytes = 0;
    // scratch of the host-buffer entry points
    float* d_himg = nullptr;
    float* d_hsino = nullptr;
    int hsino_rows = 0;
    int pack_blocks = 0;             // ADMM_OPT_PACK_BLOCKS (0: one block row per item)
    // dense-matrix plans (admm_plan_create_dense): A = total matrix rows, D = 1, d_dense = [A][N*N] row-major
    bool dense = false;
    float* d_dense = nullptr;
    // projector discretisation: 0 Joseph (the hot path), 1 rotate-and-sum bilinear ("skimage-flavoured", (f)-2)
    int impl = 0;
    float2* d_cs = nullptr;          // [A] fp32 (cos, sin) per angle row
};

static int check_nodes(const admm_plan* p, int node0, int nodes);

// the operator pair of a plan: strip projector / tile back-projector, or the dense matvecs
static cudaError_t plan_forward(const admm_plan* p, const FwdParams& F, int nodes, const FwdReduceParams& R, cudaStream_t st) {
    if (p->dense) return launch_dense_forward(p->d_dense, p->d_anode, F, nodes, R, st);
    if (p->impl == 1) return launch_rs_forward(p->d_cs, p->d_anode, p->det_w, F, nodes, R, st);
    return launch_forward(F, nodes, p->max_chunks, R, st);
}
static cudaError_t plan_back(const admm_plan* p, int mode, const BackParams& B, int nodes, cudaStream_t st) {
    if (p->dense) return launch_dense_back(p->d_dense, mode, B, nodes, st);
    if (p->impl == 1) return launch_rs_back(p->d_cs, p->det_w, mode, B, nodes, st);
    return launch_back(mode, B, nodes, st);
}

extern "C" int admm_version(void) { return 200; }
extern "C" long long admm_abi_sizeof(int what) {
    switch (what) {
        case 0: return (long long)sizeof(admm_state);
        case 1: return (long long)sizeof(admm_edge);
        case 2: return (long long)sizeof(admm_pack_item);
        case 3: return (long long)sizeof(admm_node_ctl);
    }
    return -1;
}
extern "C" const char* admm_last_error(void) { return g_err.c_str(); }
extern "C" long long admm_launch_count(void) { return admm::g_launch_count.load(); }
extern "C" int admm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static void build_angle_rec(int N, int D, double det_w, float c32, float s32, AngleRec* r) {
    const double c = (double)c32, s = (double)s32;
    const double h = 2.0 / N, ds = det_w / D;
    const bool xdom = std::fabs(c) > std::fabs(s);  // tie -> sin-dominant branch
    r->ct = c * h / ds;
    r->st = s * h / ds;
    const double M = xdom ? r->ct : r->st, m = xdom ? r->st : r->ct;
    r->inv_major = (float)(1.0 / M);
    r->slope = (float)(m / M);
    r->wgt = (float)(h / std::fabs(xdom ? c : s));
    r->inv_om = (float)(1.0 / std::fabs(M));
    r->xdom = xdom ? 1 : 0;
    r->inv_slope = (std::fabs(m / M) > 1e-6) ? (float)(M / m) : 0.0f;
}

extern "C" admm_plan* admm_plan_create(int N, int D, double det_w, int V, const int* ang_ptr, const float* cos32,
                                       const float* sin32, int device) {
    if (N < 2 || D < 1 || V < 1 || !ang_ptr || !cos32 || !sin32 || det_w <= 0.0) {
        fail(ADMM_ERR_ARG, "admm_plan_create: bad argument");
        return nullptr;
    }
    if (admm_device_count() <= device) {
        fail(ADMM_ERR_CUDA, "admm_plan_create: no CUDA device (this library has no CPU fallback)");
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        fail(ADMM_ERR_CUDA, "cudaSetDevice failed");
        return nullptr;
    }
    admm_plan* p = new admm_plan();
    p->N = N; p->D = D; p->V = V; p->det_w = det_w; p->device = device;
    p->aptr.assign(ang_ptr, ang_ptr + V + 1);
    p->A = ang_ptr[V];
    const int A = p->A;
    p->recs_h.resize(A > 0 ? A : 1);
    std::vector<int> optr(2 * (V + 1), 0), oidx(A > 0 ? A : 1, 0), anode(A > 0 ? A : 1, 0);
    double ext_f = 0.0, ext_b = 0.0;
    p->nTi = (N + FW - 1) / FW;
    {   // segment length (a multiple of the slab length FL): long segments mean fewer records and less per-block
        // setup, short ones a finer tail.  Model: time ~ waves(blocks) * (seg + c0) with 4 resident blocks per SM
        // (__launch_bounds__(FTHREADS, 4)) and c0 ~ the per-block setup + record write-out in units of steps.
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        const long long slots = 4LL * std::max(sms, 1);
        const int c0 = 24;
        long long best_cost = -1;
        int best_seg = FL;
        for (int seg = FL; seg <= FSEG_MAX; seg += FL) {
            const long long nseg = (N + seg - 1) / seg;
            const long long blocks = (long long)std::max(V, 1) * p->nTi * nseg;
            const long long waves = (blocks + slots - 1) / slots;
            const long long cost = waves * (std::min(seg, N) + c0);
            if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best_seg = seg; }   // ties: the longer segment
            if (seg >= N) break;
        }
        p->seg = best_seg;
    }
    const int Wm = std::min(FW, N), Km = std::min(p->seg, N);
    for (int a = 0; a < A; ++a) {
        build_angle_rec(N, D, det_w, cos32[a], sin32[a], &p->recs_h[a]);
        const AngleRec& r = p->recs_h[a];
        const double M = std::fabs(r.xdom ? r.ct : r.st), m = std::fabs(r.xdom ? r.st : r.ct);
        ext_f = std::max(ext_f, (Wm - 1) * M + (Km - 1) * m + 2.0 * M);
        ext_b = std::max(ext_b, (BTX - 1) * std::fabs(r.ct) + (BTY - 1) * std::fabs(r.st) + 2.0 * std::max(M, 1.0));
    }
    p->span = (int)std::ceil(ext_f) + 4;
    p->bspan = (int)std::ceil(ext_b) + 6;
    p->nSeg = (N + p->seg - 1) / p->seg;
    // orientation lists: [x-dominant angles of node 0..V-1][y-dominant angles of node 0..V-1]
    int pos = 0;
    p->max_chunks = 1;
    for (int o = 0; o < 2; ++o) {
        for (int v = 0; v < V; ++v) {
            optr[o * (V + 1) + v] = pos;
            for (int a = ang_ptr[v]; a < ang_ptr[v + 1]; ++a) {
                if ((p->recs_h[a].xdom == 1) == (o == 0)) oidx[pos++] = a;
                anode[a] = v;
            }
            const int cnt = pos - optr[o * (V + 1) + v];
            p->max_chunks = std::max(p->max_chunks, (cnt + FAC - 1) / FAC);
        }
        optr[o * (V + 1) + V] = pos;
    }
    const size_t nRec = (size_t)p->nTi * p->nSeg;
    const size_t rec_bytes = (size_t)std::max(A, 1) * nRec * p->span * sizeof(float);
    const size_t js_bytes = (size_t)std::max(A, 1) * nRec * sizeof(int);
    bool ok = true;
    ok &= cudaMalloc(&p->d_ang, sizeof(AngleRec) * std::max(A, 1)) == cudaSuccess;
    ok &= cudaMalloc(&p->d_optr, sizeof(int) * optr.size()) == cudaSuccess;
    ok &= cudaMalloc(&p->d_oidx, sizeof(int) * oidx.size()) == cudaSuccess;
    ok &= cudaMalloc(&p->d_aptr, sizeof(int) * (V + 1)) == cudaSuccess;
    ok &= cudaMalloc(&p->d_anode, sizeof(int) * anode.size()) == cudaSuccess;
    ok &= cudaMalloc(&p->d_recs, rec_bytes) == cudaSuccess;
    ok &= cudaMalloc(&p->d_jstart, js_bytes) == cudaSuccess;
    ok &= cudaMalloc(&p->d_cs, sizeof(float2) * std::max(A, 1)) == cudaSuccess;
    if (ok) {
        std::vector<float2> csh(std::max(A, 1));
        for (int a = 0; a < A; ++a) csh[a] = make_float2(cos32[a], sin32[a]);
        ok &= cudaMemcpy(p->d_cs, csh.data(), sizeof(float2) * std::max(A, 1), cudaMemcpyHostToDevice) == cudaSuccess;
        ok &= cudaMemcpy(p->d_ang, p->recs_h.data(), sizeof(AngleRec) * std::max(A, 1), cudaMemcpyHostToDevice) == cudaSuccess;
        ok &= cudaMemcpy(p->d_optr, optr.data(), sizeof(int) * optr.size(), cudaMemcpyHostToDevice) == cudaSuccess;
        ok &= cudaMemcpy(p->d_oidx, oidx.data(), sizeof(int) * oidx.size(), cudaMemcpyHostToDevice) == cudaSuccess;
        ok &= cudaMemcpy(p->d_aptr, ang_ptr, sizeof(int) * (V + 1), cudaMemcpyHostToDevice) == cudaSuccess;
        ok &= cudaMemcpy(p->d_anode, anode.data(), sizeof(int) * anode.size(), cudaMemcpyHostToDevice) == cudaSuccess;
        ok &= cudaMemset(p->d_jstart, 0, js_bytes) == cudaSuccess;
        ok &= cudaMemset(p->d_recs, 0, rec_bytes) == cudaSuccess;
    }
    if (!ok) {
        fail(ADMM_ERR_CUDA, std::string("admm_plan_create: ") + cudaGetErrorString(cudaGetLastError()));
        admm_plan_destroy(p);
        return nullptr;
    }
    p->ws_bytes = (long long)(rec_bytes + js_bytes);
    return p;
}

// Dense-matrix plan: node i's operator is an explicit (row_ptr[i+1]-row_ptr[i]) x N*N float32 matrix, uploaded with
// admm_plan_upload_dense.  "Sinogram" arrays of such a plan are [rows][1].
extern "C" admm_plan* admm_plan_create_dense(int N, int V, const int* row_ptr, int device) {
    if (N < 2 || V < 1 || !row_ptr) {
        fail(ADMM_ERR_ARG, "admm_plan_create_dense: bad argument");
        return nullptr;
    }
    if (admm_device_count() <= device) {
        fail(ADMM_ERR_CUDA, "admm_plan_create_dense: no CUDA device (this library has no CPU fallback)");
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        fail(ADMM_ERR_CUDA, "cudaSetDevice failed");
        return nullptr;
    }
    admm_plan* p = new admm_plan();
    p->N = N; p->D = 1; p->V = V; p->device = device; p->dense = true;
    p->aptr.assign(row_ptr, row_ptr + V + 1);
    p->A = row_ptr[V];
    const int A = std::max(p->A, 1);
    std::vector<int> anode(A, 0);
    for (int v = 0; v < V; ++v)
        for (int a = row_ptr[v]; a < row_ptr[v + 1]; ++a) anode[a] = v;
    const size_t bytes = (size_t)A * N * N * sizeof(float);
    bool ok = cudaMalloc(&p->d_aptr, sizeof(int) * (V + 1)) == cudaSuccess;
    ok = ok && cudaMalloc(&p->d_anode, sizeof(int) * A) == cudaSuccess;
    ok = ok && cudaMalloc(&p->d_dense, bytes) == cudaSuccess;
    ok = ok && cudaMemcpy(p->d_aptr, row_ptr, sizeof(int) * (V + 1), cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(p->d_anode, anode.data(), sizeof(int) * A, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        fail(ADMM_ERR_CUDA, std::string("admm_plan_create_dense: ") + cudaGetErrorString(cudaGetLastError()));
        admm_plan_destroy(p);
        return nullptr;
    }
    p->ws_bytes = (long long)bytes;
    return p;
}

extern "C" int admm_plan_upload_dense(admm_plan* p, int node, const float* h_A) {
    if (int e = check_nodes(p, node, 1)) return e;
    if (!p->dense || !h_A) return fail(ADMM_ERR_ARG, "admm_plan_upload_dense: not a dense plan / null matrix");
    CK(cudaSetDevice(p->device));
    const size_t n = (size_t)p->N * p->N;
    const int r0 = p->aptr[node], r1 = p->aptr[node + 1];
    CK(cudaMemcpy(p->d_dense + (size_t)r0 * n, h_A, (size_t)(r1 - r0) * n * sizeof(float), cudaMemcpyHostToDevice));
    return ADMM_OK;
}

extern "C" void admm_plan_destroy(admm_plan* p) {
    if (!p) return;
    cudaFree(p->d_ang); cudaFree(p->d_optr); cudaFree(p->d_oidx); cudaFree(p->d_aptr); cudaFree(p->d_anode);
    cudaFree(p->d_recs); cudaFree(p->d_jstart); cudaFree(p->d_himg); cudaFree(p->d_hsino); cudaFree(p->d_dense);
    cudaFree(p->d_cs);
    delete p;
}

extern "C" int admm_plan_set(admm_plan* p, int what, long long value) {
    if (!p) return fail(ADMM_ERR_ARG, "admm_plan_set: null plan");
    switch (what) {
        case ADMM_OPT_PACK_BLOCKS:
            if (value < 0 || value > 65535) return fail(ADMM_ERR_ARG, "admm_plan_set: pack blocks out of range");
            p->pack_blocks = (int)value;
            return ADMM_OK;
        case ADMM_OPT_IMPL:
            if (p->dense || (value != 0 && value != 1)) return fail(ADMM_ERR_ARG, "admm_plan_set: bad projector variant");
            p->impl = (int)value;
            return ADMM_OK;
        default: return fail(ADMM_ERR_ARG, "admm_plan_set: unknown option");
    }
}

extern "C" long long admm_plan_info(const admm_plan* p, int what) {
    if (!p) return -1;
    switch (what) {
        case ADMM_INFO_N: return p->N;
        case ADMM_INFO_D: return p->D;
        case ADMM_INFO_V: return p->V;
        case ADMM_INFO_A: return p->A;
        case ADMM_INFO_PART_FLOATS: {
            const long long tiles = (long long)((p->N + 31) / 32) * ((p->N + 31) / 32);       // back-projector grid
            const long long tvblk = (long long)((p->N + 127) / 128) * ((p->N + 7) / 8);       // TV grid
            const long long dblk = (p->dense || p->impl != 0) ? 3 * (((long long)p->N * p->N + 255) / 256) : 0;   // per-pixel back-projector grids
            return std::max(std::max(std::max(3 * tiles, 4 * tvblk), 5LL * 4096), dblk);
        }
        case ADMM_INFO_FWD_SPAN: return p->span;
        case ADMM_INFO_FWD_NREC: return (long long)p->nTi * p->nSeg;
        case ADMM_INFO_BACK_SPAN: return p->bspan;
        case ADMM_INFO_WS_BYTES: return p->ws_bytes;
    }
    return -1;
}

static int check_nodes(const admm_plan* p, int node0, int nodes) {
    if (!p) return fail(ADMM_ERR_ARG, "null plan");
    if (node0 < 0 || nodes < 1 || node0 + nodes > p->V) return fail(ADMM_ERR_ARG, "node range outside the plan");
    return ADMM_OK;
}

static FwdParams make_fwd(const admm_plan* p, const float* img, long long stride, int node0) {
    FwdParams P{};
    P.img = img; P.img_stride = stride; P.ang = p->d_ang; P.optr = p->d_optr; P.oidx = p->d_oidx;
    P.recs = p->d_recs; P.jstart = p->d_jstart; P.V = p->V; P.node0 = node0; P.N = p->N; P.D = p->D;
    P.nTi = p->nTi; P.nSeg = p->nSeg; P.span = p->span; P.seg = p->seg;
    P.r = nullptr; P.p_out = nullptr; P.scal = nullptr; P.beta_num = 0; P.beta_den = 0; P.mode = 0;
    P.hp = nullptr; P.x_io = nullptr; P.r_out = nullptr; P.part = nullptr; P.counter = nullptr; P.rr_out = 0;
    return P;
}

static FwdReduceParams make_red(const admm_plan* p, float* sino, int node0, int nodes) {
    FwdReduceParams R{};
    R.recs = p->d_recs; R.jstart = p->d_jstart; R.ang = p->d_ang; R.out = sino;
    R.A0 = p->aptr[node0]; R.A1 = p->aptr[node0 + nodes]; R.D = p->D; R.nRec = p->nTi * p->nSeg; R.span = p->span;
    return R;
}

extern "C" int admm_forward(admm_plan* p, const float* d_img, long long stride, int node0, int nodes,
                            float* d_sino, void* stream) {
    if (int e = check_nodes(p, node0, nodes)) return e;
    if (!d_img || !d_sino) return fail(ADMM_ERR_ARG, "admm_forward: null buffer");
    FwdParams P = make_fwd(p, d_img, stride, node0);
    CK(plan_forward(p, P, nodes, make_red(p, d_sino, node0, nodes), (cudaStream_t)stream));
    return ADMM_OK;
}

static BackParams make_back(const admm_plan* p, const float* q, const float* prec, float* out, long long stride,
                            int node0) {
    BackParams B{};
    B.q = q; B.ang = p->d_ang; B.aptr = p->d_aptr; B.prec = prec; B.out = out; B.stride = stride;
    B.node0 = node0; B.N = p->N; B.D = p->D; B.bspan = p->bspan; B.A_rows = p->A;
    return B;
}

extern "C" int admm_adjoint(admm_plan* p, const float* d_sino, const float* d_prec, float* d_img, long long stride,
                            int node0, int nodes, void* stream) {
    if (int e = check_nodes(p, node0, nodes)) return e;
    if (!d_img || !d_sino) return fail(ADMM_ERR_ARG, "admm_adjoint: null buffer");
    CK(plan_back(p, BACK_PLAIN, make_back(p, d_sino, d_prec, d_img, stride, node0), nodes, (cudaStream_t)stream));
    return ADMM_OK;
}

extern "C" int admm_colnorm2(admm_plan* p, float* d_img, long long stride, int node0, int nodes, void* stream) {
    if (int e = check_nodes(p, node0, nodes)) return e;
    if (!d_img) return fail(ADMM_ERR_ARG, "admm_colnorm2: null buffer");
    CK(plan_back(p, BACK_COLNORM2, make_back(p, nullptr, nullptr, d_img, stride, node0), nodes, (cudaStream_t)stream));
    return ADMM_OK;
}

static int host_scratch(admm_plan* p, int rows) {
    const size_t n = (size_t)p->N * p->N;
    if (!p->d_himg) CK(cudaMalloc(&p->d_himg, n * sizeof(float)));
    if (p->hsino_rows < rows) {
        cudaFree(p->d_hsino);
        p->d_hsino = nullptr;
        CK(cudaMalloc(&p->d_hsino, (size_t)rows * p->D * sizeof(float)));
        p->hsino_rows = rows;
    }
    return ADMM_OK;
}

extern "C" int admm_forward_host(admm_plan* p, int node, const float* h_img, float* h_sino) {
    if (int e = check_nodes(p, node, 1)) return e;
    if (!h_img || !h_sino) return fail(ADMM_ERR_ARG, "admm_forward_host: null buffer");
    CK(cudaSetDevice(p->device));
    if (int e = host_scratch(p, p->A)) return e;
    const size_t n = (size_t)p->N * p->N;
    const int a0 = p->aptr[node], a1 = p->aptr[node + 1];
    CK(cudaMemcpy(p->d_himg, h_img, n * sizeof(float), cudaMemcpyHostToDevice));
    // d_hsino is indexed by global angle row like every sinogram array
    if (int e = admm_forward(p, p->d_himg, (long long)n, node, 1, p->d_hsino, nullptr)) return e;
    CK(cudaMemcpy(h_sino, p->d_hsino + (size_t)a0 * p->D, (size_t)(a1 - a0) * p->D * sizeof(float), cudaMemcpyDeviceToHost));
    return ADMM_OK;
}

extern "C" int admm_adjoint_host(admm_plan* p, int node, const float* h_sino, float* h_img) {
    if (int e = check_nodes(p, node, 1)) return e;
    if (!h_img || !h_sino) return fail(ADMM_ERR_ARG, "admm_adjoint_host: null buffer");
    CK(cudaSetDevice(p->device));
    if (int e = host_scratch(p, p->A)) return e;
    const size_t n = (size_t)p->N * p->N;
    const int a0 = p->aptr[node], a1 = p->aptr[node + 1];
    CK(cudaMemcpy(p->d_hsino + (size_t)a0 * p->D, h_sino, (size_t)(a1 - a0) * p->D * sizeof(float), cudaMemcpyHostToDevice));
    if (int e = admm_adjoint(p, p->d_hsino, nullptr, p->d_himg, (long long)n, node, 1, nullptr)) return e;
    CK(cudaMemcpy(h_img, p->d_himg, n * sizeof(float), cudaMemcpyDeviceToHost));
    return ADMM_OK;
}

extern "C" int admm_colnorm2_host(admm_plan* p, int node, float* h_img) {
    if (int e = check_nodes(p, node, 1)) return e;
    if (!h_img) return fail(ADMM_ERR_ARG, "admm_colnorm2_host: null buffer");
    CK(cudaSetDevice(p->device));
    if (int e = host_scratch(p, p->A)) return e;
    const size_t n = (size_t)p->N * p->N;
    if (int e = admm_colnorm2(p, p->d_himg, (long long)n, node, 1, nullptr)) return e;
    CK(cudaMemcpy(h_img, p->d_himg, n * sizeof(float), cudaMemcpyDeviceToHost));
    return ADMM_OK;
}

extern "C" int admm_rhs0(admm_plan* p, const admm_state* s, const int* d_nbr_ptr, const unsigned long long* d_nbr_z,
                         const unsigned long long* d_nbr_y, const unsigned long long* d_nbr_q, int node0, int nodes,
                         void* stream) {
    if (int e = check_nodes(p, node0, nodes)) return e;
    if (!s || !d_nbr_ptr) return fail(ADMM_ERR_ARG, "admm_rhs0: null argument");
    RhsParams R{};
    const long long off = (long long)node0 * s->stride;
    R.atb = s->atb + off; R.rhs0 = s->rhs0 + off; R.nbr_ptr = d_nbr_ptr; R.nbr_z = d_nbr_z; R.nbr_y = d_nbr_y;
    R.nbr_q = d_nbr_q; R.stride = s->stride; R.n = (long long)p->N * p->N; R.node0 = node0; R.rho = s->rho;
    R.q_uniform = s->q_uniform;
    if ((s->carry_r & 1) && s->reuse_r) {   // the residual rides along: r += rhs0' - rhs0 ; p0 = r ; <r,r> -> S_RR0
        R.r_upd = s->r + off; R.p_out = (s->fuse_pupdate == 2) ? nullptr : s->p0 + off;
        R.part = s->part + (long long)node0 * admm_plan_info(p, ADMM_INFO_PART_FLOATS);
        R.counter = s->counter + node0; R.scal = s->scal;
    }
    CK(launch_rhs0(R, nodes, (cudaStream_t)stream));
    return ADMM_OK;
}

static TvParams make_tv(const admm_plan* p, const admm_state* s, int node0, bool diag, int parity, bool hand_on) {
    TvParams T{};
    const long long off = (long long)node0 * s->stride;
    if (s->ctl) parity = 0;   // the per-node parity lives in ctl[node].wpar and is applied (and flipped) by the kernel
    const float* win = parity ? s->w1 : s->w0;
    float* wout = parity ? s->w0 : s->w1;
    T.x = s->x + off; T.w_in = win + 2 * off; T.w_out = wout + 2 * off; T.tvterm = s->tvterm + off;
    T.r = diag ? s->r + off : nullptr; T.xtrue = s->xtrue;
    // carry_r bit 1: this pass hands r = rhs0 + tvterm' - H x to the next solve (bit 0, in admm_x_update: the solve takes
    // the carried residual).  With the fully fused CG the start direction p0 = r is not materialised: it IS the r buffer
    T.hp = nullptr;   // set by the caller when the solve's last CG update left its r half pending
    if (hand_on) { T.r_upd = s->r + off; T.p_out = (s->fuse_pupdate == 2) ? nullptr : s->p0 + off; }
    T.stride = s->stride; T.node0 = node0; T.N = p->N;
    T.lam = s->lam; T.mu = s->mu;
    T.part = s->part + (long long)node0 * admm_plan_info(p, ADMM_INFO_PART_FLOATS);
    T.counter = s->counter + node0; T.scal = s->scal;
    T.ctl = reinterpret_cast<NodeCtl*>(s->ctl); T.masked = (s->ctl && s->masked) ? 1 : 0;
    T.accept = 0;   // set by the caller for the pass that ends a solve
    T.max_tighten = s->max_tighten; T.eps_target2 = s->eps_target * s->eps_target; T.iter_dev = s->iter_dev;
    return T;
}

extern "C" int admm_tv_pass(admm_plan* p, admm_state* s, int node0, int nodes, int with_diag, void* stream) {
    if (int e = check_nodes(p, node0, nodes)) return e;
    if (!s) return fail(ADMM_ERR_ARG, "admm_tv_pass: null state");
    TvParams T = make_tv(p, s, node0, with_diag != 0, s->w_parity, (s->carry_r & 2) != 0);
    if (with_diag & 2) {   // deferred pass of a solve that ran CG iterations: r <- r - alpha Hp is still pending
        T.hp = s->hp + (long long)node0 * s->stride;
        if (with_diag & 4) T.r = s->r1 + (long long)node0 * s->stride;   // ... and the fused CG's ping-pong left r in r1
    }
    if (s->ctl && with_diag) T.accept = s->accept_mode;   // a deferred pass that ends a solve carries the a14 decision
    CK(launch_tv(T, nodes, (cudaStream_t)stream));
    return ADMM_OK;
}

extern "C" int admm_x_update(admm_plan* p, admm_state* s, int node0, int nodes, int sweeps, int cg_iters,
                             void* stream) {
    if (int e = check_nodes(p, node0, nodes)) return e;
    if (!s || sweeps < 1 || cg_iters < 0) return fail(ADMM_ERR_ARG, "admm_x_update: bad argument");
    if ((p->dense || p->impl != 0) && s->fuse_pupdate != 0)
        return fail(ADMM_ERR_ARG, "admm_x_update: dense-matrix and rotate-and-sum plans need fuse_pupdate = 0 (the fused CG staging lives in the strip projector)");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)p->N * p->N, off = (long long)node0 * s->stride;
    const long long part_per = admm_plan_info(p, ADMM_INFO_PART_FLOATS);
    float* part = s->part + (long long)node0 * part_per;
    unsigned* counter = s->counter + node0;
    const int A0 = p->aptr[node0], A1 = p->aptr[node0 + nodes];

    SinoParams SP{};
    SP.q = s->q; SP.ax = s->ax; SP.b = s->b; SP.anode = p->d_anode; SP.aptr = p->d_aptr; SP.A0 = A0; SP.A1 = A1;
    SP.D = p->D; SP.node0 = node0; SP.scal = s->scal;
    // a14 retry pass: every kernel of the solve skips the nodes that were accepted already
    const NodeCtl* mask = (s->ctl && s->masked) ? reinterpret_cast<const NodeCtl*>(s->ctl) : nullptr;
    SP.ctl = mask;

    int parity = s->w_parity;  // the caller flips st->w_parity (sweeps & 1) once every node group is done
    for (int sw = 0; sw < sweeps; ++sw) {
        // r = rhs0 + tvterm - H x ; p = r ; rr -> S_RR0 ; ax = A x
        if (!(s->reuse_ax || sw > 0)) {   // later sweeps: ax is current by the recurrence
            FwdParams F = make_fwd(p, s->x + off, s->stride, node0);
            F.ctl = mask;
            CK(plan_forward(p, F, nodes, make_red(p, s->ax, node0, nodes), st));
        }
        BackParams B = make_back(p, s->ax, s->prec, s->r + off, s->stride, node0);
        B.v = s->x + off; B.rhoD_vec = s->rhoD_vec ? s->rhoD_vec + off : nullptr; B.rhoD_s = s->rhoD_s; B.mu = s->mu;
        // fully fused CG: r ping-pongs (r -> r1 -> r ...), so the buffer holding r at the start of the solve stays intact
        // until the fused step of iteration 1 has read it as the old direction: p0 = r need not be written
        const bool direct_p = (s->fuse_pupdate == 2);
        B.rhs0 = s->rhs0 + off; B.tvterm = s->tvterm + off; B.p_out = direct_p ? nullptr : s->p0 + off;
        B.part = part; B.counter = counter; B.scal = s->scal; B.dot_slot = S_RR0; B.ctl = mask;
        // with carry_r the TV pass (r += tvterm' - tvterm) and the rhs0 assembly (r += rhs0' - rhs0) keep r = rhs0 +
        // tvterm - H x, p0 = r and <r,r> current, so the solve starts without this back-projection
        if (!((s->carry_r & 1) && (sw > 0 || s->reuse_r))) CK(plan_back(p, BACK_RESID0, B, nodes, st));
        int cur = 0;
        float* rcur = s->r + off;            // residual buffer currently holding r (fuse 2 ping-pongs r / r1)
        bool pending_r = false;              // the last CG update left r <- r - alpha Hp to the TV pass
        float* r_where = rcur;
        for (int it = 0; it < cg_iters; ++it) {
            const int rr_in = (it & 1) ? S_RR1 : S_RR0, rr_out = (it & 1) ? S_RR0 : S_RR1;
            // direct_p: iteration 0 projects r itself, iteration 1's fused step reads it (rcur is still that buffer) as the
            // old direction and writes p' to p1; from then on p1 / p0 ping-pong as usual
            float* pcur = (it <= 1 && direct_p) ? rcur : (cur ? s->p1 : s->p0) + off;
            float* poth = (cur ? s->p0 : s->p1) + off;
            if (it > 0 && s->fuse_pupdate == 2) {
                // x += alpha p ; r' = r - alpha Hp ; p' = r' + beta p ; <r',r'> -> rr_in  -- all inside the projector
                float* roth = (rcur == s->r + off) ? s->r1 + off : s->r + off;
                FwdParams Fp = make_fwd(p, pcur, s->stride, node0);
                Fp.mode = 2; Fp.r = rcur; Fp.r_out = roth; Fp.p_out = poth; Fp.hp = s->hp + off; Fp.x_io = s->x + off;
                Fp.scal = s->scal; Fp.beta_den = rr_out /* slot of the previous <r,r> */; Fp.rr_out = rr_in;
                Fp.part = part; Fp.counter = counter; Fp.ctl = mask;
                CK(plan_forward(p, Fp, nodes, make_red(p, s->q, node0, nodes), st));
                cur ^= 1; pcur = poth; rcur = roth;
            } else if (it > 0 && s->fuse_pupdate == 1) {
                FwdParams Fp = make_fwd(p, pcur, s->stride, node0);
                Fp.mode = 1; Fp.r = s->r + off; Fp.p_out = poth; Fp.scal = s->scal; Fp.beta_num = rr_in; Fp.beta_den = rr_out;
                Fp.ctl = mask;
                CK(plan_forward(p, Fp, nodes, make_red(p, s->q, node0, nodes), st));
                cur ^= 1; pcur = poth;
            } else if (it > 0) {
                CgParams U{};
                U.r = s->r + off; U.p = pcur; U.p_out = poth; U.stride = s->stride; U.n = n; U.node0 = node0;
                U.rr_in = rr_out; U.rr_out = rr_in;  // beta = scal[rr_in(it)] / scal[rr_out(it)] = new / old
                U.scal = s->scal; U.ctl = mask;
                CK(launch_p_update(U, nodes, st));
                cur ^= 1; pcur = poth;
                FwdParams Fp = make_fwd(p, pcur, s->stride, node0);
                Fp.ctl = mask;
                CK(plan_forward(p, Fp, nodes, make_red(p, s->q, node0, nodes), st));
            } else {
                FwdParams Fp = make_fwd(p, pcur, s->stride, node0);
                Fp.ctl = mask;
                CK(plan_forward(p, Fp, nodes, make_red(p, s->q, node0, nodes), st));
            }
            BackParams H = make_back(p, s->q, s->prec, s->hp + off, s->stride, node0);
            H.v = pcur; H.rhoD_vec = B.rhoD_vec; H.rhoD_s = s->rhoD_s; H.mu = s->mu;
            H.rvec = (s->fuse_pupdate == 2) ? rcur : nullptr;
            H.part = part; H.counter = counter; H.scal = s->scal; H.dot_slot = S_PHP; H.ctl = mask;
            CK(plan_back(p, BACK_HP, H, nodes, st));
            SP.mode = 1; SP.rr_in = rr_in;
            CK(launch_sino_axpy(SP, st));
            const bool last = (it == cg_iters - 1);
            if (s->fuse_pupdate != 2 || last) {
                // separate vector update (every iteration when unfused; once, after the last one, when fused):
                // the final residual always lands in s->r
                CgParams U{};
                U.x = s->x + off; U.r = s->r + off; U.r_in = rcur; U.p = pcur; U.hp = s->hp + off; U.stride = s->stride;
                U.n = n; U.node0 = node0; U.rr_in = rr_in; U.rr_out = rr_out; U.part = part; U.counter = counter;
                U.scal = s->scal; U.ctl = mask;
                // the solve's last update: x only; the TV pass that ends the sweep (here or deferred) applies r -= alpha Hp,
                // reading r where the CG left it (the fused form's ping-pong: r or r1) and writing the carried one to s->r
                U.x_only = last ? 1 : 0;
                if (last) { pending_r = true; r_where = rcur; }
                long long nb = (n / 4 + 256 * 4 - 1) / (256 * 4);
                nb = std::max(1LL, std::min(nb, 4096LL));
                CK(launch_cg_update(U, nodes, (int)nb, st));
            }
        }
        if (!(s->defer_tv && sw == sweeps - 1)) {
            // hand the residual on to the next solve (bit 1), and always to the next sweep of THIS solve when sweeps carry
            const bool hand_on = (s->carry_r & 2) || ((s->carry_r & 1) && sw < sweeps - 1);
            TvParams T = make_tv(p, s, node0, true, parity, hand_on);
            if (pending_r) { T.hp = s->hp + off; T.r = r_where; }
            if (s->ctl && sw == sweeps - 1) T.accept = s->accept_mode;   // the a14 decision rides on the solve's last TV pass
            CK(launch_tv(T, nodes, st));
            parity ^= 1;
        }
    }
    if (!s->skip_mse) CK(launch_sino_resid(SP, nodes, st));
    return ADMM_OK;
}

extern "C" int admm_accept(admm_plan* p, admm_state* s, int node0, int nodes, double eps_target, int max_tighten,
                           int first, void* stream) {
    if (int e = check_nodes(p, node0, nodes)) return e;
    if (!s || !s->ctl) return fail(ADMM_ERR_ARG, "admm_accept: the state has no node control table");
    AcceptParams A{};
    A.ctl = reinterpret_cast<NodeCtl*>(s->ctl); A.scal = s->scal; A.node0 = node0; A.nodes = nodes;
    A.first = first; A.max_tighten = max_tighten; A.eps_target2 = eps_target * eps_target;
    A.iter_dev = s->iter_dev;
    CK(launch_accept(A, (cudaStream_t)stream));
    return ADMM_OK;
}

extern "C" int admm_edge_update(admm_plan* p, const admm_state* s, const admm_edge* d_edges, int nedges,
                                double* d_sums, void* stream) {
    if (!p || !s) return fail(ADMM_ERR_ARG, "admm_edge_update: null argument");
    if (nedges <= 0) return ADMM_OK;
    EdgeParams E{};
    E.edges = reinterpret_cast<const EdgeDesc*>(d_edges); E.n = (long long)p->N * p->N; E.q_uniform = s->q_uniform;
    E.part = s->part; E.counter = s->counter; E.sums = d_sums;
    long long nb = (E.n + 256 * 4 - 1) / (256 * 4);
    nb = std::max(1LL, std::min(nb, 1024LL));
    // part: 5 floats per block per edge must fit the per-unit budget
    if (nb * 5 > admm_plan_info(p, ADMM_INFO_PART_FLOATS)) nb = admm_plan_info(p, ADMM_INFO_PART_FLOATS) / 5;
    // per-edge workspace stride is nb*5 floats (<= PART_FLOATS)
    CK(launch_edges(E, nedges, (int)nb, (cudaStream_t)stream));
    return ADMM_OK;
}

extern "C" int admm_pack(admm_plan* p, const admm_pack_item* d_items, int nitems, void* stream) {
    if (!p) return fail(ADMM_ERR_ARG, "admm_pack: null plan");
    PackParams K{};
    K.items = reinterpret_cast<const PackDesc*>(d_items); K.n = (long long)p->N * p->N; K.nitems = nitems;
    CK(launch_pack(K, nitems, p->pack_blocks, (cudaStream_t)stream));
    return ADMM_OK;
}

// Copy-engine form of the x push of the single-owner exchange: one cudaMemcpyAsync per item (x -> out, n floats) on the
// given stream; `h_items` is a HOST array.  No SM is used, so the transfer does not compete with the TV / edge kernels.
extern "C" int admm_push_copy(admm_plan* p, const admm_pack_item* h_items, int nitems, void* stream) {
    if (!p || (nitems > 0 && !h_items)) return fail(ADMM_ERR_ARG, "admm_push_copy: null argument");
    const size_t bytes = (size_t)p->N * p->N * sizeof(float);
    for (int k = 0; k < nitems; ++k) {
        if (h_items[k].y != 0) return fail(ADMM_ERR_ARG, "admm_push_copy: item is not a plain copy (y != 0)");
        CK(cudaMemcpyAsync(reinterpret_cast<void*>(h_items[k].out), reinterpret_cast<const void*>(h_items[k].x), bytes,
                           cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    return ADMM_OK;
}

extern "C" int admm_finalize(admm_plan* p, const admm_state* s, const double* d_sums, const int* d_edge_gi,
                             const int* d_edge_gj, const int* d_edge_flags, int nedges, int nedges_local,
                             const int* d_node_gid, const int* d_nbr_ptr, const int* d_nbr_epos, const int* d_nbr_end,
                             int Vg, double* d_row, void* stream) {
    if (!p || !s || !d_row || !d_node_gid || !d_nbr_ptr) return fail(ADMM_ERR_ARG, "admm_finalize: null argument");
    FinalizeParams F{};
    F.sums = d_sums; F.edge_gi = d_edge_gi; F.edge_gj = d_edge_gj; F.edge_flags = d_edge_flags; F.scal = s->scal;
    F.node_gid = d_node_gid; F.row = d_row; F.E = nedges; F.E_local = nedges_local; F.V = p->V; F.Vg = Vg; F.rho = s->rho;
    F.nbr_ptr = d_nbr_ptr; F.nbr_epos = d_nbr_epos; F.nbr_end = d_nbr_end;
    F.ctl = reinterpret_cast<const NodeCtl*>(s->ctl);
    F.iter_dev = s->iter_dev; F.hist_stride = s->hist_stride;
    CK(launch_finalize(F, (cudaStream_t)stream));
    return ADMM_OK;
}

// ---- PDHG consensus variant (ADMM_Tomo_Only.py:89-148; kernels in pdhg.cu) ---------------------------------------
extern "C" int admm_pdhg_dual(admm_plan* p, const float* d_xbar, long long stride, float* d_y1, float* d_y2,
                              const float* d_q, const float* d_b, const float* d_sigma, float lam_data, float lam_tv,
                              int node0, int nodes, void* stream) {
    if (int e = check_nodes(p, node0, nodes)) return e;
    if (!d_xbar || !d_y1 || !d_y2 || !d_q || !d_b || !d_sigma || lam_data <= 0.f || lam_tv <= 0.f)
        return fail(ADMM_ERR_ARG, "admm_pdhg_dual: bad argument");
    CK(launch_pdhg_dual(d_y1, d_y2, d_xbar, stride, d_q, d_b, p->d_anode, d_sigma, lam_data, lam_tv, p->N, p->D,
                        p->aptr[node0], p->aptr[node0 + nodes], node0, nodes, (cudaStream_t)stream));
    return ADMM_OK;
}
extern "C" int admm_pdhg_primal(admm_plan* p, float* d_x, float* d_xbar, long long stride, const float* d_back,
                                const float* d_y2, const float* d_pull, const float* d_tau, const float* d_adj,
                                float gamma, float theta, int node0, int nodes, void* stream) {
    if (int e = check_nodes(p, node0, nodes)) return e;
    if (!d_x || !d_xbar || !d_back || !d_y2 || !d_tau || !d_adj) return fail(ADMM_ERR_ARG, "admm_pdhg_primal: null buffer");
    CK(launch_pdhg_primal(d_x, d_xbar, stride, d_back, d_y2, d_pull, d_tau, d_adj, gamma, theta, p->N, node0, nodes,
                          (cudaStream_t)stream));
    return ADMM_OK;
}
extern "C" int admm_pdhg_normal(admm_plan* p, const float* d_x, long long stride, const float* d_back, const float* d_adj,
                                float* d_out, int node0, int nodes, void* stream) {
    if (int e = check_nodes(p, node0, nodes)) return e;
    if (!d_x || !d_back || !d_adj || !d_out) return fail(ADMM_ERR_ARG, "admm_pdhg_normal: null buffer");
    CK(launch_pdhg_normal(d_out, d_x, stride, d_back, d_adj, p->N, node0, nodes, (cudaStream_t)stream));
    return ADMM_OK;
}
extern "C" int admm_pdhg_combine(admm_plan* p, const float* d_x, long long stride, const float* d_colnorm,
                                 const float* d_phantom, float* d_xa, int nodes, void* stream) {
    if (int e = check_nodes(p, 0, nodes)) return e;
    if (!d_x || !d_colnorm || !d_phantom || !d_xa) return fail(ADMM_ERR_ARG, "admm_pdhg_combine: null buffer");
    CK(launch_pdhg_combine(d_xa, d_x, stride, d_colnorm, d_phantom, (long long)p->N * p->N, nodes, (cudaStream_t)stream));
    return ADMM_OK;
}
extern "C" int admm_pdhg_sums(admm_plan* p, const float* d_x, long long stride, const float* d_phantom, const float* d_q,
                              const float* d_b, double* d_out, int node0, int nodes, void* stream) {
    if (int e = check_nodes(p, node0, nodes)) return e;
    if (!d_x || !d_out) return fail(ADMM_ERR_ARG, "admm_pdhg_sums: null buffer");
    CK(launch_pdhg_sums(d_out, d_x, stride, d_phantom, d_q, d_b, p->d_aptr, (long long)p->N * p->N, p->D, node0, nodes,
                        (cudaStream_t)stream));
    return ADMM_OK;
}

// ---- peer-memory exchange buffers (CUDA IPC over NVLink) -------------------------------------------------------
// The cut-edge iterates a = x + y are packed into a buffer that the owning rank allocates here and every peer maps
// with cudaIpcOpenMemHandle; the peers' edge kernels then read the remote a directly over NVLink (no staging copy,
// no send/recv kernels): the transfer overlaps the edge kernel's own HBM work.
extern "C" int admm_ipc_alloc(long long bytes, void** d_ptr, unsigned char* handle64) {
    if (bytes <= 0 || !d_ptr || !handle64) return fail(ADMM_ERR_ARG, "admm_ipc_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    CK(cudaMalloc(&p, (size_t)bytes));
    CK(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(ADMM_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    }
    memcpy(handle64, &h, 64);
    *d_ptr = p;
    return ADMM_OK;
}
extern "C" int admm_ipc_open(const unsigned char* handle64, void** d_ptr) {
    if (!handle64 || !d_ptr) return fail(ADMM_ERR_ARG, "admm_ipc_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return ADMM_OK;
}
extern "C" int admm_ipc_close(void* d_ptr) {
    if (d_ptr) CK(cudaIpcCloseMemHandle(d_ptr));
    return ADMM_OK;
}
extern "C" int admm_ipc_free(void* d_ptr) {
    if (d_ptr) CK(cudaFree(d_ptr));
    return ADMM_OK;
}
