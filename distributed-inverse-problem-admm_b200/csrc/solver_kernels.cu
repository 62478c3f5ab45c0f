// solver_kernels.cu -- K3 fused TV stencil, K4 CG vector updates with in-kernel reductions, K5 edge
// (consensus z / dual y / residual) kernel, K6 right-hand-side assembly, and the small per-iteration
// bookkeeping kernels.  sm_100a.  All are single-pass streaming kernels (HBM-bound).
#include <cstdlib>
#include <mutex>

#include "solver_kernels.cuh"

namespace admm {

// =================================================================================================
// K3: one pass does  g = Kx + w ; d = shrink2(g, lam/mu) ; w' = g - d ; tvterm' = mu K^T (d - w')
// plus the diagnostics that ride on the same data: canonical TV(x), the a14 stationarity norm
//   g_x = tvterm - r_cg - mu K^T K x + lam * div_ref(normalised grad x)   (block_6_admm_loop_ver2.py:137-149
// through the identity A^T P(Ax-b) + rho(Dx - sum q v) = tvterm - r_cg - mu K^T K x), and |x - x_true|^2
// (:199-206).  Gradient = block_4_tv_helpers.py:17-23; div_ref = :25-35 as shipped (border sign quirk).
// =================================================================================================
// Streaming form: block (32, 8) covers 8 rows x 128 columns, each thread owns 4 consecutive pixels of one row
// (float4 loads / stores) and recomputes the (d - w') of its upper and left neighbours in registers, so there is
// no shared-memory tile, no barrier before the reduction and every global access is a coalesced 16-byte one.
constexpr int TVX = 32, TVY = 8;

// one node's a14 decision (block_6_admm_loop_ver2.py:155-176): accepted if |g| <= eps_target (:155) or the tighten cap is
// reached (:164), else one more try (:175)
__device__ __forceinline__ void accept_node(NodeCtl* ctl, double gn2, bool first, int max_tighten, double tgt2,
                                            const int* iter_dev) {
    NodeCtl c = *ctl;
    if (first) { c.active = 1; c.tries = 0; }
    if (iter_dev) {   // eps_target = 2 / (k+1)^1.005 from the device's own iteration counter (:101-103)
        const double t = 2.0 / pow((double)(*iter_dev + 1), 1.005);
        tgt2 = t * t;
    }
    if (c.active) {
        if (gn2 <= tgt2 || c.tries >= max_tighten) c.active = 0;
        else c.tries += 1;
    }
    ctl->active = c.active;
    ctl->tries = c.tries;
}


struct Dw { float dwx, dwy, w1, w2; };

__device__ __forceinline__ Dw shrink_dw(float gx, float gy, float w1, float w2, float kappa) {
    const float g1 = gx + w1, g2 = gy + w2;
    const float n2 = fmaf(g1, g1, g2 * g2);
    const float sc = (n2 > kappa * kappa) ? (1.f - kappa * rsqrtf(n2)) : 0.f;
    const float d1 = sc * g1, d2 = sc * g2;
    Dw o;
    o.w1 = g1 - d1; o.w2 = g2 - d2;
    o.dwx = d1 - o.w1; o.dwy = d2 - o.w2;
    return o;
}

__device__ __forceinline__ void unit_grad(float gx, float gy, float& px, float& py, float& mag) {
    const float n2 = fmaf(gx, gx, gy * gy);
    if (n2 > 1e-24f) {
        const float inv = rsqrtf(n2);
        px = gx * inv; py = gy * inv; mag = n2 * inv;
    } else { px = 0.f; py = 0.f; mag = 0.f; }
}

// guarded row-segment load: v[k] = row[c + k] for k in [0, CNT), zero outside [0, N)
template <int CNT>
__device__ __forceinline__ void load_seg(const float* __restrict__ row, int c, int N, bool rowok, bool vec, float* v) {
#pragma unroll
    for (int k = 0; k < CNT; ++k) v[k] = 0.f;
    if (!rowok) return;
    // layout of every caller: element 1..4 are the aligned quad when CNT >= 5 and the segment starts at c0-1
#pragma unroll
    for (int k = 0; k < CNT; ++k) {
        const int cc = c + k;
        if (cc >= 0 && cc < N) v[k] = row[cc];
    }
    (void)vec;
}

// The four pixels (r, c0 .. c0+3) of one thread.  INTR: the block touches no image border and N % 4 == 0, so every
// neighbour exists and every boundary test folds away at compile time (87 % of the blocks at 2048^2; the general form
// spends a third of its instructions on those tests).  Same arithmetic, same order: results are bit-identical.
template <bool INTR>
__device__ __forceinline__ void tv_quad(const TvParams& P, const float* __restrict__ x, const float* __restrict__ w1p,
                                        const float* __restrict__ w2p, float* __restrict__ wo1, float* __restrict__ wo2,
                                        long long nb, int N, int r, int c0, float kappa, float alpha, float& tv, float& gn2,
                                        float& img, float& rr) {
    const bool vec = INTR || ((N & 3) == 0);   // then c0 + 3 < N and rows are 16-byte aligned
    const long long g0 = (long long)r * N + c0;
    float xm[5], xc[6], xp[5], wc1[5], wc2[5], wu1[4], wu2[4];
    const bool up = INTR || r >= 1, dn = INTR || r + 1 < N;
    if (vec) {
        const float4 q = ld4(x + g0);
        xc[1] = q.x; xc[2] = q.y; xc[3] = q.z; xc[4] = q.w;
        xc[0] = (INTR || c0 >= 1) ? x[g0 - 1] : 0.f;
        xc[5] = (INTR || c0 + 4 < N) ? x[g0 + 4] : 0.f;
        if (up) {
            const float4 a = ld4(x + g0 - N);
            xm[0] = a.x; xm[1] = a.y; xm[2] = a.z; xm[3] = a.w;
            xm[4] = (INTR || c0 + 4 < N) ? x[g0 - N + 4] : 0.f;
            const float4 u1 = ld4(w1p + g0 - N), u2 = ld4(w2p + g0 - N);
            wu1[0] = u1.x; wu1[1] = u1.y; wu1[2] = u1.z; wu1[3] = u1.w;
            wu2[0] = u2.x; wu2[1] = u2.y; wu2[2] = u2.z; wu2[3] = u2.w;
        } else {
#pragma unroll
            for (int k = 0; k < 5; ++k) xm[k] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) { wu1[k] = 0.f; wu2[k] = 0.f; }
        }
        if (dn) {
            const float4 a = ld4(x + g0 + N);
            xp[1] = a.x; xp[2] = a.y; xp[3] = a.z; xp[4] = a.w;
            xp[0] = (INTR || c0 >= 1) ? x[g0 + N - 1] : 0.f;
        } else {
#pragma unroll
            for (int k = 0; k < 5; ++k) xp[k] = 0.f;
        }
        const float4 a1 = ld4(w1p + g0), a2 = ld4(w2p + g0);
        wc1[1] = a1.x; wc1[2] = a1.y; wc1[3] = a1.z; wc1[4] = a1.w;
        wc2[1] = a2.x; wc2[2] = a2.y; wc2[3] = a2.z; wc2[4] = a2.w;
        wc1[0] = (INTR || c0 >= 1) ? w1p[g0 - 1] : 0.f;
        wc2[0] = (INTR || c0 >= 1) ? w2p[g0 - 1] : 0.f;
    } else {
        load_seg<6>(x + (long long)r * N, c0 - 1, N, true, false, xc);
        load_seg<5>(x + (long long)(r - 1) * N, c0, N, up, false, xm);
        load_seg<5>(x + (long long)(r + 1) * N, c0 - 1, N, dn, false, xp);
        load_seg<5>(w1p + (long long)r * N, c0 - 1, N, true, false, wc1);
        load_seg<5>(w2p + (long long)r * N, c0 - 1, N, true, false, wc2);
        load_seg<4>(w1p + (long long)(r - 1) * N, c0, N, up, false, wu1);
        load_seg<4>(w2p + (long long)(r - 1) * N, c0, N, up, false, wu2);
    }
    // left neighbour (r, c0-1): only its y-component of (d - w') and of the unit gradient is needed
    float dwy_prev = 0.f, py_prev = 0.f;
    if (INTR || c0 >= 1) {
        const float gx = dn ? xp[0] - xc[0] : 0.f, gy = xc[1] - xc[0];
        dwy_prev = shrink_dw(gx, gy, wc1[0], wc2[0], kappa).dwy;
        float pxl, mg;
        unit_grad(gx, gy, pxl, py_prev, mg);
    }
    float w1o[4], w2o[4], tvo[4];
    float told[4], rc[4], xt[4];
    const bool diag = (P.r != nullptr);
    const bool upd = (P.r_upd != nullptr);
    const float* rsrc = P.r ? P.r : P.r_upd;    // read where the CG left r; the carried one is written to r_upd
    float* __restrict__ tvt = P.tvterm + nb;
    if (vec) {
        const float4 t4 = ld4(tvt + g0);
        told[0] = t4.x; told[1] = t4.y; told[2] = t4.z; told[3] = t4.w;
        if (diag || upd) {
            const float4 q = ld4(rsrc + nb + g0);
            rc[0] = q.x; rc[1] = q.y; rc[2] = q.z; rc[3] = q.w;
            if (P.hp) {      // the pending half of the last CG update: r <- r - alpha Hp
                const float4 h4 = ld4(P.hp + nb + g0);
                rc[0] = fmaf(-alpha, h4.x, rc[0]); rc[1] = fmaf(-alpha, h4.y, rc[1]);
                rc[2] = fmaf(-alpha, h4.z, rc[2]); rc[3] = fmaf(-alpha, h4.w, rc[3]);
            }
        }
        if (P.xtrue) { const float4 q = ld4(P.xtrue + g0); xt[0] = q.x; xt[1] = q.y; xt[2] = q.z; xt[3] = q.w; }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool ok = c0 + k < N;
            told[k] = ok ? tvt[g0 + k] : 0.f;
            rc[k] = (ok && (diag || upd)) ? rsrc[nb + g0 + k] : 0.f;
            if (ok && (diag || upd) && P.hp) rc[k] = fmaf(-alpha, P.hp[nb + g0 + k], rc[k]);
            xt[k] = (ok && P.xtrue) ? P.xtrue[g0 + k] : 0.f;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + k;
        const bool ok = INTR || c < N, rt = INTR || c + 1 < N, lf = INTR || c >= 1;
        const float xcv = xc[k + 1];
        const float gx = dn ? xp[k + 1] - xcv : 0.f;
        const float gy = rt ? xc[k + 2] - xcv : 0.f;
        const Dw o = shrink_dw(gx, gy, wc1[k + 1], wc2[k + 1], kappa);
        float dwx_up = 0.f, px_up = 0.f;
        if (up) {
            const float gxu = xcv - xm[k], gyu = rt ? xm[k + 1] - xm[k] : 0.f;
            dwx_up = shrink_dw(gxu, gyu, wu1[k], wu2[k], kappa).dwx;
            float pyu, mg;
            unit_grad(gxu, gyu, px_up, pyu, mg);
        }
        float kt = dwx_up;
        if (dn) kt -= o.dwx;
        if (lf) kt += dwy_prev;
        if (rt) kt -= o.dwy;
        float pxo, pyo, mag;
        unit_grad(gx, gy, pxo, pyo, mag);
        if (ok) {
            tv += mag;
            if (diag) {
                float lap = 0.f;
                if (up) lap += xcv - xm[k];
                if (dn) lap += xcv - xp[k + 1];
                if (lf) lap += xcv - xc[k];
                if (rt) lap += xcv - xc[k + 2];
                float dv = 0.f;   // block_4_tv_helpers.py:25-35 as shipped (-div, sign-flipped border)
                if (INTR || N >= 2) {
                    if (!INTR && r == 0) dv += pxo; else if (!dn) dv -= px_up; else dv += px_up - pxo;
                    if (!INTR && c == 0) dv += pyo; else if (!rt) dv -= py_prev; else dv += py_prev - pyo;
                }
                const float gv = told[k] - rc[k] - P.mu * lap + P.lam * dv;
                gn2 = fmaf(gv, gv, gn2);
            }
            if (P.xtrue) { const float e = xcv - xt[k]; img = fmaf(e, e, img); }
        }
        w1o[k] = o.w1; w2o[k] = o.w2; tvo[k] = P.mu * kt;
        dwy_prev = o.dwy; py_prev = pyo;
    }
    float rn[4];   // residual of the NEXT solve: r + (tvterm' - tvterm)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        rn[k] = rc[k] + (tvo[k] - told[k]);
        if (upd && (INTR || c0 + k < N)) rr = fmaf(rn[k], rn[k], rr);
    }
    if (vec) {
        st4(wo1 + g0, make_float4(w1o[0], w1o[1], w1o[2], w1o[3]));
        st4(wo2 + g0, make_float4(w2o[0], w2o[1], w2o[2], w2o[3]));
        st4(tvt + g0, make_float4(tvo[0], tvo[1], tvo[2], tvo[3]));
        if (upd) {
            st4(P.r_upd + nb + g0, make_float4(rn[0], rn[1], rn[2], rn[3]));
            if (P.p_out) st4(P.p_out + nb + g0, make_float4(rn[0], rn[1], rn[2], rn[3]));
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (c0 + k < N) {
                wo1[g0 + k] = w1o[k]; wo2[g0 + k] = w2o[k]; tvt[g0 + k] = tvo[k];
                if (upd) { P.r_upd[nb + g0 + k] = rn[k]; if (P.p_out) P.p_out[nb + g0 + k] = rn[k]; }
            }
    }
}

// Interior blocks march down the rows: thread (lane, band) owns the 4 columns c0..c0+3 of TVROWS consecutive rows.  The
// (d - w') and the unit gradient of row r ARE the "upper neighbour" values of row r + 1, so they are carried in
// registers instead of being recomputed (5.5 shrink + unit-gradient evaluations per 4 pixels instead of 9), x rows
// rotate through registers (one new x row per step).  Per pixel the operations and their order are those of tv_quad<true>: results are bit-identical.
constexpr int TVROWS = 8;

template <int PX> struct VecPx;
template <> struct VecPx<4> {
    static __device__ __forceinline__ void ld(const float* p, float* o) { const float4 v = ld4(p); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
    static __device__ __forceinline__ void st(float* p, const float* v) { st4(p, make_float4(v[0], v[1], v[2], v[3])); }
};
template <> struct VecPx<2> {
    static __device__ __forceinline__ void ld(const float* p, float* o) { const float2 v = *reinterpret_cast<const float2*>(p); o[0] = v.x; o[1] = v.y; }
    static __device__ __forceinline__ void st(float* p, const float* v) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};

// PX = pixels per thread and row (4: float4 accesses, 80 registers; 2: float2 accesses, half the per-thread state -> twice
// the resident warps, at one more left-neighbour evaluation per 4 pixels)
template <int PX>
__device__ __forceinline__ void tv_strip(const TvParams& P, const float* __restrict__ x, const float* __restrict__ w1p,
                                         const float* __restrict__ w2p, float* __restrict__ wo1, float* __restrict__ wo2,
                                         long long nb, int N, int r0, int c0, float kappa, float alpha, float& tv, float& gn2,
                                         float& img, float& rr) {
    typedef VecPx<PX> V;
    const bool diag = (P.r != nullptr);
    const bool upd = (P.r_upd != nullptr);
    const float* rsrc = P.r ? P.r : P.r_upd;    // read where the CG left r; the carried one is written to r_upd
    float* __restrict__ tvt = P.tvterm + nb;
    long long g0 = (long long)r0 * N + c0;
    // prologue: rows r0 - 1 and r0, and the upper-neighbour values of row r0
    float xm[PX], xc[PX + 2], dwx_up[PX], px_up[PX];
    {
        float xm5[PX + 1], wu1[PX], wu2[PX];
        V::ld(x + g0 - N, xm5);
        xm5[PX] = x[g0 - N + PX];
        V::ld(x + g0, xc + 1);
        xc[0] = x[g0 - 1];
        xc[PX + 1] = x[g0 + PX];
        V::ld(w1p + g0 - N, wu1);
        V::ld(w2p + g0 - N, wu2);
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            xm[k] = xm5[k];
            const float gxu = xc[k + 1] - xm5[k], gyu = xm5[k + 1] - xm5[k];
            dwx_up[k] = shrink_dw(gxu, gyu, wu1[k], wu2[k], kappa).dwx;
            float pyu, mg;
            unit_grad(gxu, gyu, px_up[k], pyu, mg);
        }
    }
#pragma unroll 1
    for (int i = 0; i < TVROWS; ++i, g0 += N) {
        // loads of one row: x of the row below (c0-1 .. c0+PX), w (c0-1 .. c0+PX-1), tvterm, r (- alpha Hp), x_true
        float xp[PX + 2], wc1[PX + 1], wc2[PX + 1], told[PX], rc[PX], xt[PX];
        V::ld(x + g0 + N, xp + 1);
        xp[0] = x[g0 + N - 1];
        xp[PX + 1] = x[g0 + N + PX];
        V::ld(w1p + g0, wc1 + 1);
        V::ld(w2p + g0, wc2 + 1);
        wc1[0] = w1p[g0 - 1];
        wc2[0] = w2p[g0 - 1];
        V::ld(tvt + g0, told);
#pragma unroll
        for (int k = 0; k < PX; ++k) { rc[k] = 0.f; xt[k] = 0.f; }
        if (diag || upd) {
            V::ld(rsrc + nb + g0, rc);
            if (P.hp) {      // the pending half of the last CG update: r <- r - alpha Hp
                float h[PX];
                V::ld(P.hp + nb + g0, h);
#pragma unroll
                for (int k = 0; k < PX; ++k) rc[k] = fmaf(-alpha, h[k], rc[k]);
            }
        }
        if (P.xtrue) V::ld(P.xtrue + g0, xt);
        float dwy_prev, py_prev;
        {
            const float gx = xp[0] - xc[0], gy = xc[1] - xc[0];
            dwy_prev = shrink_dw(gx, gy, wc1[0], wc2[0], kappa).dwy;
            float pxl, mg;
            unit_grad(gx, gy, pxl, py_prev, mg);
        }
        float w1o[PX], w2o[PX], tvo[PX];
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            const float xcv = xc[k + 1];
            const float gx = xp[k + 1] - xcv, gy = xc[k + 2] - xcv;
            const Dw o = shrink_dw(gx, gy, wc1[k + 1], wc2[k + 1], kappa);
            float kt = dwx_up[k];
            kt -= o.dwx;
            kt += dwy_prev;
            kt -= o.dwy;
            float pxo, pyo, mag;
            unit_grad(gx, gy, pxo, pyo, mag);
            tv += mag;
            if (diag) {
                float lap = 0.f;
                lap += xcv - xm[k];
                lap += xcv - xp[k + 1];
                lap += xcv - xc[k];
                lap += xcv - xc[k + 2];
                float dv = 0.f;
                dv += px_up[k] - pxo;
                dv += py_prev - pyo;
                const float gv = told[k] - rc[k] - P.mu * lap + P.lam * dv;
                gn2 = fmaf(gv, gv, gn2);
            }
            if (P.xtrue) { const float e = xcv - xt[k]; img = fmaf(e, e, img); }
            w1o[k] = o.w1; w2o[k] = o.w2; tvo[k] = P.mu * kt;
            dwy_prev = o.dwy; py_prev = pyo;
            dwx_up[k] = o.dwx; px_up[k] = pxo;      // the row below sees this row as its upper neighbour
        }
        V::st(wo1 + g0, w1o);
        V::st(wo2 + g0, w2o);
        V::st(tvt + g0, tvo);
        if (upd) {
            float rn[PX];
#pragma unroll
            for (int k = 0; k < PX; ++k) { rn[k] = rc[k] + (tvo[k] - told[k]); rr = fmaf(rn[k], rn[k], rr); }
            V::st(P.r_upd + nb + g0, rn);
            if (P.p_out) V::st(P.p_out + nb + g0, rn);
        }
#pragma unroll
        for (int k = 0; k < PX; ++k) xm[k] = xc[k + 1];
#pragma unroll
        for (int k = 0; k < PX + 2; ++k) xc[k] = xp[k];
    }
}

// ---- the same march with the NEXT row's operands in flight: every thread copies them global -> shared with cp.async
// (LDGSTS: no registers held while they travel) into its own slots of a double buffer and reads them back when the row's
// turn comes; nothing is shared between threads, so no barrier is involved.  8 slots of 16 bytes per thread and buffer:
// x(row+1), w1, w2, tvterm, r, Hp, x_true, {x(row+1)[c0-1], x(row+1)[c0+4], w1[c0-1], w2[c0-1]}.
constexpr int TVSLOTS = 8;
constexpr size_t kTvAsyncSmem = (size_t)2 * TVSLOTS * TVX * TVY * 16;

__device__ __forceinline__ void cp_async16(unsigned dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(unsigned dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NLEFT>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NLEFT) : "memory"); }

__device__ __forceinline__ void tv_strip_async(const TvParams& P, const float* __restrict__ x, const float* __restrict__ w1p,
                                               const float* __restrict__ w2p, float* __restrict__ wo1,
                                               float* __restrict__ wo2, long long nb, int N, int r0, int c0, float kappa,
                                               float alpha, unsigned char* smem, float& tv, float& gn2, float& img, float& rr) {
    const bool diag = (P.r != nullptr);
    const bool upd = (P.r_upd != nullptr);
    const bool need_r = diag || upd;
    const float* rsrc = P.r ? P.r : P.r_upd;
    float* __restrict__ tvt = P.tvterm + nb;
    long long g0 = (long long)r0 * N + c0;
    const int tid = threadIdx.y * TVX + threadIdx.x;
    const unsigned sbase = smem_u32(smem) + (unsigned)tid * 16u;
    const float4* sv = reinterpret_cast<const float4*>(smem) + tid;
    constexpr unsigned SLOT = TVX * TVY * 16, BUF = TVSLOTS * SLOT;   // bytes
    auto issue = [&](long long g, int buf) {
        const unsigned d = sbase + (unsigned)buf * BUF;
        cp_async16(d + 0 * SLOT, x + g + N);
        cp_async16(d + 1 * SLOT, w1p + g);
        cp_async16(d + 2 * SLOT, w2p + g);
        cp_async16(d + 3 * SLOT, tvt + g);
        if (need_r) cp_async16(d + 4 * SLOT, rsrc + nb + g);
        if (need_r && P.hp) cp_async16(d + 5 * SLOT, P.hp + nb + g);
        if (P.xtrue) cp_async16(d + 6 * SLOT, P.xtrue + g);
        cp_async4(d + 7 * SLOT + 0, x + g + N - 1);
        cp_async4(d + 7 * SLOT + 4, x + g + N + 4);
        cp_async4(d + 7 * SLOT + 8, w1p + g - 1);
        cp_async4(d + 7 * SLOT + 12, w2p + g - 1);
        cp_async_commit();
    };
    issue(g0, 0);
    // prologue: rows r0 - 1 and r0, and the upper-neighbour values of row r0
    float xm[4], xc[6], dwx_up[4], px_up[4];
    {
        const float4 a = ld4(x + g0 - N), q = ld4(x + g0);
        const float a4 = x[g0 - N + 4];
        const float4 u1 = ld4(w1p + g0 - N), u2 = ld4(w2p + g0 - N);
        xm[0] = a.x; xm[1] = a.y; xm[2] = a.z; xm[3] = a.w;
        xc[0] = x[g0 - 1]; xc[1] = q.x; xc[2] = q.y; xc[3] = q.z; xc[4] = q.w; xc[5] = x[g0 + 4];
        const float xm5[5] = {a.x, a.y, a.z, a.w, a4};
        const float wu1[4] = {u1.x, u1.y, u1.z, u1.w}, wu2[4] = {u2.x, u2.y, u2.z, u2.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gxu = xc[k + 1] - xm5[k], gyu = xm5[k + 1] - xm5[k];
            dwx_up[k] = shrink_dw(gxu, gyu, wu1[k], wu2[k], kappa).dwx;
            float pyu, mg;
            unit_grad(gxu, gyu, px_up[k], pyu, mg);
        }
    }
#pragma unroll 1
    for (int i = 0; i < TVROWS; ++i, g0 += N) {
        const int buf = i & 1;
        if (i + 1 < TVROWS) { issue(g0 + N, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        const float4* b = sv + buf * (TVSLOTS * TVX * TVY);
        const float4 xq = b[0 * TVX * TVY], w1q = b[1 * TVX * TVY], w2q = b[2 * TVX * TVY], t4 = b[3 * TVX * TVY];
        const float4 sc4 = b[7 * TVX * TVY];
        float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f), x4 = r4;
        if (need_r) {
            r4 = b[4 * TVX * TVY];
            if (P.hp) {      // the pending half of the last CG update: r <- r - alpha Hp
                const float4 h4 = b[5 * TVX * TVY];
                r4.x = fmaf(-alpha, h4.x, r4.x); r4.y = fmaf(-alpha, h4.y, r4.y);
                r4.z = fmaf(-alpha, h4.z, r4.z); r4.w = fmaf(-alpha, h4.w, r4.w);
            }
        }
        if (P.xtrue) x4 = b[6 * TVX * TVY];
        const float xp[6] = {sc4.x, xq.x, xq.y, xq.z, xq.w, sc4.y};
        const float wc1[5] = {sc4.z, w1q.x, w1q.y, w1q.z, w1q.w};
        const float wc2[5] = {sc4.w, w2q.x, w2q.y, w2q.z, w2q.w};
        const float told[4] = {t4.x, t4.y, t4.z, t4.w};
        const float rc[4] = {r4.x, r4.y, r4.z, r4.w};
        const float xt[4] = {x4.x, x4.y, x4.z, x4.w};
        float dwy_prev, py_prev;
        {
            const float gx = xp[0] - xc[0], gy = xc[1] - xc[0];
            dwy_prev = shrink_dw(gx, gy, wc1[0], wc2[0], kappa).dwy;
            float pxl, mg;
            unit_grad(gx, gy, pxl, py_prev, mg);
        }
        float w1o[4], w2o[4], tvo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float xcv = xc[k + 1];
            const float gx = xp[k + 1] - xcv, gy = xc[k + 2] - xcv;
            const Dw o = shrink_dw(gx, gy, wc1[k + 1], wc2[k + 1], kappa);
            float kt = dwx_up[k];
            kt -= o.dwx;
            kt += dwy_prev;
            kt -= o.dwy;
            float pxo, pyo, mag;
            unit_grad(gx, gy, pxo, pyo, mag);
            tv += mag;
            if (diag) {
                float lap = 0.f;
                lap += xcv - xm[k];
                lap += xcv - xp[k + 1];
                lap += xcv - xc[k];
                lap += xcv - xc[k + 2];
                float dv = 0.f;
                dv += px_up[k] - pxo;
                dv += py_prev - pyo;
                const float gv = told[k] - rc[k] - P.mu * lap + P.lam * dv;
                gn2 = fmaf(gv, gv, gn2);
            }
            if (P.xtrue) { const float e = xcv - xt[k]; img = fmaf(e, e, img); }
            w1o[k] = o.w1; w2o[k] = o.w2; tvo[k] = P.mu * kt;
            dwy_prev = o.dwy; py_prev = pyo;
            dwx_up[k] = o.dwx; px_up[k] = pxo;
        }
        st4(wo1 + g0, make_float4(w1o[0], w1o[1], w1o[2], w1o[3]));
        st4(wo2 + g0, make_float4(w2o[0], w2o[1], w2o[2], w2o[3]));
        st4(tvt + g0, make_float4(tvo[0], tvo[1], tvo[2], tvo[3]));
        if (upd) {
            float rn[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { rn[k] = rc[k] + (tvo[k] - told[k]); rr = fmaf(rn[k], rn[k], rr); }
            st4(P.r_upd + nb + g0, make_float4(rn[0], rn[1], rn[2], rn[3]));
            if (P.p_out) st4(P.p_out + nb + g0, make_float4(rn[0], rn[1], rn[2], rn[3]));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) xm[k] = xc[k + 1];
#pragma unroll
        for (int k = 0; k < 6; ++k) xc[k] = xp[k];
    }
}

template <int PX>
__global__ void __launch_bounds__(TVX * TVY, PX == 4 ? 3 : 4)
tv_fused_kernel(const TvParams P) {
    __shared__ __align__(16) float red[128];
    const int N = P.N;
    const int node = P.node0 + blockIdx.z;
    if (P.masked && !P.ctl[node].active) return;   // a14 retry pass: this node was accepted already
    const long long nb = (long long)blockIdx.z * P.stride;
    const long long n = (long long)N * N;
    const int r0 = (blockIdx.y * TVY + threadIdx.y) * TVROWS;   // first of this thread's TVROWS rows
    const int c0 = (blockIdx.x * TVX + threadIdx.x) * PX;
    const float* __restrict__ x = P.x + nb;
    // device-side ping-pong parity of this node's multiplier (host parity when there is no control table): every block
    // reads it before it arrives at the grid-reduction counter, the last block flips it after all have arrived
    const bool swap = P.ctl && P.ctl[node].wpar;
    const float* __restrict__ w1p = (swap ? P.w_out : P.w_in) + 2 * nb;
    const float* __restrict__ w2p = w1p + n;
    float* __restrict__ wo1 = (swap ? const_cast<float*>(P.w_in) : P.w_out) + 2 * nb;
    float* __restrict__ wo2 = wo1 + n;
    const float kappa = P.lam / P.mu;
    const float alpha = P.hp ? (float)P.scal[(long long)node * NSCAL + S_ALPHA] : 0.f;
    float tv = 0.f, gn2 = 0.f, img = 0.f, rr = 0.f;
    // block-uniform: rows [by*TVY*TVROWS, +TVY*TVROWS) and columns [bx*4*TVX, +4*TVX) all strictly inside the image
    const bool interior = ((N & 3) == 0) && blockIdx.y >= 1 && (int)(blockIdx.y + 1) * TVY * TVROWS <= N - 1 &&
                          blockIdx.x >= 1 && (int)(blockIdx.x + 1) * PX * TVX <= N - 1;
    // the per-row forms work on quads: with PX = 2 every other thread takes one
    const bool quad = (PX == 4) || ((threadIdx.x & 1) == 0);
    extern __shared__ __align__(16) unsigned char tv_smem[];
    if (PX == 4 && interior && P.strip == 2)
        tv_strip_async(P, x, w1p, w2p, wo1, wo2, nb, N, r0, c0, kappa, alpha, tv_smem, tv, gn2, img, rr);
    else if (interior && P.strip) tv_strip<PX>(P, x, w1p, w2p, wo1, wo2, nb, N, r0, c0, kappa, alpha, tv, gn2, img, rr);
    else if (interior) {
        if (quad)
            for (int r = r0; r < r0 + TVROWS; ++r) tv_quad<true>(P, x, w1p, w2p, wo1, wo2, nb, N, r, c0, kappa, alpha, tv, gn2, img, rr);
    } else if (quad && c0 < N) {
        for (int r = r0; r < min(r0 + TVROWS, N); ++r) tv_quad<false>(P, x, w1p, w2p, wo1, wo2, nb, N, r, c0, kappa, alpha, tv, gn2, img, rr);
    }
    float v[4] = {tv, gn2, img, rr};
    block_sum<4>(v, red);
    const int nblk = gridDim.x * gridDim.y, blk = blockIdx.y * gridDim.x + blockIdx.x;
    __shared__ int s_slots[4];
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        s_slots[0] = S_TV; s_slots[1] = S_GN2; s_slots[2] = S_IMG;
        s_slots[3] = P.r_upd ? S_RR0 : S_SCRATCH;
    }
    // (s_slots is ordered before its use by the barriers inside grid_reduce_store)
    const bool last = grid_reduce_store<4>(v, P.part + (long long)blockIdx.z * nblk * 4, P.counter + blockIdx.z, blk,
                                           nblk, P.scal + (long long)node * NSCAL, red, s_slots);
    if (last && P.ctl && threadIdx.x == 0 && threadIdx.y == 0) {
        P.ctl[node].wpar ^= 1;
        // the a14 decision of this node, taken here instead of by a separate launch (thread 0 stored |g|^2 itself)
        if (P.accept)
            accept_node(P.ctl + node, P.scal[(long long)node * NSCAL + S_GN2], P.accept == 1, P.max_tighten,
                        P.eps_target2, P.iter_dev);
    }
}

// =================================================================================================
// K4a: x += alpha p ; r -= alpha Hp ; <r,r> -> scal[rr_out]   (alpha = scal[rr_in] / scal[S_PHP])
// =================================================================================================
__global__ void __launch_bounds__(256)
cg_update_kernel(const CgParams P) {
    __shared__ __align__(16) float red[64];
    const int node = P.node0 + blockIdx.y;
    if (P.ctl && !P.ctl[node].active) return;
    const double* sc = P.scal + (long long)node * NSCAL;
    const double php = sc[S_PHP], rr = sc[P.rr_in];
    const float alpha = (php > 0.0) ? (float)(rr / php) : 0.f;
    const long long nb = (long long)blockIdx.y * P.stride;
    float* __restrict__ x = P.x + nb;
    if (P.x_only) {      // launch-uniform: the r half of the update rides in the TV pass that follows
        if (blockIdx.x == 0 && threadIdx.x == 0) P.scal[(long long)node * NSCAL + S_ALPHA] = (double)alpha;
        const float* __restrict__ pp = P.p + nb;
        const long long m4 = ((P.n & 3) == 0 && (P.stride & 3) == 0) ? (P.n >> 2) : 0;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m4; i += (long long)gridDim.x * blockDim.x) {
            float4 xv = ld4(x + 4 * i);
            const float4 pv = ld4(pp + 4 * i);
            xv.x = fmaf(alpha, pv.x, xv.x); xv.y = fmaf(alpha, pv.y, xv.y);
            xv.z = fmaf(alpha, pv.z, xv.z); xv.w = fmaf(alpha, pv.w, xv.w);
            st4(x + 4 * i, xv);
        }
        for (long long i = 4 * m4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += (long long)gridDim.x * blockDim.x)
            x[i] = fmaf(alpha, pp[i], x[i]);
        return;
    }
    float* r = P.r + nb;
    const float* rin = (P.r_in ? P.r_in : P.r) + nb;
    const float* p = P.p + nb;       // may BE the residual buffer (first CG iteration of the fully fused form: p = r)
    const float* __restrict__ hp = P.hp + nb;
    float s = 0.f;
    const long long n4 = ((P.n & 3) == 0 && (P.stride & 3) == 0) ? (P.n >> 2) : 0;   // float4 only when every node image is 16-byte aligned
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 xv = ld4(x + 4 * i), rv = ld4(rin + 4 * i);
        const float4 pv = ld4(p + 4 * i), hv = ld4(hp + 4 * i);
        xv.x = fmaf(alpha, pv.x, xv.x); xv.y = fmaf(alpha, pv.y, xv.y);
        xv.z = fmaf(alpha, pv.z, xv.z); xv.w = fmaf(alpha, pv.w, xv.w);
        rv.x = fmaf(-alpha, hv.x, rv.x); rv.y = fmaf(-alpha, hv.y, rv.y);
        rv.z = fmaf(-alpha, hv.z, rv.z); rv.w = fmaf(-alpha, hv.w, rv.w);
        st4(x + 4 * i, xv); st4(r + 4 * i, rv);
        s = fmaf(rv.x, rv.x, s); s = fmaf(rv.y, rv.y, s); s = fmaf(rv.z, rv.z, s); s = fmaf(rv.w, rv.w, s);
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += (long long)gridDim.x * blockDim.x) {
        const float xv = fmaf(alpha, p[i], x[i]), rv = fmaf(-alpha, hp[i], rin[i]);
        x[i] = xv; r[i] = rv; s = fmaf(rv, rv, s);
    }
    float v[1] = {s};
    block_sum<1>(v, red);
    grid_reduce_store<1>(v, P.part + (long long)blockIdx.y * gridDim.x, P.counter + blockIdx.y, blockIdx.x,
                         gridDim.x, P.scal + (long long)node * NSCAL + P.rr_out, red);
}

// K4b (unfused fallback of the direction update; the fused form lives in the forward projector's staging)
__global__ void __launch_bounds__(256)
p_update_kernel(const CgParams P) {
    const int node = P.node0 + blockIdx.y;
    if (P.ctl && !P.ctl[node].active) return;
    const double* sc = P.scal + (long long)node * NSCAL;
    const double den = sc[P.rr_in], num = sc[P.rr_out];
    const float beta = (den > 0.0) ? (float)(num / den) : 0.f;
    const long long nb = (long long)blockIdx.y * P.stride;
    const float* __restrict__ r = P.r + nb;
    const float* __restrict__ p = P.p + nb;
    float* __restrict__ po = P.p_out + nb;
    const long long n4 = ((P.n & 3) == 0 && (P.stride & 3) == 0) ? (P.n >> 2) : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = ld4(p + 4 * i), b = ld4(r + 4 * i);
        st4(po + 4 * i, make_float4(fmaf(beta, a.x, b.x), fmaf(beta, a.y, b.y), fmaf(beta, a.z, b.z), fmaf(beta, a.w, b.w)));
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += (long long)gridDim.x * blockDim.x)
        po[i] = fmaf(beta, p[i], r[i]);
}

// Ax += alpha(node) * q on sinogram rows (keeps A x current without an extra projection; block_6_ver2:190-194)
__global__ void __launch_bounds__(256)
sino_axpy_kernel(const SinoParams P) {
    const int a = P.A0 + blockIdx.x;            // angle rows on grid.x
    const int node = P.anode[a];
    if (P.ctl && !P.ctl[node].active) return;
    const double* sc = P.scal + (long long)node * NSCAL;
    float alpha = 1.f;
    if (P.mode == 1) {
        const double php = sc[S_PHP], rr = sc[P.rr_in];
        alpha = (php > 0.0) ? (float)(rr / php) : 0.f;
    }
    for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < P.D; j += gridDim.y * blockDim.x) {
        const long long g = (long long)a * P.D + j;
        P.ax[g] = (P.mode == 0) ? P.q[g] : fmaf(alpha, P.q[g], P.ax[g]);
    }
}

// |A x - b|^2 per node -> scal[S_MSE]   (one block per node: deterministic)
__global__ void __launch_bounds__(256)
sino_resid_kernel(const SinoParams P) {
    __shared__ __align__(16) float red[32];
    __shared__ double dred[8];
    const int node = P.node0 + blockIdx.x;
    const long long beg = (long long)P.aptr[node] * P.D, end = (long long)P.aptr[node + 1] * P.D;
    double acc = 0.0;
    for (long long g = beg + threadIdx.x; g < end; g += blockDim.x) {
        const float e = P.ax[g] - P.b[g];
        acc += (double)e * (double)e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) dred[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += dred[w];
        P.scal[(long long)node * NSCAL + S_MSE] = s;
    }
    (void)red;
}

// =================================================================================================
// K6: rhs0_i = A^T P b_i + rho * sum_{j in N(i)} Q_ij .* (z_ij - y_ij,i)   in G.neighbors(i) order
// (block_6_admm_loop_ver2.py:87-95 neighbour assembly; block_5_node_problem.py:24-27 normal equations)
// =================================================================================================
__global__ void __launch_bounds__(256)
rhs0_kernel(const RhsParams P) {
    __shared__ __align__(16) float red[32];
    const int node = P.node0 + blockIdx.y;
    const long long nb = (long long)blockIdx.y * P.stride;
    const int kb = P.nbr_ptr[node], ke = P.nbr_ptr[node + 1];
    const float* __restrict__ atb = P.atb + nb;
    float* out = P.rhs0 + nb;
    const bool upd = (P.r_upd != nullptr);
    float* rio = upd ? P.r_upd + nb : nullptr;
    float* po = (upd && P.p_out) ? P.p_out + nb : nullptr;
    float rr = 0.f;
    const long long n4 = ((P.n & 3) == 0 && (P.stride & 3) == 0) ? (P.n >> 2) : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 cons = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = kb; k < ke; ++k) {
            const float4 z = ld4(reinterpret_cast<const float*>(P.nbr_z[k]) + 4 * i);
            const float4 y = ld4(reinterpret_cast<const float*>(P.nbr_y[k]) + 4 * i);
            const float* q = reinterpret_cast<const float*>(P.nbr_q[k]);
            float4 qv = make_float4(P.q_uniform, P.q_uniform, P.q_uniform, P.q_uniform);
            if (q) qv = ld4(q + 4 * i);
            cons.x += P.rho * qv.x * (z.x - y.x); cons.y += P.rho * qv.y * (z.y - y.y);
            cons.z += P.rho * qv.z * (z.z - y.z); cons.w += P.rho * qv.w * (z.w - y.w);
        }
        const float4 a = ld4(atb + 4 * i);
        const float4 nw = make_float4(a.x + cons.x, a.y + cons.y, a.z + cons.z, a.w + cons.w);
        if (upd) {
            const float4 od = ld4(out + 4 * i), rv = ld4(rio + 4 * i);
            const float4 rn = make_float4(rv.x + (nw.x - od.x), rv.y + (nw.y - od.y), rv.z + (nw.z - od.z), rv.w + (nw.w - od.w));
            st4(rio + 4 * i, rn);
            if (po) st4(po + 4 * i, rn);
            rr = fmaf(rn.x, rn.x, rr); rr = fmaf(rn.y, rn.y, rr); rr = fmaf(rn.z, rn.z, rr); rr = fmaf(rn.w, rn.w, rr);
        }
        st4(out + 4 * i, nw);
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += (long long)gridDim.x * blockDim.x) {
        float cons = 0.f;
        for (int k = kb; k < ke; ++k) {
            const float* z = reinterpret_cast<const float*>(P.nbr_z[k]);
            const float* y = reinterpret_cast<const float*>(P.nbr_y[k]);
            const float* q = reinterpret_cast<const float*>(P.nbr_q[k]);
            const float qv = q ? q[i] : P.q_uniform;
            cons += P.rho * qv * (z[i] - y[i]);
        }
        const float nw = atb[i] + cons;
        if (upd) {
            const float rn = rio[i] + (nw - out[i]);
            rio[i] = rn; if (po) po[i] = rn; rr = fmaf(rn, rn, rr);
        }
        out[i] = nw;
    }
    if (upd) {   // launch-uniform
        float v[1] = {rr};
        block_sum<1>(v, red);
        grid_reduce_store<1>(v, P.part + (long long)blockIdx.y * gridDim.x, P.counter + blockIdx.y, blockIdx.x, gridDim.x,
                             P.scal + (long long)node * NSCAL + S_RR0, red);
    }
}

// =================================================================================================
// K5: one edge (i, j), i < j   (block_6_admm_loop_ver2.py:210-264)
//   a_i = x_i + y_i, a_j = x_j + y_j          (:217-218)   [remote end: a arrives packed by NCCL recv]
//   z'  = (a_i + a_j) / 2                      (:221-223)   or (W_i a_i + W_j a_j)/(W_i + W_j) (PDF eq. 2)
//   y_i' = y_i + x_i - z', y_j' likewise       (:229-230)
//   sums: |x_i - z'|^2, |x_j - z'|^2, |z' - z|^2 (:240-249), and the block_5 penalty
//         sum q_ij (x_i - (z - y_i))^2 evaluated with the OLD z, y (objective value, block_5:24-27).
// =================================================================================================
struct EdgePtrs {
    const float *xi, *xj, *ai_r, *aj_r, *Wi, *Wj, *qij, *qji;
    float *yi, *yj, *z, *vi, *vj;
    float q_uniform;
};

__device__ __forceinline__ void edge_elem(const EdgePtrs& e, float zo, float xiv, float yiv, float xjv, float yjv,
                                          float air, float ajr, float wi, float wj, float qi, float qj, float& zn,
                                          float& yin, float& yjn, float (&s)[5]) {
    float ai, aj;
    if (e.xi) {
        ai = xiv + yiv;
        const float ei = xiv - (zo - yiv);
        s[3] = fmaf(qi * ei, ei, s[3]);
    } else ai = air;
    if (e.xj) {
        aj = xjv + yjv;
        const float ej = xjv - (zo - yjv);
        s[4] = fmaf(qj * ej, ej, s[4]);
    } else aj = ajr;
    if (e.Wi) zn = (wi * ai + wj * aj) / (wi + wj);
    else zn = (ai + aj) / 2.0f;
    if (e.xi) { const float ri = xiv - zn; yin = yiv + xiv - zn; s[0] = fmaf(ri, ri, s[0]); }
    if (e.xj) { const float rj = xjv - zn; yjn = yjv + xjv - zn; s[1] = fmaf(rj, rj, s[1]); }
    const float dz = zn - zo;
    s[2] = fmaf(dz, dz, s[2]);
}

__global__ void __launch_bounds__(256, 4)
edge_kernel(const EdgeParams P) {
    __shared__ __align__(16) float red[160];
    const EdgeDesc d = P.edges[blockIdx.y];
    EdgePtrs e;
    e.xi = reinterpret_cast<const float*>(d.xi); e.xj = reinterpret_cast<const float*>(d.xj);
    e.yi = reinterpret_cast<float*>(d.yi); e.yj = reinterpret_cast<float*>(d.yj);
    e.z = reinterpret_cast<float*>(d.z);
    e.ai_r = reinterpret_cast<const float*>(d.ai); e.aj_r = reinterpret_cast<const float*>(d.aj);
    e.Wi = reinterpret_cast<const float*>(d.Wi); e.Wj = reinterpret_cast<const float*>(d.Wj);
    e.qij = reinterpret_cast<const float*>(d.qij); e.qji = reinterpret_cast<const float*>(d.qji);
    e.vi = reinterpret_cast<float*>(d.vi); e.vj = reinterpret_cast<float*>(d.vj);
    e.q_uniform = P.q_uniform;
    float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 qu4 = make_float4(P.q_uniform, P.q_uniform, P.q_uniform, P.q_uniform);
    const long long n4 = ((P.n & 3) == 0) ? (P.n >> 2) : 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += (long long)gridDim.x * blockDim.x) {
        const long long o = 4 * k;
        const float4 zo = ld4(e.z + o);
        const float4 xi = e.xi ? ld4(e.xi + o) : zero4, yi = e.xi ? ld4(e.yi + o) : zero4;
        const float4 xj = e.xj ? ld4(e.xj + o) : zero4, yj = e.xj ? ld4(e.yj + o) : zero4;
        const float4 ar = e.xi ? zero4 : ld4(e.ai_r + o), br = e.xj ? zero4 : ld4(e.aj_r + o);
        const float4 wi = e.Wi ? ld4(e.Wi + o) : zero4, wj = e.Wi ? ld4(e.Wj + o) : zero4;
        const float4 qi = (e.xi && e.qij) ? ld4(e.qij + o) : qu4, qj = (e.xj && e.qji) ? ld4(e.qji + o) : qu4;
        float4 zn, yin = zero4, yjn = zero4;
        edge_elem(e, zo.x, xi.x, yi.x, xj.x, yj.x, ar.x, br.x, wi.x, wj.x, qi.x, qj.x, zn.x, yin.x, yjn.x, s);
        edge_elem(e, zo.y, xi.y, yi.y, xj.y, yj.y, ar.y, br.y, wi.y, wj.y, qi.y, qj.y, zn.y, yin.y, yjn.y, s);
        edge_elem(e, zo.z, xi.z, yi.z, xj.z, yj.z, ar.z, br.z, wi.z, wj.z, qi.z, qj.z, zn.z, yin.z, yjn.z, s);
        edge_elem(e, zo.w, xi.w, yi.w, xj.w, yj.w, ar.w, br.w, wi.w, wj.w, qi.w, qj.w, zn.w, yin.w, yjn.w, s);
        st4(e.z + o, zn);
        if (e.xi) st4(e.yi + o, yin);
        if (e.xj) st4(e.yj + o, yjn);
        // single-owner exchange: v = z' - y' of a peer's node goes straight into the peer's memory (posted NVLink stores)
        if (e.vi) st4(e.vi + o, make_float4(zn.x - yin.x, zn.y - yin.y, zn.z - yin.z, zn.w - yin.w));
        if (e.vj) st4(e.vj + o, make_float4(zn.x - yjn.x, zn.y - yjn.y, zn.z - yjn.z, zn.w - yjn.w));
    }
    for (long long k = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; k < P.n; k += (long long)gridDim.x * blockDim.x) {
        float zn, yin = 0.f, yjn = 0.f;
        edge_elem(e, e.z[k], e.xi ? e.xi[k] : 0.f, e.xi ? e.yi[k] : 0.f, e.xj ? e.xj[k] : 0.f, e.xj ? e.yj[k] : 0.f,
                  e.xi ? 0.f : e.ai_r[k], e.xj ? 0.f : e.aj_r[k], e.Wi ? e.Wi[k] : 0.f, e.Wi ? e.Wj[k] : 0.f,
                  (e.xi && e.qij) ? e.qij[k] : P.q_uniform, (e.xj && e.qji) ? e.qji[k] : P.q_uniform, zn, yin, yjn, s);
        e.z[k] = zn;
        if (e.xi) e.yi[k] = yin;
        if (e.xj) e.yj[k] = yjn;
        if (e.vi) e.vi[k] = zn - yin;
        if (e.vj) e.vj[k] = zn - yjn;
    }
    block_sum<5>(s, red);
    grid_reduce_store<5>(s, P.part + (long long)blockIdx.y * gridDim.x * 5, P.counter + blockIdx.y, blockIdx.x,
                         gridDim.x, P.sums + (long long)blockIdx.y * 5, red);
}

// a = x + y for the cut-edge ends this rank sends
__global__ void __launch_bounds__(256)
pack_kernel(const PackParams P) {
    // grid.y == nitems: one item per block row; grid.y == 1 (narrow launch, `out` in a peer's memory): every block
    // walks all items, so a few resident blocks keep the NVLink stores in flight next to the HBM-bound kernels
    for (int it = blockIdx.y; it < P.nitems; it += gridDim.y) {
        const PackDesc d = P.items[it];
        const float* __restrict__ x = reinterpret_cast<const float*>(d.x);
        const float* __restrict__ y = reinterpret_cast<const float*>(d.y);
        float* __restrict__ o = reinterpret_cast<float*>(d.out);
        const long long n4 = ((P.n & 3) == 0) ? (P.n >> 2) : 0;
        const long long step = (long long)gridDim.x * blockDim.x;
        if (!y) {   // single-owner exchange: the peer that updates the edge gets x itself
#pragma unroll 4
            for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += step) st4(o + 4 * k, ld4(x + 4 * k));
            for (long long k = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; k < P.n; k += step) o[k] = x[k];
            continue;
        }
#pragma unroll 4
        for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += step) {
            const float4 a = ld4(x + 4 * k), b = ld4(y + 4 * k);
            st4(o + 4 * k, make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w));
        }
        for (long long k = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; k < P.n; k += step)
            o[k] = x[k] + y[k];
    }
}

// Per-iteration bookkeeping (block_6_admm_loop_ver2.py:232-264): fold the per-edge sums (in G.edges()
// order, fp64, like the reference's Python floats) and the per-node scalars into one history row
//   row = [r2, s2, pri_node[Vg], dual_node[Vg], pen[Vg], mse[Vg], tv[Vg], gn2[Vg], img[Vg], tighten_tries[Vg]]
// Rows of different ranks are summed by the caller (ncclAllReduce) when nodes are sharded.
__global__ void __launch_bounds__(256) finalize_kernel(const FinalizeParams P) {
    __shared__ double red2[2][8];
    const int Vg = P.Vg, tid = threadIdx.x;
    double* row = P.iter_dev ? P.row + (long long)(*P.iter_dev) * P.hist_stride : P.row;
    double* pri = row + 2; double* dual = pri + Vg; double* pen = dual + Vg;
    double* mse = pen + Vg; double* tv = mse + Vg; double* gn2 = tv + Vg; double* img = gn2 + Vg;
    double* tries = img + Vg;
    const double rho2 = (double)P.rho * (double)P.rho;
    for (int i = tid; i < 2 + 8 * Vg; i += blockDim.x) row[i] = 0.0;
    __syncthreads();
    // per local node, over its incident edges in G.neighbors() order
    for (int v = tid; v < P.V; v += blockDim.x) {
        const int g = P.node_gid[v];
        double a = 0.0, b = 0.0, d = 0.0;
        for (int k = P.nbr_ptr[v]; k < P.nbr_ptr[v + 1]; ++k) {
            const int e = P.nbr_epos[k], end = P.nbr_end[k];
            if (e < 0) continue;      // updated by the peer that owns the edge: it adds this node's pieces to its row
            const double* s = P.sums + (long long)e * 5;
            a += s[end];
            b += s[3 + end];
            if (P.edge_flags[e] & 4) d += rho2 * s[2];
        }
        const double* sc = P.scal + (long long)v * NSCAL;
        pri[g] = a; pen[g] = b; dual[g] = d;
        mse[g] = sc[S_MSE]; tv[g] = sc[S_TV]; gn2[g] = sc[S_GN2]; img[g] = sc[S_IMG];
        tries[g] = P.ctl ? (double)P.ctl[v].tries : 0.0;
    }
    // totals: fixed-order strided partials + fixed tree
    double r2 = 0.0, s2 = 0.0;
    for (int e = tid; e < P.E; e += blockDim.x) {
        const double* s = P.sums + (long long)e * 5;
        const int fl = P.edge_flags[e];
        if (fl & 1) r2 += s[0];
        if (fl & 2) r2 += s[1];
        if (fl & 4) s2 += rho2 * s[2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((tid & 31) == 0) { red2[0][tid >> 5] = r2; red2[1][tid >> 5] = s2; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red2[0][w]; b += red2[1][w]; }
        row[0] = a; row[1] = b;
        // dual residual of owned cut edges also belongs to the REMOTE end's node (block_6_ver2:251-253)
        for (int e = P.E_local; e < P.E; ++e) {
            const int fl = P.edge_flags[e];
            const double* s = P.sums + (long long)e * 5;
            if (fl & 24) {            // single-owner exchange: every per-node piece of the remote end
                const int g = (fl & 8) ? P.edge_gi[e] : P.edge_gj[e], end = (fl & 8) ? 0 : 1;
                pri[g] += s[end]; pen[g] += s[3 + end]; dual[g] += rho2 * s[2];
            } else if ((fl & 4) && !(fl & 2)) dual[P.edge_gj[e]] += rho2 * s[2];
        }
    }
    if (P.iter_dev) {      // every thread has read the counter (row pointer) before the barrier above
        __syncthreads();
        if (tid == 0) *P.iter_dev += 1;
    }
}

// a14 acceptance (block_6_admm_loop_ver2.py:155-176), one thread per node, after a solve + TV pass left |g_x,i|^2 in
// scal[S_GN2]:  accept if |g| <= eps_target (:155) or the tighten cap is reached (:164); else tighten and retry (:175).
// `first`: this is the decision after the iteration's first solve (every node took part in it).
__global__ void __launch_bounds__(128) accept_kernel(const AcceptParams P) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= P.nodes) return;
    accept_node(P.ctl + P.node0 + v, P.scal[(long long)(P.node0 + v) * NSCAL + S_GN2], P.first != 0, P.max_tighten,
                P.eps_target2, P.iter_dev);
}

// ---- launchers --------------------------------------------------------------------------------------
static inline int stream_blocks(long long n, int per_thread) {
    long long b = (n + 256LL * per_thread - 1) / (256LL * per_thread);
    if (b < 1) b = 1;
    if (b > 4096) b = 4096;
    return (int)b;
}

cudaError_t launch_tv(const TvParams& P0, int nodes, cudaStream_t st) {
    // ADMM_B200_TVSTRIP: 0 per-row form, 1 row march with register loads, 2 (default) row march with cp.async staging
    static const int strip = [] { const char* e = getenv("ADMM_B200_TVSTRIP"); return e ? atoi(e) : 2; }();
    TvParams P = P0;
    P.strip = strip;
    // pixels per thread and row: 2 for large images (twice the resident warps), 4 otherwise; ADMM_B200_TVPX overrides
    static const int px_env = [] { const char* e = getenv("ADMM_B200_TVPX"); return e ? atoi(e) : 0; }();
    const int px = (px_env == 2 || px_env == 4) ? px_env : 4;
    dim3 grid((P.N + px * TVX - 1) / (px * TVX), (P.N + TVY * TVROWS - 1) / (TVY * TVROWS), nodes);
    size_t smem = 0;
    if (px == 4 && P.strip == 2) {
        // the dynamic shared-memory limit is a per-device function attribute
        static std::mutex mu;
        static bool done[64] = {false};
        std::lock_guard<std::mutex> lk(mu);
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64 || !done[dev]) {
            e = cudaFuncSetAttribute(tv_fused_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTvAsyncSmem);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) done[dev] = true;
        }
        smem = kTvAsyncSmem;
    }
    ProfScope ps(KC_TV, st);
    if (px == 2) tv_fused_kernel<2><<<grid, dim3(TVX, TVY), 0, st>>>(P);
    else tv_fused_kernel<4><<<grid, dim3(TVX, TVY), smem, st>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_cg_update(const CgParams& P, int nodes, int nblk, cudaStream_t st) {
    { ProfScope ps(KC_CG_UPDATE, st); cg_update_kernel<<<dim3(nblk, nodes), 256, 0, st>>>(P); }
    return cudaGetLastError();
}
cudaError_t launch_p_update(const CgParams& P, int nodes, cudaStream_t st) {
    { ProfScope ps(KC_P_UPDATE, st); p_update_kernel<<<dim3(stream_blocks(P.n, 8), nodes), 256, 0, st>>>(P); }
    return cudaGetLastError();
}
cudaError_t launch_sino_axpy(const SinoParams& P, cudaStream_t st) {
    if (P.A1 <= P.A0) return cudaSuccess;
    { ProfScope ps(KC_SINO_AXPY, st); sino_axpy_kernel<<<dim3(P.A1 - P.A0, (P.D + 255) / 256), 256, 0, st>>>(P); }
    return cudaGetLastError();
}
cudaError_t launch_sino_resid(const SinoParams& P, int nodes, cudaStream_t st) {
    { ProfScope ps(KC_SINO_RESID, st); sino_resid_kernel<<<nodes, 256, 0, st>>>(P); }
    return cudaGetLastError();
}
cudaError_t launch_rhs0(const RhsParams& P, int nodes, cudaStream_t st) {
    { ProfScope ps(KC_RHS0, st); rhs0_kernel<<<dim3(stream_blocks(P.n, 4), nodes), 256, 0, st>>>(P); }
    return cudaGetLastError();
}
cudaError_t launch_edges(const EdgeParams& P, int nedges, int nblk, cudaStream_t st) {
    if (nedges <= 0) return cudaSuccess;
    { ProfScope ps(KC_EDGE, st); edge_kernel<<<dim3(nblk, nedges), 256, 0, st>>>(P); }
    return cudaGetLastError();
}
cudaError_t launch_pack(const PackParams& P, int nitems, int narrow_blocks, cudaStream_t st) {
    if (nitems <= 0) return cudaSuccess;
    const dim3 grid = narrow_blocks > 0 ? dim3(narrow_blocks, 1) : dim3(stream_blocks(P.n, 4), nitems);
    { ProfScope ps(KC_PACK, st); pack_kernel<<<grid, 256, 0, st>>>(P); }
    return cudaGetLastError();
}
cudaError_t launch_accept(const AcceptParams& P, cudaStream_t st) {
    if (P.nodes <= 0) return cudaSuccess;
    { ProfScope ps(KC_ACCEPT, st); accept_kernel<<<(P.nodes + 127) / 128, 128, 0, st>>>(P); }
    return cudaGetLastError();
}
cudaError_t launch_finalize(const FinalizeParams& P, cudaStream_t st) {
    { ProfScope ps(KC_FINALIZE, st); finalize_kernel<<<1, 256, 0, st>>>(P); }
    return cudaGetLastError();
}

}  // namespace admm
