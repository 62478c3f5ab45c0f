// projector.cu -- K1 (Joseph forward projector, strip-staged, deterministic), K2 (matched gather
// back-projector with fused H-apply / CG-residual epilogues), K2b (column norms^2).  sm_100a.
//
// Replaces odl.tomo.RayTransform.__call__ / `Ai @ x` (block_2_load_odl_data.py:149,
// block_6_admm_loop_ver2.py:145,193), `Ai.T @ r` (block_6_admm_loop_ver2.py:145) and
// np.sum(A_i*A_i, axis=0) (block_3_graph_and_precisions.py:22).  Discretisation: SURVEY.md App. C.
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "projector.cuh"

namespace admm {

// =================================================================================================
// K1 forward.  One block = (node, orientation, strip ti, segment sg, angle chunk).  The block walks its
// segment slab by slab: the slab's image tile is staged into shared memory in the canonical layout
// S[step k][interp u] (x-dominant angles: k = iy, u = ix, i.e. transposed on the way in; y-dominant:
// k = ix, u = iy), with zero halo columns so that only pixels the tile OWNS contribute.  Thread (slot,
// t) owns detector bin jmin+t of one angle and accumulates its L-step partial line integral in a
// register, then adds it to the block's per-angle accumulator row in shared memory (one owner per bin
// per slab -> no atomics).  At the end the rows are stored as fixed-size records; fwd_reduce_kernel sums
// the records of all strips/segments in a fixed order (deterministic) and applies the step weight.
// TMA (tma.cuh): in a plain projection the y-dominant tile -- whose canonical layout is the image's own row-major
// order -- lands by one cp.async.bulk.tensor.3d per slab (zero-filled outside the image, mbarrier arrival); while a slab
// is sampled the next slab's boxes of every operand stream are prefetched to L2 (cp.async.bulk.prefetch.tensor).
// x-dominant tiles (transposed on the way in) and the CG-fused mode (updated on the way in) are staged by the threads.
// =================================================================================================
__global__ void __launch_bounds__(FTHREADS, 4)
fwd_strip_kernel(const __grid_constant__ FwdParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* S = reinterpret_cast<float*>(smem_raw);                 // [FL][FPITCH]  ([FL][FPITCH_T] when landed by TMA)
    float* acc_s = S + FL * FPITCH_T;                              // [FAC][span]
    __shared__ __align__(8) unsigned long long s_bar;              // TMA arrival barrier of the tile copy
    double* s_base = reinterpret_cast<double*>(acc_s + FAC * P.span + ((FAC * P.span) & 1));  // [FAC]
    int* s_jseg = reinterpret_cast<int*>(s_base + FAC);            // [FAC]
    int* s_aid = s_jseg + FAC;                                     // [FAC]
    int4* s_win = reinterpret_cast<int4*>(s_aid + FAC);            // [FAC] per slab: jmin, jmax, jb, fb(bits)
    float4* s_ang = reinterpret_cast<float4*>(s_win + FAC);        // [FAC] inv_major, slope, inv_slope, -

    const int node = P.node0 + blockIdx.y;
    if (P.ctl && !P.ctl[node].active) return;   // a14 retry pass: this node was accepted already
    const int orient = blockIdx.z;  // 0: x-dominant list, 1: y-dominant list
    const int nTi = P.nTi, nSeg = P.nSeg, N = P.N, D = P.D, span = P.span;
    const int bx = blockIdx.x;
    const int ti = bx % nTi, sg = (bx / nTi) % nSeg, ch = bx / (nTi * nSeg);
    const int obeg = P.optr[orient * (P.V + 1) + node], oend = P.optr[orient * (P.V + 1) + node + 1];
    const int a0 = obeg + ch * FAC;
    if (a0 >= oend) return;
    const int na = min(FAC, oend - a0);
    const bool xdom = (orient == 0);
    // y-dominant tiles of a plain projection land by ONE bulk tensor copy per slab (pitch FPITCH_T, zero-filled outside
    // the image); everything else is staged by the threads (pitch FPITCH, transposed / CG-updated on the way in)
    const bool ytma = !xdom && P.mode == 0 && (P.use_tma & 2);
    unsigned bar_phase = 0;
    if (ytma && threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
    }

    const float* __restrict__ img = P.img + (long long)blockIdx.y * P.img_stride;
    const int U0 = ti * FW, Wt = min(FW, N - U0);
    const int K0seg = sg * P.seg, Kseg = min(P.seg, N - K0seg);
    const int tid = threadIdx.x;
    const double cx = 0.5 * (N - 1), cj = 0.5 * (D - 1);

    for (int i = tid; i < FAC * span; i += FTHREADS) acc_s[i] = 0.f;
    double my_M = 0.0, my_m = 0.0;   // threads < na keep their angle's (major, minor) tau steps
    if (tid < na) {
        const int aid = P.oidx[a0 + tid];
        const AngleRec r = P.ang[aid];
        my_M = xdom ? r.ct : r.st;
        my_m = xdom ? r.st : r.ct;
        const double base = cj + (U0 - cx) * my_M + (K0seg - cx) * my_m;  // tau of pixel (u=0,k=0) of the segment
        const double e1 = (Wt - 1) * my_M, e2 = (Kseg - 1) * my_m;
        const double lo = base + fmin(e1, 0.0) + fmin(e2, 0.0);
        s_base[tid] = base;
        s_jseg[tid] = (int)ceil(lo - fabs(my_M)) - 1;   // one bin of slack below the exact bound
        s_aid[tid] = aid;
        s_ang[tid] = make_float4(r.inv_major, r.slope, r.inv_slope, 0.f);
    }

    // fused CG vector updates (see FwdParams): mode 1: p' = r + beta p ; mode 2: x += alpha p, r' = r - alpha Hp,
    // p' = r' + beta p.  Exactly one (orientation, chunk) pass per pixel -- the "writer" -- stores the results.
    __shared__ __align__(16) float red[64];
    float beta = 0.f, alpha = 0.f, rsum = 0.f;
    const int mode = P.mode;
    const float* __restrict__ rimg = nullptr;
    const float* __restrict__ hpimg = nullptr;
    float* pout = nullptr;
    float* xio = nullptr;
    float* rout = nullptr;
    bool writer = false;
    if (mode != 0) {
        const double* sc = P.scal + (long long)node * NSCAL;
        if (mode == 1) {
            const double den = sc[P.beta_den], num = sc[P.beta_num];
            beta = (den > 0.0) ? (float)(num / den) : 0.f;
        } else {
            // <r',r'> = rr - 2 alpha <r,Hp> + alpha^2 <Hp,Hp>, and <r,Hp> = <p,Hp> for CG directions (p - r is a
            // multiple of the previous direction, which is H-conjugate to p), so  rr' = alpha^2 <Hp,Hp> - rr.
            const double rr = sc[P.beta_den], php = sc[S_PHP], hh = sc[S_HPHP];
            const double al = (php > 0.0) ? rr / php : 0.0;
            const double rr_est = al * al * hh - rr;
            alpha = (float)al;
            beta = (rr > 0.0) ? (float)fmax(0.0, rr_est / rr) : 0.f;
        }
        const long long off = (long long)blockIdx.y * P.img_stride;
        rimg = P.r + off;
        const int x_has = P.optr[node + 1] - P.optr[node];
        writer = (ch == 0) && (xdom ? true : (x_has == 0));
        if (writer) pout = P.p_out + off;
        if (mode == 2) {
            hpimg = P.hp + off;
            if (writer) { xio = P.x_io + off; rout = P.r_out + off; }
        }
    }
    // p_old (and r, Hp, x) at flat index g -> the value to project; stores the updates when this block is the writer
    auto fuse1 = [&](long long g, float pold) -> float {
        if (mode == 0) return pold;
        float rv = rimg[g];
        if (mode == 2) {
            rv = fmaf(-alpha, hpimg[g], rv);
            if (writer) { xio[g] = fmaf(alpha, pold, xio[g]); rout[g] = rv; rsum = fmaf(rv, rv, rsum); }
        }
        const float pn = fmaf(beta, pold, rv);
        if (writer) pout[g] = pn;
        return pn;
    };
    // vector form split into (all loads) / (update + stores) so that two items' loads are in flight together
    struct Item4 { float4 p, r, h, x; };
    auto load4 = [&](long long g) -> Item4 {
        Item4 it;
        it.p = ld4(img + g);
        if (mode != 0) it.r = ld4(rimg + g);
        if (mode == 2) { it.h = ld4(hpimg + g); if (writer) it.x = ld4(xio + g); }
        return it;
    };
    auto finish4 = [&](long long g, const Item4& it) -> float4 {
        if (mode == 0) return it.p;
        float4 rv = it.r;
        if (mode == 2) {
            rv.x = fmaf(-alpha, it.h.x, rv.x); rv.y = fmaf(-alpha, it.h.y, rv.y);
            rv.z = fmaf(-alpha, it.h.z, rv.z); rv.w = fmaf(-alpha, it.h.w, rv.w);
            if (writer) {
                float4 xv = it.x;
                xv.x = fmaf(alpha, it.p.x, xv.x); xv.y = fmaf(alpha, it.p.y, xv.y);
                xv.z = fmaf(alpha, it.p.z, xv.z); xv.w = fmaf(alpha, it.p.w, xv.w);
                st4(xio + g, xv);
                st4(rout + g, rv);
                rsum = fmaf(rv.x, rv.x, rsum); rsum = fmaf(rv.y, rv.y, rsum);
                rsum = fmaf(rv.z, rv.z, rsum); rsum = fmaf(rv.w, rv.w, rsum);
            }
        }
        float4 pn;
        pn.x = fmaf(beta, it.p.x, rv.x); pn.y = fmaf(beta, it.p.y, rv.y);
        pn.z = fmaf(beta, it.p.z, rv.z); pn.w = fmaf(beta, it.p.w, rv.w);
        if (writer) st4(pout + g, pn);
        return pn;
    };
    const bool vec4 = ((N & 3) == 0);

    const int slot = tid / FTPA, t = tid % FTPA;
    const int nslab = (Kseg + FL - 1) / FL;
    // shared-window byte address of S[0][-1+1] minus the magic-number bias: addr(k, i) = sbase + k*pitch*4 + bits(i)*4
    const unsigned sbase = smem_u32(S) + 4u * (ytma ? FHALO_T : FHALO) - ((unsigned)kMagicBits << 2);
    for (int slab = 0; slab < nslab; ++slab) {
        const int K0 = K0seg + slab * FL, Lt = min(FL, K0seg + Kseg - K0);
        __syncthreads();  // previous slab fully consumed (also orders the acc/zero + setup writes)
        // ---- per-angle detector window of this slab (one thread per angle) ------------------------------
        if (tid < na) {
            const double base = s_base[tid] + (double)(slab * FL) * my_m;  // tau of slab pixel (0,0)
            const double e1 = (Wt - 1) * my_M, e2 = (Lt - 1) * my_m, om = fabs(my_M);
            const double tlo = base + fmin(e1, 0.0) + fmin(e2, 0.0) - om;
            const double thi = base + fmax(e1, 0.0) + fmax(e2, 0.0) + om;
            const double fbase = floor(base);
            s_win[tid] = make_int4(max(0, (int)ceil(tlo)), min(D - 1, (int)floor(thi)), (int)fbase,
                                   __float_as_int((float)(base - fbase)));
        }
        if (ytma && tid == 0) {
            // the previous slab's reads (and this thread's halo stores) are ordered before the copy unit's writes
            fence_proxy_async();
            mbar_expect_tx(&s_bar, (unsigned)(FL * FPITCH_T * sizeof(float)));
            tma_load_3d(S, &P.mapY[0], U0 - FHALO_T, K0, (int)blockIdx.y, &s_bar);
        }
        // ---- stage tile -------------------------------------------------------------------------
        // x-dominant: pixel (ix = U0+u, iy = K0+k) is contiguous along the step axis k -> item = (u, 4 consecutive k),
        //             stored transposed (stride FPITCH, odd -> conflict-free up to 2-way);
        // y-dominant: pixel (ix = K0+k, iy = U0+u) is contiguous along the interpolation axis u -> item = (k, 4 u).
        if (!ytma) {
            const int per_row = xdom ? (FL / 4) : (FW / 4);            // float4 items per contiguous run
            const int nitems = xdom ? FW * (FL / 4) : FL * (FW / 4);
            auto decode = [&](int idx, int& row, int& c4, bool& full, bool& any, long long& g) {
                row = idx / per_row;                                    // u (x-dom) or k (y-dom)
                c4 = (idx % per_row) * 4;                               // k4 (x-dom) or u4 (y-dom)
                const int rlim = xdom ? Wt : Lt, clim = xdom ? Lt : Wt;
                any = (idx < nitems) && row < rlim && c4 < clim;
                full = any && vec4 && (c4 + 3 < clim);
                g = xdom ? (long long)(U0 + row) * N + K0 + c4 : (long long)(K0 + row) * N + U0 + c4;
            };
            auto store = [&](int row, int c4, float4 val) {
                if (xdom) {
                    float* dst = S + c4 * FPITCH + FHALO + row;
                    dst[0] = val.x; dst[FPITCH] = val.y; dst[2 * FPITCH] = val.z; dst[3 * FPITCH] = val.w;
                } else {
                    float* dst = S + row * FPITCH + FHALO + c4;
                    dst[0] = val.x; dst[1] = val.y; dst[2] = val.z; dst[3] = val.w;
                }
            };
            auto tail = [&](long long g, int c4, bool any) -> float4 {   // ragged / unaligned items, element by element
                float tmp[4] = {0.f, 0.f, 0.f, 0.f};
                if (any) {
                    const int clim = xdom ? Lt : Wt;
                    for (int i = 0; i < 4; ++i)
                        if (c4 + i < clim) tmp[i] = fuse1(g + i, img[g + i]);
                }
                return make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
            };
            for (int idx = tid; idx < nitems; idx += FTHREADS) {
                int rowA, cA;
                bool fullA, anyA;
                long long gA;
                decode(idx, rowA, cA, fullA, anyA, gA);
                float4 va;
                if (fullA) { const Item4 a = load4(gA); va = finish4(gA, a); }
                else va = tail(gA, cA, anyA);
                store(rowA, cA, va);
            }
        }
        if (ytma) {
            // the box brought the neighbouring strips' pixels into the halo columns: a tile contributes only what it owns
            mbar_wait(&s_bar, bar_phase);
            bar_phase ^= 1;
            for (int k = tid; k < FL; k += FTHREADS) {
                S[k * FPITCH_T + FHALO_T - 2] = 0.f;
                S[k * FPITCH_T + FHALO_T - 1] = 0.f;
                S[k * FPITCH_T + FHALO_T + FW] = 0.f;
                S[k * FPITCH_T + FHALO_T + FW + 1] = 0.f;
                S[k * FPITCH_T + FHALO_T + FW + 2] = 0.f;
            }
        } else {
            for (int k = tid; k < FL; k += FTHREADS) {
                S[k * FPITCH] = 0.f;
                S[k * FPITCH + 1] = 0.f;
                S[k * FPITCH + FW + 2] = 0.f;
                S[k * FPITCH + FW + 3] = 0.f;
                S[k * FPITCH + FW + 4] = 0.f;
            }
        }
        __syncthreads();
        // ---- pull the next slab's operand boxes towards L2 while this slab is sampled: one instruction per stream ----
        if ((P.use_tma & 1) && tid == 32 && slab + 1 < nslab) {
            const int K1 = K0 + FL, nd = (int)blockIdx.y;
            if (xdom) {
                tma_prefetch_3d(&P.mapX[0], K1, U0, nd);
                if (mode != 0) tma_prefetch_3d(&P.mapX[1], K1, U0, nd);
                if (mode == 2) { tma_prefetch_3d(&P.mapX[2], K1, U0, nd); if (writer) tma_prefetch_3d(&P.mapX[3], K1, U0, nd); }
            } else {
                tma_prefetch_3d(&P.mapY[0], U0, K1, nd);
                if (mode != 0) tma_prefetch_3d(&P.mapY[1], U0, K1, nd);
                if (mode == 2) { tma_prefetch_3d(&P.mapY[2], U0, K1, nd); if (writer) tma_prefetch_3d(&P.mapY[3], U0, K1, nd); }
            }
        }
        // ---- sample (instantiated for the two row pitches: the row offsets are immediates) -----------------
        auto sample = [&](auto pitch_c) {
        constexpr int PITCH = decltype(pitch_c)::value;
        for (int ai = slot; ai < na; ai += FTHREADS / FTPA) {
            const int4 win = s_win[ai];
            const float4 ar = s_ang[ai];
            const float fb = __int_as_float(win.w);
            const float im = ar.x, s = ar.y, rs = ar.z;
            const float ulo = -1.0f, uhi = (float)Wt;
            const int jseg = s_jseg[ai];
            for (int j = win.x + t; j <= win.y; j += FTPA) {
                const float u0 = ((float)(j - win.z) - fb) * im;   // u_k = u0 - k s  (interp-axis pixel coordinate)
                // steps with -1 <= u_k <= Wt.  The float bounds may be off by one step at either end, which moves u
                // by < 1 pixel: the two-column zero halo makes such a sample contribute exactly 0, and a dropped
                // boundary sample has weight ~0, so no exact fix-up loop is needed.
                int klo = 0, khi = Lt - 1;
                if (rs != 0.f) {
                    const float ka = (u0 - uhi) * rs, kb = (u0 - ulo) * rs;
                    klo = max(0, (int)ceilf(fminf(ka, kb)));
                    khi = min(Lt - 1, (int)floorf(fmaxf(ka, kb)));
                } else if (u0 < ulo || u0 > uhi) {
                    khi = -1;
                }
                float acc = 0.f;
                f32x2 acc2 = splat2(0.f);      // packed partial sums of the (even, odd) steps of the 8-step blocks
                const f32x2 s2 = splat2(s);
                float kf = (float)klo;
                unsigned rowa = sbase + (unsigned)(klo * PITCH * 4);
                int k = klo;
                // one sample: fi = floor(u) + magic (exact, FADD.RM), f = u - floor(u), both taps through one address
                // register, acc += a + f (b - a).  The 8-step block does two steps per instruction with the packed fp32
                // pipe (FFMA2 / FADD2.RM / FADD2): 13 issue slots per pair instead of 10 per sample.
#define FWD_SAMPLE(UU, ROWOFF)                                                              \
    {                                                                                       \
        const float fi = __fadd_rd((UU), kMagic);            /* floor(u) + magic, exact */ \
        const float f = (UU) - (fi - kMagic);                /* in [0, 1) */               \
        const unsigned addr = (__float_as_uint(fi) << 2) + rowa + (ROWOFF);                 \
        float a, b;                                                                         \
        lds_pair(addr, a, b);                                                               \
        acc += fmaf(f, b - a, a);                                                           \
    }
#define FWD_PAIRS(NSTEP)                                                                                    \
    {                                                                                                       \
        const float ub = fmaf(-kf, s, u0);                                                                  \
        const f32x2 ub2 = pack2(ub, ub - s);                                                                \
        _Pragma("unroll") for (int q = 0; q < (NSTEP); q += 2) {                                            \
            const f32x2 uu = (q == 0) ? ub2 : fma2(s2, splat2(-(float)q), ub2);                             \
            const f32x2 fi = add2_rm(uu, splat2(kMagic));                                                   \
            const f32x2 f = sub2(uu, sub2(fi, splat2(kMagic)));                                             \
            float fia, fib;                                                                                 \
            unpack2(fi, fia, fib);                                                                          \
            float a0, b0, a1, b1;                                                                           \
            lds_pair((__float_as_uint(fia) << 2) + rowa + (unsigned)(q * PITCH * 4), a0, b0);               \
            lds_pair((__float_as_uint(fib) << 2) + rowa + (unsigned)((q + 1) * PITCH * 4), a1, b1);         \
            const f32x2 av = pack2(a0, a1);                                                                 \
            acc2 = add2(acc2, fma2(f, sub2(pack2(b0, b1), av), av));                                        \
        }                                                                                                   \
        kf += (float)(NSTEP);                                                                               \
        rowa += (NSTEP) * PITCH * 4;                                                                        \
        k += (NSTEP);                                                                                       \
    }
                while (k + 7 <= khi) FWD_PAIRS(8)
                // tail: one 4-step and one 2-step packed block before the (at most one) single step
                if (k + 3 <= khi) FWD_PAIRS(4)
                if (k + 1 <= khi) FWD_PAIRS(2)
#undef FWD_PAIRS
                for (; k <= khi; ++k) {
                    const float u = fmaf(-kf, s, u0);
                    FWD_SAMPLE(u, 0u)
                    kf += 1.f;
                    rowa += PITCH * 4;
                }
#undef FWD_SAMPLE
                {
                    float lo, hi;
                    unpack2(acc2, lo, hi);
                    acc += lo + hi;
                }
                const int idx = j - jseg;
                if (idx >= 0 && idx < span) acc_s[ai * span + idx] += acc;
            }
        }
        };
        if (ytma) sample(std::integral_constant<int, FPITCH_T>{});
        else sample(std::integral_constant<int, FPITCH>{});
    }
    __syncthreads();
    const int nRec = nTi * nSeg, rec = sg * nTi + ti;
    for (int i = tid; i < na * span; i += FTHREADS) {
        const int ai = i / span, k = i % span;
        P.recs[((long long)s_aid[ai] * nRec + rec) * span + k] = acc_s[i];
    }
    if (tid < na) P.jstart[(long long)s_aid[tid] * nRec + rec] = s_jseg[tid];
    if (mode == 2 && writer) {   // block-uniform: exact <r', r'> of this node over its nTi*nSeg writer blocks
        float v[1] = {rsum};
        block_sum<1>(v, red);
        grid_reduce_store<1>(v, P.part + (long long)blockIdx.y * nRec, P.counter + blockIdx.y, rec, nRec,
                             P.scal + (long long)node * NSCAL + P.rr_out, red);
    }
}

__global__ void __launch_bounds__(256)
fwd_reduce_kernel(const FwdReduceParams P) {
    const int aid = P.A0 + blockIdx.x;          // angle rows on grid.x (can exceed 65535 in slice-batched runs)
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= P.D) return;
    const int* __restrict__ js = P.jstart + (long long)aid * P.nRec;
    const float* __restrict__ rc = P.recs + (long long)aid * P.nRec * P.span;
    float acc = 0.f;
    for (int r = 0; r < P.nRec; ++r) {
        const int idx = j - __ldg(js + r);
        if ((unsigned)idx < (unsigned)P.span) acc += rc[(long long)r * P.span + idx];
    }
    P.out[(long long)aid * P.D + j] = acc * P.ang[aid].wgt;
}

// =================================================================================================
// K2 back-projection: exact transpose of K1 as an atomics-free gather:
//   (A^T q)[ix,iy] = sum_theta (h/|a|) sum_j max(0, 1 - |tau - j|/omega) q[theta, j]
// One block owns a BTX x BTY pixel tile, stages every angle's detector window (pre-scaled by the step
// weight and the node precision) into shared memory and accumulates in registers (4*BPG pixels / thread).
// Epilogues fuse the rest of the CG operator:  H v = A^T P A v + rhoD .* v + mu K^T K v, the <v, Hv>
// / <r, r> block reductions (warp shuffle + last-block-done grid reduce), and the CG initial residual.
// =================================================================================================
template <int MODE>
__global__ void __launch_bounds__(BTHREADS)
back_tile_kernel(const __grid_constant__ BackParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* qs = reinterpret_cast<float*>(smem_raw);       // [BAC][bpitch]  (every window 128-byte aligned with TMA)
    float4* s_c = reinterpret_cast<float4*>(qs + BAC * P.bpitch + ((4 - ((BAC * P.bpitch) & 3)) & 3));  // [BAC]
    unsigned* s_qa = reinterpret_cast<unsigned*>(s_c + BAC);  // [BAC] biased shared-window address of each window
    int* s_jw = reinterpret_cast<int*>(s_qa + BAC);           // [BAC] first detector bin of each window
    float* s_sc = reinterpret_cast<float*>(s_jw + BAC);       // [BAC] window scale (step weight * precision)
    __shared__ __align__(16) float red[96];
    __shared__ __align__(8) unsigned long long s_bar;         // TMA arrival barrier of the window copies

    const int node = P.node0 + blockIdx.z;
    if (P.ctl && !P.ctl[node].active) return;   // a14 retry pass: this node was accepted already
    const int N = P.N, D = P.D, bspan = P.bpitch;   // row pitch of the staged windows
    // TMA staging: raw sinogram windows land by bulk tensor copies (zero-filled outside [0, D)); the window scale is
    // folded into the hat weights instead of the staged values
    const bool tma = (MODE != BACK_COLNORM2) && P.use_tma;
    unsigned bar_phase = 0;
    if (tma && threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
    }
    const int X0 = blockIdx.y * BTX, Y0 = blockIdx.x * BTY;
    const int tid = threadIdx.x;
    const int lx = tid >> 3, ly = (tid & 7) * 4;          // pixels (X0+lx, Y0+32*g+ly+{0..3}), g < BPG
    const int ix = X0 + lx;
    const int abeg = P.aptr[node], aend = P.aptr[node + 1];
    const double cx = 0.5 * (N - 1), cj = 0.5 * (D - 1);
    const float prec = (MODE == BACK_COLNORM2 || P.prec == nullptr) ? 1.f : P.prec[node];

    // accumulators as packed pairs: acc2[k] = pixels (2k, 2k+1) of the thread's 4*BPG pixels (FFMA2, see common.cuh)
    f32x2 acc2[2 * BPG];
#pragma unroll
    for (int k = 0; k < 2 * BPG; ++k) acc2[k] = splat2(0.f);
    float fx = (float)lx, fy = (float)ly;
    opaque(fx);
    opaque(fy);
    for (int c0 = abeg; c0 < aend; c0 += BAC) {
        const int na = min(BAC, aend - c0);
        __syncthreads();
        // window geometry: one thread per angle (fp64 once per (tile, angle))
        if (tid < na) {
            const AngleRec r = P.ang[c0 + tid];
            const double tau0 = cj + (X0 - cx) * r.ct + (Y0 - cx) * r.st;  // tau of tile pixel (0,0)
            const double e1 = (BTX - 1) * r.ct, e2 = (BTY - 1) * r.st;
            const double om = (double)(1.0f / r.inv_om);
            int jw0 = (int)floor(tau0 + fmin(e1, 0.0) + fmin(e2, 0.0) - om) - 1;
            if (tma) jw0 &= ~3;   // a box must start on a 16-byte boundary of the sinogram row (bpitch has the slack)
            // window-relative tau of pixel (0,0), pre-biased by -1/2 for the rint trick
            s_c[tid] = make_float4((float)(tau0 - (double)jw0 - 0.5), (float)r.ct, (float)r.st, r.inv_om);
            // read back through shared memory so the magic-number bias stays folded into ONE register
            s_qa[tid] = smem_u32(qs) + (unsigned)(tid * bspan * 4) - ((unsigned)kMagicBits << 2);
            s_jw[tid] = jw0;
            s_sc[tid] = (MODE == BACK_COLNORM2) ? r.wgt * r.wgt : r.wgt * prec;
        }
        if (tma) {
            // the threads that computed the windows (all in warp 0) issue one box copy each; lane 0 arms the barrier
            // with the total byte count first
            if (tid < 32) {
                if (tid == 0) mbar_expect_tx(&s_bar, (unsigned)(na * bspan * 4));
                __syncwarp();
                if (tid < na) tma_load_2d(qs + tid * bspan, &P.qmap, s_jw[tid], c0 + tid, &s_bar);
            }
            __syncthreads();                 // window records visible to every thread
            mbar_wait(&s_bar, bar_phase);    // window data has landed
            bar_phase ^= 1;
        } else {
            __syncthreads();
            // stage the detector windows: one warp per angle row, coalesced
            for (int ai = tid >> 5; ai < na; ai += BTHREADS / 32) {
                const int jw0 = s_jw[ai];
                const float sc = s_sc[ai];
                const float* __restrict__ qrow = P.q + (long long)(c0 + ai) * D;
                for (int k = (tid & 31); k < bspan; k += 32) {
                    const int j = jw0 + k;
                    float val = 0.f;
                    if (j >= 0 && j < D) val = (MODE == BACK_COLNORM2) ? sc : sc * qrow[j];
                    qs[ai * bspan + k] = val;
                }
            }
            __syncthreads();
        }
        for (int ai = 0; ai < na; ++ai) {
            const float4 c = s_c[ai];
            const float a = c.w;  // 1/omega
            const float tb = fmaf(fx, c.y, fmaf(fy, c.z, c.x));   // tau - 1/2 of this thread's first pixel
            const float wsc = tma ? s_sc[ai] : 1.f;               // TMA: raw windows, the scale rides on the weights
            if (a >= 1.0f) {
                // omega <= 1 bin: the hat touches bins rint(tau - 1/2) and the next one.  Pixel pairs (iy, iy+1) go
                // through the packed fp32 pipe: per pair 1 FFMA2 (tau) + 3 FADD2 (rint, fraction) + 2 FFMA2 (weights)
                // + 2 FFMA2 (accumulate) next to the per-pixel LEA / 2 LDS / 2 FMNMX.
                const float c1 = fmaf(-0.5f, a, 1.f) * wsc, as = a * wsc;
                const unsigned qa = s_qa[ai];
                const f32x2 tb2 = pack2(tb, tb + c.z), cz2 = splat2(c.z), a2 = splat2(as), na2 = splat2(-as), c12 = splat2(c1);
#pragma unroll
                for (int k = 0; k < 2 * BPG; ++k) {
                    const float o = (float)(2 * (k & 1) + 32 * (k >> 1));   // iy offset of the pair's first pixel
                    const f32x2 v = (k == 0) ? tb2 : fma2(cz2, splat2(o), tb2);
                    const f32x2 fi = add2(v, splat2(kMagic));
                    const f32x2 up = sub2(v, sub2(fi, splat2(kMagic)));     // in [-1/2, 1/2]
                    float fia, fib;
                    unpack2(fi, fia, fib);
                    float q0a, q1a, q0b, q1b;
                    lds_pair((__float_as_uint(fia) << 2) + qa, q0a, q1a);
                    lds_pair((__float_as_uint(fib) << 2) + qa, q0b, q1b);
                    float w0a, w0b, w1a, w1b;
                    unpack2(fma2(na2, up, c12), w0a, w0b);
                    unpack2(fma2(a2, up, c12), w1a, w1b);
                    f32x2 w0 = pack2(fmaxf(0.f, w0a), fmaxf(0.f, w0b)), w1 = pack2(fmaxf(0.f, w1a), fmaxf(0.f, w1b));
                    if (MODE == BACK_COLNORM2) { w0 = mul2(w0, w0); w1 = mul2(w1, w1); }
                    acc2[k] = fma2(w0, pack2(q0a, q0b), fma2(w1, pack2(q1a, q1b), acc2[k]));
                }
            } else {
                const float* __restrict__ qa = qs + ai * bspan;
                const float om = 1.f / a;
#pragma unroll
                for (int k = 0; k < 2 * BPG; ++k) {
                    float ap[2];
                    unpack2(acc2[k], ap[0], ap[1]);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float off = (float)(2 * (k & 1) + e + 32 * (k >> 1));
                        const float tau = fmaf(off, c.z, tb) + 0.5f;
                        const int jlo = (int)ceilf(tau - om), jhi = (int)floorf(tau + om);
                        for (int j = max(jlo, 0); j <= min(jhi, bspan - 1); ++j) {
                            float w = fmaxf(0.f, 1.f - fabsf(tau - (float)j) * a) * wsc;
                            if (MODE == BACK_COLNORM2) w *= w;
                            ap[e] = fmaf(w, qa[j], ap[e]);
                        }
                    }
                    acc2[k] = pack2(ap[0], ap[1]);
                }
            }
        }
    }
    float acc[4 * BPG];
#pragma unroll
    for (int k = 0; k < 2 * BPG; ++k) unpack2(acc2[k], acc[2 * k], acc[2 * k + 1]);

    // ---- epilogue ----------------------------------------------------------------------------------
    const long long nb = (long long)blockIdx.z * P.stride;
    const bool rowok = ix < N;
    const bool vecok = (N & 3) == 0;
    float dsum = 0.f, dsum2 = 0.f, dsum3 = 0.f;   // HP: <p,Hp>, (unused), <Hp,Hp> ; RESID0: <r,r>
    if (MODE == BACK_PLAIN || MODE == BACK_COLNORM2) {
        if (rowok) {
#pragma unroll
            for (int h = 0; h < BPG; ++h) {
                const int iy = Y0 + ly + 32 * h;
                float* o = P.out + nb + (long long)ix * N + iy;
                if (iy + 3 < N && vecok) st4(o, make_float4(acc[4 * h], acc[4 * h + 1], acc[4 * h + 2], acc[4 * h + 3]));
                else for (int px = 0; px < 4; ++px) if (iy + px < N) o[px] = acc[4 * h + px];
            }
        }
        return;
    } else {
        const float* __restrict__ v = P.v + nb;
        const float rhoDs = P.rhoD_vec ? 0.f : P.rhoD_s[node];
        // tiles that touch no image border take the unguarded path
        const bool interior = vecok && X0 >= 1 && X0 + BTX < N && Y0 >= 1 && Y0 + BTY < N;
#pragma unroll
        for (int h = 0; h < BPG; ++h) {
            const int iy = Y0 + ly + 32 * h;
            const long long g = (long long)ix * N + iy;
            float res[4];
            if (interior) {
                const float4 c4 = ld4(v + g), u4 = ld4(v + g - N), d4 = ld4(v + g + N);
                const float lf = v[g - 1], rt = v[g + 4];
                const float vc[6] = {lf, c4.x, c4.y, c4.z, c4.w, rt};
                const float vu[4] = {u4.x, u4.y, u4.z, u4.w}, vd[4] = {d4.x, d4.y, d4.z, d4.w};
                float dd[4] = {rhoDs, rhoDs, rhoDs, rhoDs};
                if (P.rhoD_vec) { const float4 t = ld4(P.rhoD_vec + nb + g); dd[0] = t.x; dd[1] = t.y; dd[2] = t.z; dd[3] = t.w; }
                float rh[4] = {0.f, 0.f, 0.f, 0.f};
                if (MODE == BACK_RESID0) {
                    const float4 r0 = ld4(P.rhs0 + nb + g), t0 = ld4(P.tvterm + nb + g);
                    rh[0] = r0.x + t0.x; rh[1] = r0.y + t0.y; rh[2] = r0.z + t0.z; rh[3] = r0.w + t0.w;
                }
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                    const float c = vc[px + 1];
                    const float lap = ((c - vu[px]) + (c - vd[px])) + ((c - vc[px]) + (c - vc[px + 2]));
                    const float hv = acc[4 * h + px] + fmaf(dd[px], c, P.mu * lap);
                    if (MODE == BACK_HP) {
                        res[px] = hv;
                        dsum = fmaf(c, hv, dsum); dsum3 = fmaf(hv, hv, dsum3);
                    } else { const float rr = rh[px] - hv; res[px] = rr; dsum = fmaf(rr, rr, dsum); }
                }
                st4(P.out + nb + g, make_float4(res[0], res[1], res[2], res[3]));
                if (MODE == BACK_RESID0 && P.p_out) st4(P.p_out + nb + g, make_float4(res[0], res[1], res[2], res[3]));
            } else if (rowok && iy < N) {
                float vc[6], vu[4], vd[4];  // centre row with one halo each side, up row, down row
#pragma unroll
                for (int px = -1; px < 5; ++px) {
                    const int y = iy + px;
                    vc[px + 1] = (y >= 0 && y < N) ? v[g + px] : 0.f;
                }
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                    const bool ok = iy + px < N;
                    vu[px] = (ok && ix >= 1) ? v[g + px - N] : 0.f;
                    vd[px] = (ok && ix + 1 < N) ? v[g + px + N] : 0.f;
                }
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                    const int y = iy + px;
                    if (y >= N) { res[px] = 0.f; continue; }
                    const float c = vc[px + 1];
                    float lu = 0.f, ld = 0.f, ll = 0.f, lr = 0.f;   // same association as the interior path
                    if (ix >= 1) lu = c - vu[px];
                    if (ix + 1 < N) ld = c - vd[px];
                    if (y >= 1) ll = c - vc[px];
                    if (y + 1 < N) lr = c - vc[px + 2];
                    const float lap = (lu + ld) + (ll + lr);
                    const float dd = P.rhoD_vec ? P.rhoD_vec[nb + g + px] : rhoDs;
                    const float hv = acc[4 * h + px] + fmaf(dd, c, P.mu * lap);
                    if (MODE == BACK_HP) {
                        res[px] = hv;
                        dsum = fmaf(c, hv, dsum);
                        dsum3 = fmaf(hv, hv, dsum3);
                    } else {
                        const float rr = (P.rhs0[nb + g + px] + P.tvterm[nb + g + px]) - hv;
                        res[px] = rr;
                        dsum = fmaf(rr, rr, dsum);
                    }
                }
                float* o = P.out + nb + g;
                if (iy + 3 < N && vecok) {
                    st4(o, make_float4(res[0], res[1], res[2], res[3]));
                    if (MODE == BACK_RESID0 && P.p_out) st4(P.p_out + nb + g, make_float4(res[0], res[1], res[2], res[3]));
                } else {
                    for (int px = 0; px < 4; ++px)
                        if (iy + px < N) {
                            o[px] = res[px];
                            if (MODE == BACK_RESID0 && P.p_out) P.p_out[nb + g + px] = res[px];
                        }
                }
            }
        }
        const int nblk = gridDim.x * gridDim.y, blk = blockIdx.y * gridDim.x + blockIdx.x;
        if (MODE == BACK_HP) {   // S_PHP, S_RHP, S_HPHP are consecutive slots
            float vsum[3] = {dsum, dsum2, dsum3};
            block_sum<3>(vsum, red);
            grid_reduce_store<3>(vsum, P.part + (long long)blockIdx.z * nblk * 3, P.counter + blockIdx.z, blk, nblk,
                                 P.scal + (long long)node * NSCAL + P.dot_slot, red);
        } else {
            float vsum[1] = {dsum};
            block_sum<1>(vsum, red);
            grid_reduce_store<1>(vsum, P.part + (long long)blockIdx.z * nblk, P.counter + blockIdx.z, blk, nblk,
                                 P.scal + (long long)node * NSCAL + P.dot_slot, red);
        }
    }
}

// ---- host launchers ---------------------------------------------------------------------------------
static size_t fwd_smem_bytes(int span) {
    size_t f = (size_t)FL * FPITCH_T + (size_t)FAC * span;
    f += (f & 1);
    return f * sizeof(float) + FAC * sizeof(double) + 2 * FAC * sizeof(int) + FAC * sizeof(int4) + FAC * sizeof(float4);
}

// ADMM_B200_TMA = bit mask of the TMA uses that are on (default 7): 1 back-projector windows, 2 forward L2 prefetch of
// the next slab, 4 forward y-dominant tile landing.  The staging-loop paths stay selectable for A/B measurements.
static int tma_mask() {
    static const int m = [] { const char* e = getenv("ADMM_B200_TMA"); return e ? atoi(e) : 7; }();
    return m;
}

// tensor maps of one operand stream [nodes][N][N] (node stride `stride` floats)
static bool fwd_maps(const float* base, int N, long long stride, int nodes, CUtensorMap* mapY, CUtensorMap* mapX) {
    const unsigned long long dims[3] = {(unsigned long long)N, (unsigned long long)N, (unsigned long long)nodes};
    const unsigned long long strides[2] = {(unsigned long long)N * sizeof(float), (unsigned long long)stride * sizeof(float)};
    const unsigned boxY[3] = {(unsigned)FPITCH_T, (unsigned)FL, 1u}, boxX[3] = {(unsigned)FL, (unsigned)FW, 1u};
    return tma_encode_f32(mapY, base, 3, dims, strides, boxY) && tma_encode_f32(mapX, base, 3, dims, strides, boxX);
}

cudaError_t launch_forward(const FwdParams& P0, int nodes, int max_chunks, const FwdReduceParams& R,
                           cudaStream_t st) {
    FwdParams P = P0;
    P.use_tma = 0;
    if (tma_mask() & 6) {
        bool ok = fwd_maps(P.img, P.N, P.img_stride, nodes, &P.mapY[0], &P.mapX[0]);
        if (ok && P.mode != 0) ok = fwd_maps(P.r, P.N, P.img_stride, nodes, &P.mapY[1], &P.mapX[1]);
        if (ok && P.mode == 2)
            ok = fwd_maps(P.hp, P.N, P.img_stride, nodes, &P.mapY[2], &P.mapX[2]) &&
                 fwd_maps(P.x_io, P.N, P.img_stride, nodes, &P.mapY[3], &P.mapX[3]);
        if (ok) P.use_tma = ((tma_mask() & 2) ? 1 : 0) | ((tma_mask() & 4) ? 2 : 0);
    }
    // the dynamic shared-memory limit is a per-device function attribute: remember what was set on EACH device
    static std::mutex mu;
    static int configured_span[64];
    static bool init = false;
    const size_t smem = fwd_smem_bytes(P.span);
    {
        std::lock_guard<std::mutex> lk(mu);
        if (!init) { for (int& v : configured_span) v = -1; init = true; }
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const int slot = (dev >= 0 && dev < 64) ? dev : 63;
        if (dev >= 63 || configured_span[slot] < P.span) {
            e = cudaFuncSetAttribute(fwd_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured_span[slot] = P.span;
        }
    }
    dim3 grid(P.nTi * P.nSeg * max_chunks, nodes, 2);
    { ProfScope ps(P.mode ? KC_FWD_FUSED : KC_FWD, st); fwd_strip_kernel<<<grid, FTHREADS, smem, st>>>(P); }
    dim3 rgrid(R.A1 - R.A0, (R.D + 255) / 256);
    if (rgrid.x > 0) { ProfScope ps(KC_FWD_REDUCE, st); fwd_reduce_kernel<<<rgrid, 256, 0, st>>>(R); }
    return cudaGetLastError();
}

cudaError_t launch_back(int mode, const BackParams& P0, int nodes, cudaStream_t st) {
    BackParams P = P0;
    P.bpitch = P.bspan;
    P.use_tma = 0;
    if (mode != BACK_COLNORM2 && (tma_mask() & 1) && P.q) {
        // sinogram rows as a 2-D tensor [rows >= every angle row of this launch][D]; a window is the box (bpitch, 1)
        // at (jw0, row).  The row count only bounds the coordinates: one past the last angle row the kernel can touch
        const unsigned bp = (unsigned)((P.bspan + 3 + 31) & ~31);   // + 3: the window origin is rounded down to a multiple of 4 bins
        const unsigned long long dims[2] = {(unsigned long long)P.D, (unsigned long long)P.A_rows};
        const unsigned long long strides[1] = {(unsigned long long)P.D * sizeof(float)};
        const unsigned box[2] = {bp, 1u};
        if (bp <= 256 && P.A_rows > 0 && tma_encode_f32(&P.qmap, P.q, 2, dims, strides, box)) {
            P.bpitch = (int)bp;
            P.use_tma = 1;
        }
    }
    size_t f = (size_t)BAC * P.bpitch;
    f += (4 - (f & 3)) & 3;
    const size_t smem = f * sizeof(float) + BAC * sizeof(float4) + BAC * (sizeof(unsigned) + sizeof(int) + sizeof(float));
    dim3 grid((P.N + BTY - 1) / BTY, (P.N + BTX - 1) / BTX, nodes);
    switch (mode) {
        case BACK_PLAIN: { ProfScope ps(KC_BACK_PLAIN, st); back_tile_kernel<BACK_PLAIN><<<grid, BTHREADS, smem, st>>>(P); } break;
        case BACK_HP: { ProfScope ps(KC_BACK_HP, st); back_tile_kernel<BACK_HP><<<grid, BTHREADS, smem, st>>>(P); } break;
        case BACK_RESID0: { ProfScope ps(KC_BACK_RESID0, st); back_tile_kernel<BACK_RESID0><<<grid, BTHREADS, smem, st>>>(P); } break;
        case BACK_COLNORM2: { ProfScope ps(KC_COLNORM, st); back_tile_kernel<BACK_COLNORM2><<<grid, BTHREADS, smem, st>>>(P); } break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace admm
