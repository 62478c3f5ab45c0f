// projector.cu -- K1 (Joseph forward projector, strip-staged, deterministic), K2 (matched gather
// back-projector with fused H-apply / CG-residual epilogues), K2b (column norms^2).  sm_100a.
//
// Replaces odl.tomo.RayTransform.__call__ / `Ai @ x` (block_2_load_odl_data.py:149,
// block_6_admm_loop_ver2.py:145,193), `Ai.T @ r` (block_6_admm_loop_ver2.py:145) and
// np.sum(A_i*A_i, axis=0) (block_3_graph_and_precisions.py:22).  Discretisation: SURVEY.md App. C.
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
// =================================================================================================
__global__ void __launch_bounds__(FTHREADS)
fwd_strip_kernel(const FwdParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* S = reinterpret_cast<float*>(smem_raw);                 // [FL][FPITCH]
    float* acc_s = S + FL * FPITCH;                                // [FAC][span]
    double* s_base = reinterpret_cast<double*>(acc_s + FAC * P.span + ((FAC * P.span) & 1));  // [FAC]
    int* s_jseg = reinterpret_cast<int*>(s_base + FAC);            // [FAC]
    int* s_aid = s_jseg + FAC;                                     // [FAC]

    const int node = P.node0 + blockIdx.y;
    const int orient = blockIdx.z;  // 0: x-dominant list, 1: y-dominant list
    const int nTi = P.nTi, nSeg = P.nSeg, N = P.N, D = P.D, span = P.span;
    const int bx = blockIdx.x;
    const int ti = bx % nTi, sg = (bx / nTi) % nSeg, ch = bx / (nTi * nSeg);
    const int obeg = P.optr[orient * (P.V + 1) + node], oend = P.optr[orient * (P.V + 1) + node + 1];
    const int a0 = obeg + ch * FAC;
    if (a0 >= oend) return;
    const int na = min(FAC, oend - a0);
    const bool xdom = (orient == 0);

    const float* __restrict__ img = P.img + (long long)blockIdx.y * P.img_stride;
    const int U0 = ti * FW, Wt = min(FW, N - U0);
    const int K0seg = sg * FSEG, Kseg = min(FSEG, N - K0seg);
    const int tid = threadIdx.x;
    const double cx = 0.5 * (N - 1), cj = 0.5 * (D - 1);

    for (int i = tid; i < FAC * span; i += FTHREADS) acc_s[i] = 0.f;
    if (tid < na) {
        const int aid = P.oidx[a0 + tid];
        const AngleRec r = P.ang[aid];
        const double M = xdom ? r.ct : r.st, m = xdom ? r.st : r.ct;
        const double base = cj + (U0 - cx) * M + (K0seg - cx) * m;  // tau of pixel (u=0,k=0) of the segment
        const double e1 = (Wt - 1) * M, e2 = (Kseg - 1) * m;
        const double lo = base + fmin(e1, 0.0) + fmin(e2, 0.0);
        s_base[tid] = base;
        s_jseg[tid] = (int)ceil(lo - fabs(M)) - 1;   // one bin of slack below the exact bound
        s_aid[tid] = aid;
    }

    // fused CG direction update p_new = r + beta p_old
    float beta = 0.f;
    const float* __restrict__ rimg = nullptr;
    float* pout = nullptr;
    if (P.r != nullptr) {
        const double den = P.scal[(long long)node * NSCAL + P.beta_den];
        const double num = P.scal[(long long)node * NSCAL + P.beta_num];
        beta = (den > 0.0) ? (float)(num / den) : 0.f;
        rimg = P.r + (long long)blockIdx.y * P.img_stride;
        // exactly one (orientation, chunk) pass per pixel writes p_new: the first non-empty orientation, chunk 0
        const int x_has = P.optr[node + 1] - P.optr[node];
        const bool writer = (ch == 0) && (xdom ? true : (x_has == 0));
        pout = writer ? (P.p_out + (long long)blockIdx.y * P.img_stride) : nullptr;
    }

    const int slot = tid / FTPA, t = tid % FTPA;
    const int nslab = (Kseg + FL - 1) / FL;
    for (int slab = 0; slab < nslab; ++slab) {
        const int K0 = K0seg + slab * FL, Lt = min(FL, K0seg + Kseg - K0);
        __syncthreads();  // previous slab fully consumed (also orders the acc/zero + setup writes)
        // ---- stage tile -------------------------------------------------------------------------
        if (xdom) {
            // pixel (ix = U0+u, iy = K0+k): contiguous along k.  thread -> (u, 4 consecutive k)
            for (int idx = tid; idx < FW * (FL / 4); idx += FTHREADS) {
                const int u = idx / (FL / 4), k4 = (idx % (FL / 4)) * 4;
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                if (u < Wt) {
                    const long long g = (long long)(U0 + u) * N + K0 + k4;
                    if (k4 + 3 < Lt && ((N & 3) == 0)) {
                        val = ld4(img + g);
                        if (rimg) {
                            const float4 rv = ld4(rimg + g);
                            val.x = fmaf(beta, val.x, rv.x); val.y = fmaf(beta, val.y, rv.y);
                            val.z = fmaf(beta, val.z, rv.z); val.w = fmaf(beta, val.w, rv.w);
                            if (pout) st4(pout + g, val);
                        }
                    } else {
                        float tmp[4] = {0.f, 0.f, 0.f, 0.f};
                        for (int i = 0; i < 4; ++i)
                            if (k4 + i < Lt) {
                                float x = img[g + i];
                                if (rimg) { x = fmaf(beta, x, rimg[g + i]); if (pout) pout[g + i] = x; }
                                tmp[i] = x;
                            }
                        val = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
                    }
                }
                float* dst = S + k4 * FPITCH + 1 + u;
                dst[0] = val.x; dst[FPITCH] = val.y; dst[2 * FPITCH] = val.z; dst[3 * FPITCH] = val.w;
            }
        } else {
            // pixel (ix = K0+k, iy = U0+u): contiguous along u
            for (int idx = tid; idx < FL * FW; idx += FTHREADS) {
                const int k = idx / FW, u = idx % FW;
                float x = 0.f;
                if (k < Lt && u < Wt) {
                    const long long g = (long long)(K0 + k) * N + U0 + u;
                    x = img[g];
                    if (rimg) { x = fmaf(beta, x, rimg[g]); if (pout) pout[g] = x; }
                }
                S[k * FPITCH + 1 + u] = x;
            }
        }
        for (int k = tid; k < FL; k += FTHREADS) {
            S[k * FPITCH] = 0.f;
            S[k * FPITCH + FW + 1] = 0.f;
            S[k * FPITCH + FW + 2] = 0.f;
        }
        __syncthreads();
        // ---- sample ---------------------------------------------------------------------------------
        for (int ai = slot; ai < na; ai += FTHREADS / FTPA) {
            const AngleRec r = P.ang[s_aid[ai]];
            const double M = xdom ? r.ct : r.st, m = xdom ? r.st : r.ct;
            const double base = s_base[ai] + (double)(slab * FL) * m;  // tau of slab pixel (0,0)
            const double e1 = (Wt - 1) * M, e2 = (Lt - 1) * m, om = fabs(M);
            const double tlo = base + fmin(e1, 0.0) + fmin(e2, 0.0) - om;
            const double thi = base + fmax(e1, 0.0) + fmax(e2, 0.0) + om;
            const int jmin = max(0, (int)ceil(tlo)), jmax = min(D - 1, (int)floor(thi));
            const double fbase = floor(base);
            const int jb = (int)fbase;
            const float fb = (float)(base - fbase);
            const float s = r.slope, im = r.inv_major;
            const float vlo = -1.5f, vhi = (float)Wt - 0.5f;
            const int jseg = s_jseg[ai];
            for (int j = jmin + t; j <= jmax; j += FTPA) {
                const float v0 = ((float)(j - jb) - fb) * im - 0.5f;  // v_k = v0 - k s,  u = v + 0.5
                int klo = 0, khi = Lt - 1;
                if (fabsf(s) > 1e-6f) {
                    const float rs = 1.0f / s;
                    const float ka = (v0 - vhi) * rs, kb = (v0 - vlo) * rs;
                    const float kmn = fminf(ka, kb), kmx = fmaxf(ka, kb);
                    klo = max(0, (int)fmaxf(ceilf(kmn) - 1.f, -1.f));
                    khi = min(Lt - 1, (int)fminf(floorf(kmx) + 1.f, (float)FL));
                }
                while (klo <= khi) {
                    const float v = fmaf(-(float)klo, s, v0);
                    if (v > vlo && v <= vhi) break;
                    ++klo;
                }
                while (klo <= khi) {
                    const float v = fmaf(-(float)khi, s, v0);
                    if (v > vlo && v <= vhi) break;
                    --khi;
                }
                float acc = 0.f;
                float kf = (float)klo;
                const float* row = S + klo * FPITCH + 1;
                for (int k = klo; k <= khi; ++k) {
                    const float v = fmaf(-kf, s, v0);
                    const float fi = v + kMagic;
                    const int ii = __float_as_int(fi) - kMagicBits;
                    const float f = (v - (fi - kMagic)) + 0.5f;
                    const float a = row[ii], b = row[ii + 1];
                    acc += fmaf(f, b - a, a);
                    kf += 1.f;
                    row += FPITCH;
                }
                const int idx = j - jseg;
                if (idx >= 0 && idx < span) acc_s[ai * span + idx] += acc;
            }
        }
    }
    __syncthreads();
    const int nRec = nTi * nSeg, rec = sg * nTi + ti;
    for (int i = tid; i < na * span; i += FTHREADS) {
        const int ai = i / span, k = i % span;
        P.recs[((long long)s_aid[ai] * nRec + rec) * span + k] = acc_s[i];
    }
    if (tid < na) P.jstart[(long long)s_aid[tid] * nRec + rec] = s_jseg[tid];
}

__global__ void __launch_bounds__(256)
fwd_reduce_kernel(const FwdReduceParams P) {
    const int aid = P.A0 + blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
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
// weight and the node precision) into shared memory and accumulates in registers (4 pixels / thread).
// Epilogues fuse the rest of the CG operator:  H v = A^T P A v + rhoD .* v + mu K^T K v, the <v, Hv>
// / <r, r> block reductions (warp shuffle + last-block-done grid reduce), and the CG initial residual.
// =================================================================================================
template <int MODE>
__global__ void __launch_bounds__(BTHREADS)
back_tile_kernel(const BackParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* qs = reinterpret_cast<float*>(smem_raw);       // [BAC][bspan]
    float4* s_c = reinterpret_cast<float4*>(qs + BAC * P.bspan + ((4 - ((BAC * P.bspan) & 3)) & 3));  // [BAC]
    __shared__ float red[64];

    const int node = P.node0 + blockIdx.z;
    const int N = P.N, D = P.D, bspan = P.bspan;
    const int X0 = blockIdx.y * BTX, Y0 = blockIdx.x * BTY;
    const int tid = threadIdx.x;
    const int lx = tid / (BTY / 4), ly = (tid % (BTY / 4)) * 4;
    const int ix = X0 + lx, iy = Y0 + ly;
    const int abeg = P.aptr[node], aend = P.aptr[node + 1];
    const double cx = 0.5 * (N - 1), cj = 0.5 * (D - 1);
    const float prec = (MODE == BACK_COLNORM2 || P.prec == nullptr) ? 1.f : P.prec[node];

    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = abeg; c0 < aend; c0 += BAC) {
        const int na = min(BAC, aend - c0);
        __syncthreads();
        // stage detector windows: one warp per angle row
        for (int ai = tid / 32; ai < na; ai += BTHREADS / 32) {
            const AngleRec r = P.ang[c0 + ai];
            const double tau0 = cj + (X0 - cx) * r.ct + (Y0 - cx) * r.st;  // tau of tile pixel (0,0)
            const double e1 = (BTX - 1) * r.ct, e2 = (BTY - 1) * r.st;
            const double om = (double)(1.0f / r.inv_om);
            const int jw0 = (int)floor(tau0 + fmin(e1, 0.0) + fmin(e2, 0.0) - om) - 1;
            const float sc = (MODE == BACK_COLNORM2) ? r.wgt * r.wgt : r.wgt * prec;
            for (int k = (tid & 31); k < bspan; k += 32) {
                const int j = jw0 + k;
                float val = 0.f;
                if (j >= 0 && j < D) val = (MODE == BACK_COLNORM2) ? sc : sc * P.q[(long long)(c0 + ai) * D + j];
                qs[ai * bspan + k] = val;
            }
            if ((tid & 31) == 0)
                s_c[ai] = make_float4((float)(tau0 - (double)jw0), (float)r.ct, (float)r.st, r.inv_om);
        }
        __syncthreads();
        for (int ai = 0; ai < na; ++ai) {
            const float4 c = s_c[ai];
            const float* __restrict__ qa = qs + ai * bspan;
            const float a = c.w;  // 1/omega
            const float tb = fmaf((float)lx, c.y, fmaf((float)ly, c.z, c.x));
            if (a >= 1.0f) {
                const float c1 = 1.f - 0.5f * a;
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                    const float v = fmaf((float)px, c.z, tb) - 0.5f;
                    const float fi = v + kMagic;
                    const int j0 = __float_as_int(fi) - kMagicBits;
                    const float up = v - (fi - kMagic);
                    float w0 = fmaxf(0.f, fmaf(-a, up, c1)), w1 = fmaxf(0.f, fmaf(a, up, c1));
                    if (MODE == BACK_COLNORM2) { w0 *= w0; w1 *= w1; }
                    acc[px] = fmaf(w0, qa[j0], fmaf(w1, qa[j0 + 1], acc[px]));
                }
            } else {
                const float om = 1.f / a;
#pragma unroll
                for (int px = 0; px < 4; ++px) {
                    const float tau = fmaf((float)px, c.z, tb);
                    const int jlo = (int)ceilf(tau - om), jhi = (int)floorf(tau + om);
                    for (int j = max(jlo, 0); j <= min(jhi, bspan - 1); ++j) {
                        float w = fmaxf(0.f, 1.f - fabsf(tau - (float)j) * a);
                        if (MODE == BACK_COLNORM2) w *= w;
                        acc[px] = fmaf(w, qa[j], acc[px]);
                    }
                }
            }
        }
    }

    // ---- epilogue ----------------------------------------------------------------------------------
    const long long nb = (long long)blockIdx.z * P.stride;
    const bool rowok = ix < N;
    float dsum = 0.f;
    if (MODE == BACK_PLAIN || MODE == BACK_COLNORM2) {
        if (rowok) {
            float* o = P.out + nb + (long long)ix * N + iy;
            if (iy + 3 < N && (N & 3) == 0) st4(o, make_float4(acc[0], acc[1], acc[2], acc[3]));
            else for (int px = 0; px < 4; ++px) if (iy + px < N) o[px] = acc[px];
        }
        return;
    } else {
        if (rowok) {
            const float* __restrict__ v = P.v + nb;
            const long long g = (long long)ix * N + iy;
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
            const float rhoDs = P.rhoD_vec ? 0.f : P.rhoD_s[node];
            float res[4];
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                const int y = iy + px;
                if (y >= N) { res[px] = 0.f; continue; }
                const float c = vc[px + 1];
                float lap = 0.f;
                if (ix >= 1) lap += c - vu[px];
                if (ix + 1 < N) lap += c - vd[px];
                if (y >= 1) lap += c - vc[px];
                if (y + 1 < N) lap += c - vc[px + 2];
                const float dd = P.rhoD_vec ? P.rhoD_vec[nb + g + px] : rhoDs;
                const float hv = acc[px] + fmaf(dd, c, P.mu * lap);
                if (MODE == BACK_HP) {
                    res[px] = hv;
                    dsum = fmaf(c, hv, dsum);
                } else {
                    const float rr = (P.rhs0[nb + g + px] + P.tvterm[nb + g + px]) - hv;
                    res[px] = rr;
                    dsum = fmaf(rr, rr, dsum);
                }
            }
            float* o = P.out + nb + g;
            if (iy + 3 < N && (N & 3) == 0) {
                st4(o, make_float4(res[0], res[1], res[2], res[3]));
                if (MODE == BACK_RESID0) st4(P.p_out + nb + g, make_float4(res[0], res[1], res[2], res[3]));
            } else {
                for (int px = 0; px < 4; ++px)
                    if (iy + px < N) {
                        o[px] = res[px];
                        if (MODE == BACK_RESID0) P.p_out[nb + g + px] = res[px];
                    }
            }
        }
        float vsum[1] = {dsum};
        block_sum<1>(vsum, red);
        const int nblk = gridDim.x * gridDim.y, blk = blockIdx.y * gridDim.x + blockIdx.x;
        grid_reduce_store<1>(vsum, P.part + (long long)blockIdx.z * nblk, P.counter + blockIdx.z, blk, nblk,
                             P.scal + (long long)node * NSCAL + P.dot_slot, red);
    }
}

// ---- host launchers ---------------------------------------------------------------------------------
static size_t fwd_smem_bytes(int span) {
    size_t f = (size_t)FL * FPITCH + (size_t)FAC * span;
    f += (f & 1);
    return f * sizeof(float) + FAC * sizeof(double) + 2 * FAC * sizeof(int);
}

cudaError_t launch_forward(const FwdParams& P, int nodes, int max_chunks, const FwdReduceParams& R,
                           cudaStream_t st) {
    static int configured_span = -1;
    const size_t smem = fwd_smem_bytes(P.span);
    if (configured_span < P.span) {
        cudaError_t e = cudaFuncSetAttribute(fwd_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)fwd_smem_bytes(P.span));
        if (e != cudaSuccess) return e;
        configured_span = P.span;
    }
    dim3 grid(P.nTi * P.nSeg * max_chunks, nodes, 2);
    { ProfScope ps(P.r ? KC_FWD_FUSED : KC_FWD, st); fwd_strip_kernel<<<grid, FTHREADS, smem, st>>>(P); }
    dim3 rgrid((R.D + 255) / 256, R.A1 - R.A0);
    if (rgrid.y > 0) { ProfScope ps(KC_FWD_REDUCE, st); fwd_reduce_kernel<<<rgrid, 256, 0, st>>>(R); }
    return cudaGetLastError();
}

cudaError_t launch_back(int mode, const BackParams& P, int nodes, cudaStream_t st) {
    size_t f = (size_t)BAC * P.bspan;
    f += (4 - (f & 3)) & 3;
    const size_t smem = f * sizeof(float) + BAC * sizeof(float4);
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
