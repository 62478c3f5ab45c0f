// solver_kernels.cu -- K3 fused TV stencil, K4 CG vector updates with in-kernel reductions, K5 edge
// (consensus z / dual y / residual) kernel, K6 right-hand-side assembly, and the small per-iteration
// bookkeeping kernels.  sm_100a.  All are single-pass streaming kernels (HBM-bound).
#include "solver_kernels.cuh"

namespace admm {

// =================================================================================================
// K3: one pass does  g = Kx + w ; d = shrink2(g, lam/mu) ; w' = g - d ; tvterm' = mu K^T (d - w')
// plus the diagnostics that ride on the same data: canonical TV(x), the a14 stationarity norm
//   g_x = tvterm - r_cg - mu K^T K x + lam * div_ref(normalised grad x)   (block_6_admm_loop_ver2.py:137-149
// through the identity A^T P(Ax-b) + rho(Dx - sum q v) = tvterm - r_cg - mu K^T K x), and |x - x_true|^2
// (:199-206).  Gradient = block_4_tv_helpers.py:17-23; div_ref = :25-35 as shipped (border sign quirk).
// =================================================================================================
constexpr int TT = 32;  // tile edge

__global__ void __launch_bounds__(256)
tv_fused_kernel(const TvParams P) {
    __shared__ float xs[TT + 2][TT + 3];      // x rows R0-1..R0+32, cols C0-1..C0+32
    __shared__ float dwx[TT + 1][TT + 2], dwy[TT + 1][TT + 2];  // (d - w') at rows R0-1..R0+31, cols C0-1..C0+31
    __shared__ float pxs[TT + 1][TT + 2], pys[TT + 1][TT + 2];  // normalised gradient (diagnostic)
    __shared__ float red[96];
    const int N = P.N, tid = threadIdx.x;
    const int node = P.node0 + blockIdx.z;
    const long long nb = (long long)blockIdx.z * P.stride;
    const int R0 = blockIdx.y * TT, C0 = blockIdx.x * TT;
    const float* __restrict__ x = P.x + nb;
    const float* __restrict__ win = P.w_in + 2 * nb;
    float* __restrict__ wout = P.w_out + 2 * nb;
    const long long n = (long long)N * N;
    const float kappa = P.lam / P.mu;

    for (int i = tid; i < (TT + 2) * (TT + 2); i += 256) {
        const int lr = i / (TT + 2), lc = i % (TT + 2);
        const int r = R0 - 1 + lr, c = C0 - 1 + lc;
        xs[lr][lc] = (r >= 0 && r < N && c >= 0 && c < N) ? x[(long long)r * N + c] : 0.f;
    }
    __syncthreads();
    float tv = 0.f;
    for (int i = tid; i < (TT + 1) * (TT + 1); i += 256) {
        const int lr = i / (TT + 1), lc = i % (TT + 1);   // extended position (R0-1+lr, C0-1+lc)
        const int r = R0 - 1 + lr, c = C0 - 1 + lc;
        float ox = 0.f, oy = 0.f, nx = 0.f, ny = 0.f;
        if (r >= 0 && r < N && c >= 0 && c < N) {
            const float xc = xs[lr][lc];
            const float gx0 = (r < N - 1) ? xs[lr + 1][lc] - xc : 0.f;
            const float gy0 = (c < N - 1) ? xs[lr][lc + 1] - xc : 0.f;
            const long long g = (long long)r * N + c;
            const float g1 = gx0 + win[g], g2 = gy0 + win[n + g];
            const float nrm = sqrtf(g1 * g1 + g2 * g2);
            const float sc = (nrm > kappa) ? (1.f - kappa / nrm) : 0.f;
            const float d1 = sc * g1, d2 = sc * g2;
            const float w1 = g1 - d1, w2 = g2 - d2;
            ox = d1 - w1;
            oy = d2 - w2;
            const float n0 = sqrtf(gx0 * gx0 + gy0 * gy0);
            if (n0 > 1e-12f) { nx = gx0 / n0; ny = gy0 / n0; }
            if (lr >= 1 && lc >= 1) {  // owned pixel
                wout[g] = w1;
                wout[n + g] = w2;
                tv += n0;
            }
        }
        dwx[lr][lc] = ox; dwy[lr][lc] = oy; pxs[lr][lc] = nx; pys[lr][lc] = ny;
    }
    __syncthreads();
    float gn2 = 0.f, img = 0.f;
    const float* __restrict__ rcg = P.r ? P.r + nb : nullptr;
    const float* __restrict__ xt = P.xtrue;
    float* __restrict__ tvt = P.tvterm + nb;
    for (int i = tid; i < TT * TT; i += 256) {
        const int lr = i / TT + 1, lc = i % TT + 1;
        const int r = R0 + lr - 1, c = C0 + lc - 1;
        if (r >= N || c >= N) continue;
        const long long g = (long long)r * N + c;
        float kt = 0.f;
        if (r >= 1) kt += dwx[lr - 1][lc];
        if (r < N - 1) kt -= dwx[lr][lc];
        if (c >= 1) kt += dwy[lr][lc - 1];
        if (c < N - 1) kt -= dwy[lr][lc];
        const float tv_old = tvt[g];
        tvt[g] = P.mu * kt;
        const float xc = xs[lr][lc];
        if (rcg) {
            float lap = 0.f;
            if (r >= 1) lap += xc - xs[lr - 1][lc];
            if (r < N - 1) lap += xc - xs[lr + 1][lc];
            if (c >= 1) lap += xc - xs[lr][lc - 1];
            if (c < N - 1) lap += xc - xs[lr][lc + 1];
            // div_ref (block_4_tv_helpers.py:25-35): -div, sign-flipped border
            float dv = 0.f;
            if (N >= 2) {
                if (r == 0) dv += pxs[lr][lc];
                else if (r == N - 1) dv -= pxs[lr - 1][lc];
                else dv += pxs[lr - 1][lc] - pxs[lr][lc];
                if (c == 0) dv += pys[lr][lc];
                else if (c == N - 1) dv -= pys[lr][lc - 1];
                else dv += pys[lr][lc - 1] - pys[lr][lc];
            }
            const float gv = tv_old - rcg[g] - P.mu * lap + P.lam * dv;
            gn2 = fmaf(gv, gv, gn2);
        }
        if (xt) { const float e = xc - xt[g]; img = fmaf(e, e, img); }
    }
    float v[3] = {tv, gn2, img};
    block_sum<3>(v, red);
    const int nblk = gridDim.x * gridDim.y, blk = blockIdx.y * gridDim.x + blockIdx.x;
    grid_reduce_store<3>(v, P.part + (long long)blockIdx.z * nblk * 3, P.counter + blockIdx.z, blk, nblk,
                         P.scal + (long long)node * NSCAL + S_TV, red);
}

// =================================================================================================
// K4a: x += alpha p ; r -= alpha Hp ; <r,r> -> scal[rr_out]   (alpha = scal[rr_in] / scal[S_PHP])
// =================================================================================================
__global__ void __launch_bounds__(256)
cg_update_kernel(const CgParams P) {
    __shared__ float red[64];
    const int node = P.node0 + blockIdx.y;
    const double* sc = P.scal + (long long)node * NSCAL;
    const double php = sc[S_PHP], rr = sc[P.rr_in];
    const float alpha = (php > 0.0) ? (float)(rr / php) : 0.f;
    const long long nb = (long long)blockIdx.y * P.stride;
    float* __restrict__ x = P.x + nb;
    float* __restrict__ r = P.r + nb;
    const float* __restrict__ p = P.p + nb;
    const float* __restrict__ hp = P.hp + nb;
    float s = 0.f;
    const long long n4 = P.n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 xv = ld4(x + 4 * i), rv = ld4(r + 4 * i);
        const float4 pv = ld4(p + 4 * i), hv = ld4(hp + 4 * i);
        xv.x = fmaf(alpha, pv.x, xv.x); xv.y = fmaf(alpha, pv.y, xv.y);
        xv.z = fmaf(alpha, pv.z, xv.z); xv.w = fmaf(alpha, pv.w, xv.w);
        rv.x = fmaf(-alpha, hv.x, rv.x); rv.y = fmaf(-alpha, hv.y, rv.y);
        rv.z = fmaf(-alpha, hv.z, rv.z); rv.w = fmaf(-alpha, hv.w, rv.w);
        st4(x + 4 * i, xv); st4(r + 4 * i, rv);
        s = fmaf(rv.x, rv.x, s); s = fmaf(rv.y, rv.y, s); s = fmaf(rv.z, rv.z, s); s = fmaf(rv.w, rv.w, s);
    }
    if (blockIdx.x == 0)
        for (long long i = 4 * n4 + threadIdx.x; i < P.n; i += blockDim.x) {
            const float xv = fmaf(alpha, p[i], x[i]), rv = fmaf(-alpha, hp[i], r[i]);
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
    const double* sc = P.scal + (long long)node * NSCAL;
    const double den = sc[P.rr_in], num = sc[P.rr_out];
    const float beta = (den > 0.0) ? (float)(num / den) : 0.f;
    const long long nb = (long long)blockIdx.y * P.stride;
    const float* __restrict__ r = P.r + nb;
    const float* __restrict__ p = P.p + nb;
    float* __restrict__ po = P.p_out + nb;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += (long long)gridDim.x * blockDim.x)
        po[i] = fmaf(beta, p[i], r[i]);
}

// Ax += alpha(node) * q on sinogram rows (keeps A x current without an extra projection; block_6_ver2:190-194)
__global__ void __launch_bounds__(256)
sino_axpy_kernel(const SinoParams P) {
    const int a = P.A0 + blockIdx.y;
    const int node = P.anode[a];
    const double* sc = P.scal + (long long)node * NSCAL;
    float alpha = 1.f;
    if (P.mode == 1) {
        const double php = sc[S_PHP], rr = sc[P.rr_in];
        alpha = (php > 0.0) ? (float)(rr / php) : 0.f;
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.D; j += gridDim.x * blockDim.x) {
        const long long g = (long long)a * P.D + j;
        P.ax[g] = (P.mode == 0) ? P.q[g] : fmaf(alpha, P.q[g], P.ax[g]);
    }
}

// |A x - b|^2 per node -> scal[S_MSE]   (one block per node: deterministic)
__global__ void __launch_bounds__(256)
sino_resid_kernel(const SinoParams P) {
    __shared__ float red[32];
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
    const int node = P.node0 + blockIdx.y;
    const long long nb = (long long)blockIdx.y * P.stride;
    const int kb = P.nbr_ptr[node], ke = P.nbr_ptr[node + 1];
    const float* __restrict__ atb = P.atb + nb;
    float* __restrict__ out = P.rhs0 + nb;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += (long long)gridDim.x * blockDim.x) {
        float cons = 0.f;
        for (int k = kb; k < ke; ++k) {
            const float* z = reinterpret_cast<const float*>(P.nbr_z[k]);
            const float* y = reinterpret_cast<const float*>(P.nbr_y[k]);
            const float* q = reinterpret_cast<const float*>(P.nbr_q[k]);
            const float qv = q ? q[i] : P.q_uniform;
            cons += P.rho * qv * (z[i] - y[i]);
        }
        out[i] = atb[i] + cons;
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
__global__ void __launch_bounds__(256)
edge_kernel(const EdgeParams P) {
    __shared__ float red[160];
    const EdgeDesc e = P.edges[blockIdx.y];
    const float* __restrict__ xi = reinterpret_cast<const float*>(e.xi);
    const float* __restrict__ xj = reinterpret_cast<const float*>(e.xj);
    float* __restrict__ yi = reinterpret_cast<float*>(e.yi);
    float* __restrict__ yj = reinterpret_cast<float*>(e.yj);
    float* __restrict__ z = reinterpret_cast<float*>(e.z);
    const float* __restrict__ ai_r = reinterpret_cast<const float*>(e.ai);
    const float* __restrict__ aj_r = reinterpret_cast<const float*>(e.aj);
    const float* __restrict__ Wi = reinterpret_cast<const float*>(e.Wi);
    const float* __restrict__ Wj = reinterpret_cast<const float*>(e.Wj);
    const float* __restrict__ qij = reinterpret_cast<const float*>(e.qij);
    const float* __restrict__ qji = reinterpret_cast<const float*>(e.qji);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < P.n; k += (long long)gridDim.x * blockDim.x) {
        const float zo = z[k];
        float xiv = 0.f, yiv = 0.f, xjv = 0.f, yjv = 0.f, ai, aj;
        if (xi) {
            xiv = xi[k]; yiv = yi[k]; ai = xiv + yiv;
            const float ei = xiv - (zo - yiv);
            s3 = fmaf((qij ? qij[k] : P.q_uniform) * ei, ei, s3);
        } else ai = ai_r[k];
        if (xj) {
            xjv = xj[k]; yjv = yj[k]; aj = xjv + yjv;
            const float ej = xjv - (zo - yjv);
            s4 = fmaf((qji ? qji[k] : P.q_uniform) * ej, ej, s4);
        } else aj = aj_r[k];
        float zn;
        if (Wi) { const float wi = Wi[k], wj = Wj[k]; zn = (wi * ai + wj * aj) / (wi + wj); }
        else zn = (ai + aj) / 2.0f;
        if (xi) { const float ri = xiv - zn; yi[k] = yiv + xiv - zn; s0 = fmaf(ri, ri, s0); }
        if (xj) { const float rj = xjv - zn; yj[k] = yjv + xjv - zn; s1 = fmaf(rj, rj, s1); }
        const float dz = zn - zo;
        s2 = fmaf(dz, dz, s2);
        z[k] = zn;
    }
    float v[5] = {s0, s1, s2, s3, s4};
    block_sum<5>(v, red);
    grid_reduce_store<5>(v, P.part + (long long)blockIdx.y * gridDim.x * 5, P.counter + blockIdx.y, blockIdx.x,
                         gridDim.x, P.sums + (long long)blockIdx.y * 5, red);
}

// a = x + y for the cut-edge ends this rank sends
__global__ void __launch_bounds__(256)
pack_kernel(const PackParams P) {
    const PackDesc d = P.items[blockIdx.y];
    const float* __restrict__ x = reinterpret_cast<const float*>(d.x);
    const float* __restrict__ y = reinterpret_cast<const float*>(d.y);
    float* __restrict__ o = reinterpret_cast<float*>(d.out);
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < P.n; k += (long long)gridDim.x * blockDim.x)
        o[k] = x[k] + y[k];
}

// Per-iteration bookkeeping (block_6_admm_loop_ver2.py:232-264): fold the per-edge sums (in G.edges()
// order, fp64, like the reference's Python floats) and the per-node scalars into one history row
//   row = [r2, s2, pri_node[Vg], dual_node[Vg], pen[Vg], mse[Vg], tv[Vg], gn2[Vg], img[Vg]]
// Rows of different ranks are summed by the caller (ncclAllReduce) when nodes are sharded.
__global__ void finalize_kernel(const FinalizeParams P) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int Vg = P.Vg;
    double* row = P.row;
    for (int i = 0; i < 2 + 7 * Vg; ++i) row[i] = 0.0;
    double* pri = row + 2; double* dual = pri + Vg; double* pen = dual + Vg;
    double* mse = pen + Vg; double* tv = mse + Vg; double* gn2 = tv + Vg; double* img = gn2 + Vg;
    const double rho2 = (double)P.rho * (double)P.rho;
    double r2 = 0.0, s2 = 0.0;
    for (int e = 0; e < P.E; ++e) {
        const double* s = P.sums + (long long)e * 5;
        const int gi = P.edge_gi[e], gj = P.edge_gj[e], fl = P.edge_flags[e];
        double acc = 0.0;
        if (fl & 1) { acc += s[0]; pri[gi] += s[0]; pen[gi] += s[3]; }
        if (fl & 2) { acc += s[1]; pri[gj] += s[1]; pen[gj] += s[4]; }
        r2 += acc;
        if (fl & 4) {  // this rank owns the edge's dual residual
            s2 += rho2 * s[2];
            dual[gi] += rho2 * s[2];
            dual[gj] += rho2 * s[2];
        }
    }
    row[0] = r2; row[1] = s2;
    for (int v = 0; v < P.V; ++v) {
        const double* sc = P.scal + (long long)v * NSCAL;
        const int g = P.node_gid[v];
        mse[g] = sc[S_MSE]; tv[g] = sc[S_TV]; gn2[g] = sc[S_GN2]; img[g] = sc[S_IMG];
    }
}

// ---- launchers --------------------------------------------------------------------------------------
static inline int stream_blocks(long long n, int per_thread) {
    long long b = (n + 256LL * per_thread - 1) / (256LL * per_thread);
    if (b < 1) b = 1;
    if (b > 4096) b = 4096;
    return (int)b;
}

cudaError_t launch_tv(const TvParams& P, int nodes, cudaStream_t st) {
    dim3 grid((P.N + TT - 1) / TT, (P.N + TT - 1) / TT, nodes);
    { ProfScope ps(KC_TV, st); tv_fused_kernel<<<grid, 256, 0, st>>>(P); }
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
    { ProfScope ps(KC_SINO_AXPY, st); sino_axpy_kernel<<<dim3((P.D + 255) / 256, P.A1 - P.A0), 256, 0, st>>>(P); }
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
cudaError_t launch_pack(const PackParams& P, int nitems, cudaStream_t st) {
    if (nitems <= 0) return cudaSuccess;
    { ProfScope ps(KC_PACK, st); pack_kernel<<<dim3(stream_blocks(P.n, 4), nitems), 256, 0, st>>>(P); }
    return cudaGetLastError();
}
cudaError_t launch_finalize(const FinalizeParams& P, cudaStream_t st) {
    { ProfScope ps(KC_FINALIZE, st); finalize_kernel<<<1, 32, 0, st>>>(P); }
    return cudaGetLastError();
}

}  // namespace admm
