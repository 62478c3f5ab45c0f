// epilogue.cuh -- per-pixel form of the back-projector epilogues (same semantics as back_tile_kernel's fused ones), for
// the operator variants that are not on the hot path: dense matrices (dense.cu), rotate-and-sum projector (rotsum.cu).
//   BACK_PLAIN / BACK_COLNORM2 : out = acc
//   BACK_HP     : out = acc + rhoD v + mu K^T K v ; sums <v, out>, <out, out>
//   BACK_RESID0 : r = rhs0 + tvterm - (acc + rhoD v + mu K^T K v) ; p = r ; sums <r, r>
#pragma once
#include "solver_kernels.cuh"

namespace admm {

template <int MODE>
__device__ __forceinline__ void pixel_epilogue(const BackParams& P, int node, long long nb, long long c, float acc,
                                               float& dsum, float& dsum3) {
    if (MODE == BACK_PLAIN || MODE == BACK_COLNORM2) {
        P.out[nb + c] = acc;
        return;
    }
    const int N = P.N;
    const float* __restrict__ v = P.v + nb;
    const int ix = (int)(c / N), iy = (int)(c % N);
    const float cv = v[c];
    float lu = 0.f, ld = 0.f, ll = 0.f, lr = 0.f;   // same association as the projector's epilogue
    if (ix >= 1) lu = cv - v[c - N];
    if (ix + 1 < N) ld = cv - v[c + N];
    if (iy >= 1) ll = cv - v[c - 1];
    if (iy + 1 < N) lr = cv - v[c + 1];
    const float lap = (lu + ld) + (ll + lr);
    const float dd = P.rhoD_vec ? P.rhoD_vec[nb + c] : P.rhoD_s[node];
    const float hv = acc + fmaf(dd, cv, P.mu * lap);
    if (MODE == BACK_HP) {
        P.out[nb + c] = hv;
        dsum = cv * hv; dsum3 = hv * hv;
    } else {
        const float rr = (P.rhs0[nb + c] + P.tvterm[nb + c]) - hv;
        P.out[nb + c] = rr;
        P.p_out[nb + c] = rr;
        dsum = rr * rr;
    }
}

// block / grid reduction of the epilogue sums of one node (grid.x blocks per node, node on grid.y)
template <int MODE>
__device__ __forceinline__ void pixel_epilogue_reduce(const BackParams& P, int node, float dsum, float dsum3, float* red) {
    if (MODE == BACK_HP) {
        float vs[3] = {dsum, 0.f, dsum3};
        block_sum<3>(vs, red);
        grid_reduce_store<3>(vs, P.part + (long long)blockIdx.y * gridDim.x * 3, P.counter + blockIdx.y, blockIdx.x,
                             gridDim.x, P.scal + (long long)node * NSCAL + P.dot_slot, red);
    } else if (MODE == BACK_RESID0) {
        float vs[1] = {dsum};
        block_sum<1>(vs, red);
        grid_reduce_store<1>(vs, P.part + (long long)blockIdx.y * gridDim.x, P.counter + blockIdx.y, blockIdx.x,
                             gridDim.x, P.scal + (long long)node * NSCAL + P.dot_slot, red);
    }
}

}  // namespace admm
