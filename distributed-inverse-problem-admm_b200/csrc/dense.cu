// dense.cu -- dense-matrix operator for the reference's literal calling convention: `A_dense_list[i]` is an
// (m_i, n) ndarray (block_2_load_odl_data.py:68-96; used as `Ai @ x`, `Ai.T @ r`, `Ai.shape[1]` at
// block_6_admm_loop_ver2.py:26,145,193 and as np.sum(A*A, axis=0) at block_3_graph_and_precisions.py:22).
// "dense ndarray still accepted for tiny N (uploaded, but then no projector kernel)".  Plain,
// trivially correct matvec kernels behind the same launch_forward / launch_back contract as the matrix-free
// projector pair (same parameter blocks, same epilogue semantics, same deterministic reductions), so the solver
// above them does not change.  Not a performance path: a dense A_i is 4 m_i n bytes (377 MB per node at cfg 1).
#include "epilogue.cuh"

namespace admm {

// out[row] = sum_c A[row][c] * img[node(row)][c]      one warp per matrix row, fixed summation order
__global__ void __launch_bounds__(256)
dense_fwd_kernel(const float* __restrict__ A, const int* __restrict__ anode, const float* __restrict__ img,
                 long long img_stride, int node0, int row0, int rows, long long n, float* __restrict__ out,
                 const NodeCtl* ctl) {
    const int row = row0 + blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= row0 + rows) return;
    const int node = anode[row];
    if (ctl && !ctl[node].active) return;
    const float* __restrict__ a = A + (long long)row * n;
    const float* __restrict__ x = img + (long long)(node - node0) * img_stride;
    float acc = 0.f;
    for (long long c = threadIdx.x & 31; c < n; c += 32) acc = fmaf(a[c], x[c], acc);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) out[row] = acc;
}

// one thread per pixel column: at = sum_rows prec * A[row][c] * q[row]  (or sum A^2), then the same epilogues as
// back_tile_kernel: H v = at + rhoD v + mu K^T K v with <v,Hv>, <Hv,Hv> ; r = rhs0 + tvterm - H v, p = r, <r,r>
template <int MODE>
__global__ void __launch_bounds__(256)
dense_back_kernel(const float* __restrict__ A, const BackParams P) {
    __shared__ __align__(16) float red[96];
    const int node = P.node0 + blockIdx.y;
    if (P.ctl && !P.ctl[node].active) return;
    const long long n = (long long)P.N * P.N;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nb = (long long)blockIdx.y * P.stride;
    const int r0 = P.aptr[node], r1 = P.aptr[node + 1];
    const float prec = (MODE == BACK_COLNORM2 || P.prec == nullptr) ? 1.f : P.prec[node];
    float dsum = 0.f, dsum3 = 0.f;
    if (c < n) {
        float acc = 0.f;
        for (int row = r0; row < r1; ++row) {
            const float a = A[(long long)row * n + c];
            acc = (MODE == BACK_COLNORM2) ? fmaf(a, a, acc) : fmaf(a * prec, P.q[row], acc);
        }
        pixel_epilogue<MODE>(P, node, nb, c, acc, dsum, dsum3);
    }
    pixel_epilogue_reduce<MODE>(P, node, dsum, dsum3, red);
}

cudaError_t launch_dense_forward(const float* A, const int* anode, const FwdParams& P, int nodes, const FwdReduceParams& R,
                                 cudaStream_t st) {
    if (P.mode != 0) return cudaErrorInvalidValue;   // the fused CG staging belongs to the strip projector
    const int rows = R.A1 - R.A0;
    if (rows <= 0) return cudaSuccess;
    (void)nodes;
    {
        ProfScope ps(KC_FWD, st);
        dense_fwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(A, anode, P.img, P.img_stride, P.node0, R.A0, rows,
                                                         (long long)P.N * P.N, R.out, P.ctl);
    }
    return cudaGetLastError();
}

cudaError_t launch_dense_back(const float* A, int mode, const BackParams& P, int nodes, cudaStream_t st) {
    const long long n = (long long)P.N * P.N;
    dim3 grid((unsigned)((n + 255) / 256), nodes);
    switch (mode) {
        case BACK_PLAIN: { ProfScope ps(KC_BACK_PLAIN, st); dense_back_kernel<BACK_PLAIN><<<grid, 256, 0, st>>>(A, P); } break;
        case BACK_HP: { ProfScope ps(KC_BACK_HP, st); dense_back_kernel<BACK_HP><<<grid, 256, 0, st>>>(A, P); } break;
        case BACK_RESID0: { ProfScope ps(KC_BACK_RESID0, st); dense_back_kernel<BACK_RESID0><<<grid, 256, 0, st>>>(A, P); } break;
        case BACK_COLNORM2: { ProfScope ps(KC_COLNORM, st); dense_back_kernel<BACK_COLNORM2><<<grid, 256, 0, st>>>(A, P); } break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace admm
