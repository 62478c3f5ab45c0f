// rotsum.cu -- (f)-2: the "skimage-flavoured" projector variant behind the same operator interface
// (Gen_Sino_Partitioned.py:133 pins impl='skimage' for generate_sinogram).  skimage.transform.radon(circle=False) rotates
// the sqrt(2)-padded image with bilinear interpolation and sums its columns; restated as ray marching on the rotated
// pixel grid: for detector bin j (centre s_j) the samples are r_k = s_j (cos, sin) + t_k (-sin, cos),
// t_k = (k - (P-1)/2) h, k < P = ceil(sqrt(2) N); the image is interpolated bilinearly there (zero outside) and the bin
// gets h * sum_k.  The adjoint is the exact transpose as an atomics-free gather over the <= 4 x 4 lattice samples whose
// bilinear footprint covers the pixel.  A second discretisation to bracket the "ODL output" ambiguity,
// not a hot path: plain kernels, fp64 geometry.
#include "epilogue.cuh"

namespace admm {

__device__ __forceinline__ int rs_steps(int N) { return (int)ceil(1.4142135623730951 * (double)N); }

__global__ void __launch_bounds__(128)
rs_fwd_kernel(const float2* __restrict__ cs, const int* __restrict__ anode, const float* __restrict__ img,
              long long img_stride, int node0, int row0, int N, int D, double det_w, float* __restrict__ out,
              const NodeCtl* ctl) {
    const int a = row0 + blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= D) return;
    const int node = anode[a];
    if (ctl && !ctl[node].active) return;
    const float* __restrict__ x = img + (long long)(node - node0) * img_stride;
    const double c = (double)cs[a].x, s = (double)cs[a].y;
    const double h = 2.0 / N, ds = det_w / D, x0 = -1.0 + 0.5 * h;
    const int P = rs_steps(N);
    const double sj = -0.5 * det_w + (j + 0.5) * ds, t0 = -0.5 * (P - 1) * h;
    // pixel coordinates of sample k: (px0 - k s, py0 + k c)
    const double px0 = ((sj * c - t0 * s) - x0) / h, py0 = ((sj * s + t0 * c) - x0) / h;
    float acc = 0.f;
    for (int k = 0; k < P; ++k) {
        const double px = px0 - k * s, py = py0 + k * c;
        const double fx0 = floor(px), fy0 = floor(py);
        const int i0 = (int)fx0, j0 = (int)fy0;
        if (i0 < -1 || i0 >= N || j0 < -1 || j0 >= N) continue;
        const float fx = (float)(px - fx0), fy = (float)(py - fy0);
        const bool ia = i0 >= 0, ib = i0 + 1 < N, ja = j0 >= 0, jb = j0 + 1 < N;
        const long long g = (long long)i0 * N + j0;
        const float v00 = (ia && ja) ? x[g] : 0.f, v01 = (ia && jb) ? x[g + 1] : 0.f;
        const float v10 = (ib && ja) ? x[g + N] : 0.f, v11 = (ib && jb) ? x[g + N + 1] : 0.f;
        acc += (1.f - fx) * ((1.f - fy) * v00 + fy * v01) + fx * ((1.f - fy) * v10 + fy * v11);
    }
    out[(long long)a * D + j] = acc * (float)h;
}

template <int MODE>
__global__ void __launch_bounds__(256)
rs_back_kernel(const float2* __restrict__ cs, const BackParams P, double det_w) {
    __shared__ __align__(16) float red[96];
    const int node = P.node0 + blockIdx.y;
    if (P.ctl && !P.ctl[node].active) return;
    const int N = P.N, D = P.D;
    const long long n = (long long)N * N;
    const long long cpx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nb = (long long)blockIdx.y * P.stride;
    const float prec = (MODE == BACK_COLNORM2 || P.prec == nullptr) ? 1.f : P.prec[node];
    float dsum = 0.f, dsum3 = 0.f;
    if (cpx < n) {
        const int ix = (int)(cpx / N), iy = (int)(cpx % N);
        const double h = 2.0 / N, ds = det_w / D, x0 = -1.0 + 0.5 * h, r = ds / h;
        const int Pn = rs_steps(N);
        const double x = x0 + ix * h, y = x0 + iy * h;
        const double rad = 1.4142135623730951 + 1e-6;
        float acc = 0.f;
        for (int a = P.aptr[node]; a < P.aptr[node + 1]; ++a) {
            const double c = (double)cs[a].x, s = (double)cs[a].y;
            const double nrm2 = c * c + s * s;   // fp32-rounded trig: the exact inverse of the forward's sample map
            const double tau = ((x * c + y * s) / nrm2 + 0.5 * det_w) / ds - 0.5;
            const double kap = ((-x * s + y * c) / nrm2) / h + 0.5 * (Pn - 1);
            const int jlo = max((int)ceil(tau - rad / r), 0), jhi = min((int)floor(tau + rad / r), D - 1);
            const int klo = max((int)ceil(kap - rad), 0), khi = min((int)floor(kap + rad), Pn - 1);
            for (int j = jlo; j <= jhi; ++j) {
                float wsum = 0.f;
                for (int k = klo; k <= khi; ++k) {
                    const double du = (j - tau) * r, dv = (double)k - kap;
                    const float wx = 1.f - (float)fabs(du * c - dv * s), wy = 1.f - (float)fabs(du * s + dv * c);
                    if (wx > 0.f && wy > 0.f) wsum = fmaf(wx, wy, wsum);
                }
                wsum *= (float)h;
                acc = (MODE == BACK_COLNORM2) ? fmaf(wsum, wsum, acc) : fmaf(wsum * prec, P.q[(long long)a * D + j], acc);
            }
        }
        pixel_epilogue<MODE>(P, node, nb, cpx, acc, dsum, dsum3);
    }
    pixel_epilogue_reduce<MODE>(P, node, dsum, dsum3, red);
}

cudaError_t launch_rs_forward(const float2* cs, const int* anode, double det_w, const FwdParams& P, int nodes,
                              const FwdReduceParams& R, cudaStream_t st) {
    if (P.mode != 0) return cudaErrorInvalidValue;   // the fused CG staging belongs to the strip projector
    const int rows = R.A1 - R.A0;
    if (rows <= 0) return cudaSuccess;
    (void)nodes;
    {
        ProfScope ps(KC_FWD, st);
        rs_fwd_kernel<<<dim3((P.D + 127) / 128, rows), 128, 0, st>>>(cs, anode, P.img, P.img_stride, P.node0, R.A0, P.N, P.D,
                                                                    det_w, R.out, P.ctl);
    }
    return cudaGetLastError();
}

cudaError_t launch_rs_back(const float2* cs, double det_w, int mode, const BackParams& P, int nodes, cudaStream_t st) {
    const long long n = (long long)P.N * P.N;
    dim3 grid((unsigned)((n + 255) / 256), nodes);
    switch (mode) {
        case BACK_PLAIN: { ProfScope ps(KC_BACK_PLAIN, st); rs_back_kernel<BACK_PLAIN><<<grid, 256, 0, st>>>(cs, P, det_w); } break;
        case BACK_HP: { ProfScope ps(KC_BACK_HP, st); rs_back_kernel<BACK_HP><<<grid, 256, 0, st>>>(cs, P, det_w); } break;
        case BACK_RESID0: { ProfScope ps(KC_BACK_RESID0, st); rs_back_kernel<BACK_RESID0><<<grid, 256, 0, st>>>(cs, P, det_w); } break;
        case BACK_COLNORM2: { ProfScope ps(KC_COLNORM, st); rs_back_kernel<BACK_COLNORM2><<<grid, 256, 0, st>>>(cs, P, det_w); } break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace admm
