// tv_helpers.cu -- device versions of the reference's NumPy TV helpers (block_4_tv_helpers.py:17-46,
// block_4_tv_helpers_with_plot.py:23-46) behind host-buffer C-ABI entry points.  They work in fp64 like the
// reference (float64 NumPy): differences, IEEE sqrt and division only, so results are bit-identical to NumPy.
#include <string>

#include "../../include/admm_b200.h"
#include "common.cuh"

namespace admm {

__global__ void grad2d_kernel(const double* __restrict__ x, int N, double* __restrict__ gx, double* __restrict__ gy) {
    const long long n = (long long)N * N;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(k / N), c = (int)(k % N);
        gx[k] = (r < N - 1) ? x[k + N] - x[k] : 0.0;   // block_4_tv_helpers.py:21
        gy[k] = (c < N - 1) ? x[k + 1] - x[k] : 0.0;   // :22
    }
}

// mode 0: block_4_tv_helpers.py:25-35 as shipped ; mode 1: exact K^T
__global__ void div2d_kernel(const double* __restrict__ px, const double* __restrict__ py, int N, int mode,
                             double* __restrict__ out) {
    const long long n = (long long)N * N;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(k / N), c = (int)(k % N);
        if (mode == 1) {
            double v = 0.0;
            if (r >= 1) v += px[k - N];
            if (r < N - 1) v -= px[k];
            if (c >= 1) v += py[k - 1];
            if (c < N - 1) v -= py[k];
            out[k] = v;
        } else {
            double div = 0.0;
            if (r == 0) div -= px[k];
            else if (r == N - 1) div += px[k - N];
            else div += px[k] - px[k - N];
            if (c == 0) div -= py[k];
            else if (c == N - 1) div += py[k - 1];
            else div += py[k] - py[k - 1];
            out[k] = -div;
        }
    }
}

// p = g / |g| where |g| > eps else 0   (block_4_tv_helpers.py:38-45); mag optionally returned (edge map)
__global__ void normgrad_kernel(const double* __restrict__ x, int N, double eps, double* __restrict__ px,
                                double* __restrict__ py, double* __restrict__ mag) {
    const long long n = (long long)N * N;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(k / N), c = (int)(k % N);
        const double gx = (r < N - 1) ? x[k + N] - x[k] : 0.0;
        const double gy = (c < N - 1) ? x[k + 1] - x[k] : 0.0;
        // no FMA contraction: NumPy rounds gx*gx, gy*gy and their sum separately
        const double m = sqrt(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)));
        if (mag) mag[k] = m;
        if (px) {
            px[k] = (m > eps) ? gx / m : 0.0;
            py[k] = (m > eps) ? gy / m : 0.0;
        }
    }
}

}  // namespace admm

using namespace admm;

static thread_local std::string g_err2;
#define CKH(expr)                                     \
    do {                                              \
        cudaError_t _e = (expr);                      \
        if (_e != cudaSuccess) {                      \
            for (int _i = 0; _i < nb; ++_i) cudaFree(bufs[_i]); \
            return ADMM_ERR_CUDA;                     \
        }                                             \
    } while (0)

static int blocks_for(long long n) {
    long long b = (n + 255) / 256;
    return (int)(b < 1 ? 1 : (b > 8192 ? 8192 : b));
}

extern "C" int admm_grad2d_host(int N, const double* h_x, double* h_gx, double* h_gy) {
    if (N < 1 || !h_x || !h_gx || !h_gy) return ADMM_ERR_ARG;
    const size_t n = (size_t)N * N, B = n * sizeof(double);
    double* bufs[3] = {nullptr, nullptr, nullptr};
    const int nb = 3;
    for (int i = 0; i < nb; ++i) CKH(cudaMalloc(&bufs[i], B));
    CKH(cudaMemcpy(bufs[0], h_x, B, cudaMemcpyHostToDevice));
    { ProfScope ps(KC_TV, nullptr); grad2d_kernel<<<blocks_for(n), 256>>>(bufs[0], N, bufs[1], bufs[2]); }
    CKH(cudaGetLastError());
    CKH(cudaMemcpy(h_gx, bufs[1], B, cudaMemcpyDeviceToHost));
    CKH(cudaMemcpy(h_gy, bufs[2], B, cudaMemcpyDeviceToHost));
    for (int i = 0; i < nb; ++i) cudaFree(bufs[i]);
    return ADMM_OK;
}

extern "C" int admm_div2d_host(int N, const double* h_px, const double* h_py, int exact_adjoint, double* h_out) {
    if (N < 2 || !h_px || !h_py || !h_out) return ADMM_ERR_ARG;
    const size_t n = (size_t)N * N, B = n * sizeof(double);
    double* bufs[3] = {nullptr, nullptr, nullptr};
    const int nb = 3;
    for (int i = 0; i < nb; ++i) CKH(cudaMalloc(&bufs[i], B));
    CKH(cudaMemcpy(bufs[0], h_px, B, cudaMemcpyHostToDevice));
    CKH(cudaMemcpy(bufs[1], h_py, B, cudaMemcpyHostToDevice));
    { ProfScope ps(KC_TV, nullptr); div2d_kernel<<<blocks_for(n), 256>>>(bufs[0], bufs[1], N, exact_adjoint ? 1 : 0, bufs[2]); }
    CKH(cudaGetLastError());
    CKH(cudaMemcpy(h_out, bufs[2], B, cudaMemcpyDeviceToHost));
    for (int i = 0; i < nb; ++i) cudaFree(bufs[i]);
    return ADMM_OK;
}

// kt_subgrad_isotropic_tv_from_x (block_4_tv_helpers.py:37-46); h_mag (optional) = |grad x| (edge map, _with_plot:23-46)
extern "C" int admm_kt_subgrad_host(int N, const double* h_x, double eps, int exact_adjoint, double* h_out,
                                    double* h_mag) {
    if (N < 2 || !h_x || (!h_out && !h_mag)) return ADMM_ERR_ARG;
    const size_t n = (size_t)N * N, B = n * sizeof(double);
    double* bufs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    const int nb = 5;
    for (int i = 0; i < nb; ++i) CKH(cudaMalloc(&bufs[i], B));
    CKH(cudaMemcpy(bufs[0], h_x, B, cudaMemcpyHostToDevice));
    { ProfScope ps(KC_TV, nullptr); normgrad_kernel<<<blocks_for(n), 256>>>(bufs[0], N, eps, bufs[1], bufs[2], bufs[4]); }
    if (h_out) {
        { ProfScope ps(KC_TV, nullptr); div2d_kernel<<<blocks_for(n), 256>>>(bufs[1], bufs[2], N, exact_adjoint ? 1 : 0, bufs[3]); }
        CKH(cudaGetLastError());
        CKH(cudaMemcpy(h_out, bufs[3], B, cudaMemcpyDeviceToHost));
    }
    if (h_mag) CKH(cudaMemcpy(h_mag, bufs[4], B, cudaMemcpyDeviceToHost));
    for (int i = 0; i < nb; ++i) cudaFree(bufs[i]);
    return ADMM_OK;
}
