// tma_selftest.cu -- minimal TMA bring-up check (GPU box): nvcc -gencode arch=compute_100a,code=sm_100a -I../distributed-inverse-problem-admm_b200/csrc
// usage: tma_selftest <variant>   0: map as direct __grid_constant__ param, 1: map inside a params struct, 2: prefetch only,
//                                 3: map read from global memory
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tma.cuh"
using namespace admm;

struct Params { float* out; int rows, cols; int pad; CUtensorMap map; };

__device__ void body(const CUtensorMap* map, float* out, int cols, bool prefetch_only) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    float* tile = reinterpret_cast<float*>(smem);
    if (prefetch_only) {
        if (threadIdx.x == 0) tma_prefetch_3d(map, 0, 0, 0);
        if (threadIdx.x < 64) out[threadIdx.x] = 1.f;
        return;
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 32 * 8 * 4);
        tma_load_3d(tile, map, -4, 1, 0, &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 32 * 8; i += blockDim.x) out[i] = tile[i];
}
__global__ void k_direct(const __grid_constant__ CUtensorMap map, float* out, int cols) { body(&map, out, cols, false); }
__global__ void k_struct(const __grid_constant__ Params P) { body(&P.map, P.out, P.cols, false); }
__global__ void k_prefetch(const __grid_constant__ Params P) { body(&P.map, P.out, P.cols, true); }
__global__ void k_global(const CUtensorMap* map, float* out, int cols) { body(map, out, cols, false); }
// bisect variants
__global__ void k_mbar_only(float* out) {            // 4: mbarrier init / arrive / wait, no copy
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) mbar_expect_tx(&bar, 0);
    mbar_wait(&bar, 0);
    if (threadIdx.x < 64) out[threadIdx.x] = 2.f;
}
__global__ void k_bulk1d(const float* src, float* out) {   // 5: non-tensor bulk copy, 1024 bytes
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    float* tile = reinterpret_cast<float*>(smem);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 1024);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         tma_smem_u32(tile)), "l"(src), "r"(1024), "r"(tma_smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = tile[i];
}
__global__ void k_2d(const __grid_constant__ CUtensorMap map, float* out, int c0, int c1) {   // 6 / 8: 2-D box
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    float* tile = reinterpret_cast<float*>(smem);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, 1024); tma_load_2d(tile, &map, c0, c1, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = tile[i];
}
__global__ void k_elect_hint(const __grid_constant__ CUtensorMap map, float* out) {   // 7: CUTLASS form
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    float* tile = reinterpret_cast<float*>(smem);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned pred = 0;
        asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
        if (pred) {
            mbar_expect_tx(&bar, 1024);
            const unsigned long long hint = 0x1000000000000000ULL;   // EVICT_NORMAL
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                         " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(tma_smem_u32(tile)), "l"(&map),
                         "r"(tma_smem_u32(&bar)), "r"(0), "r"(0), "r"(0), "l"(hint) : "memory");
        }
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int rows = 16, cols = 64, nodes = 2;
    std::vector<float> h((size_t)rows * cols * nodes);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *out;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&out, 32 * 8 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(out, 0, 32 * 8 * 4);
    CUtensorMap map;
    const unsigned long long dims[3] = {(unsigned long long)cols, (unsigned long long)rows, (unsigned long long)nodes};
    const unsigned long long strides[2] = {(unsigned long long)cols * 4, (unsigned long long)rows * cols * 4};
    const unsigned box[3] = {32, 8, 1};
    if (!tma_encode_f32(&map, d, 3, dims, strides, box)) { printf("encode failed\n"); return 2; }
    if (variant >= 4) {
        if (variant == 4) k_mbar_only<<<1, 128>>>(out);
        else if (variant == 5) k_bulk1d<<<1, 128, 1024>>>(d, out);
        else if (variant == 6 || variant == 8) {
            CUtensorMap m2;
            const unsigned long long d2[2] = {(unsigned long long)cols, (unsigned long long)rows * nodes};
            const unsigned long long s2[1] = {(unsigned long long)cols * 4};
            const unsigned b2[2] = {32, 8};
            if (!tma_encode_f32(&m2, d, 2, d2, s2, b2)) { printf("encode2 failed\n"); return 2; }
            k_2d<<<1, 128, 1024>>>(m2, out, variant == 6 ? 4 : -4, 1);
        } else k_elect_hint<<<1, 128, 1024>>>(map, out);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> r(4);
        if (e == cudaSuccess) cudaMemcpy(r.data(), out, 16, cudaMemcpyDeviceToHost);
        printf("variant %d: %s  out[0..3] = %g %g %g %g\n", variant, cudaGetErrorString(e), r[0], r[1], r[2], r[3]);
        return 0;
    }
    if (variant == 0) k_direct<<<1, 128, 32 * 8 * 4>>>(map, out, cols);
    else if (variant == 1 || variant == 2) {
        Params P{}; P.out = out; P.rows = rows; P.cols = cols; P.map = map;
        if (variant == 1) k_struct<<<1, 128, 32 * 8 * 4>>>(P); else k_prefetch<<<1, 128, 32 * 8 * 4>>>(P);
    } else {
        CUtensorMap* dm; cudaMalloc(&dm, sizeof(map)); cudaMemcpy(dm, &map, sizeof(map), cudaMemcpyHostToDevice);
        k_global<<<1, 128, 32 * 8 * 4>>>(dm, out, cols);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d: %s\n", variant, cudaGetErrorString(e));
    if (e == cudaSuccess && variant != 2) {
        std::vector<float> r(32 * 8);
        cudaMemcpy(r.data(), out, r.size() * 4, cudaMemcpyDeviceToHost);
        // box at (col -4, row 1): element (k, j) = row 1+k, col -4+j  (zero when col < 0)
        int bad = 0;
        for (int k = 0; k < 8; ++k) for (int j = 0; j < 32; ++j) {
            const int c = j - 4; const float want = c < 0 ? 0.f : (float)((1 + k) * cols + c);
            if (r[k * 32 + j] != want) ++bad;
        }
        printf("mismatches %d (r[0..3] = %g %g %g %g)\n", bad, r[0], r[1], r[2], r[3]);
    }
    return 0;
}
