// common.cuh -- shared device helpers for the sm_100a TV-ADMM tomography kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace admm {

extern std::atomic<long long> g_launch_count;  // kernels launched by this library (bench evidence)

// kernel classes for the optional CUDA-event profiler (admm_profile_* in the C ABI)
enum KClass : int {
    KC_FWD = 0, KC_FWD_REDUCE, KC_BACK_PLAIN, KC_BACK_HP, KC_BACK_RESID0, KC_COLNORM, KC_TV, KC_CG_UPDATE,
    KC_P_UPDATE, KC_SINO_AXPY, KC_SINO_RESID, KC_RHS0, KC_EDGE, KC_PACK, KC_FINALIZE, KC_FWD_FUSED, KC_ACCEPT, KC_COUNT
};
void prof_mark(int kc, cudaStream_t st, bool begin);
extern std::atomic<bool> g_prof_on;
struct ProfScope {  // records a CUDA event pair around one launch on its own stream when profiling is on
    int kc; cudaStream_t st;
    ProfScope(int k, cudaStream_t s) : kc(k), st(s) { ++g_launch_count; if (g_prof_on) prof_mark(kc, st, true); }
    ~ProfScope() { if (g_prof_on) prof_mark(kc, st, false); }
};

// Per-angle geometry record (built on the host in fp64 from the fp32-rounded (cos, sin) table so the CUDA
// kernels and the fp64 oracle take bit-identical dominant-axis decisions).
// Pixel (ix, iy) projects to fractional detector bin
//     tau = cj + (ix - cx) * ct + (iy - cx) * st,   cx = (N-1)/2, cj = (D-1)/2,
// with ct = cos * h/ds, st = sin * h/ds (h = 2/N pixel size, ds = det_w/D bin size).
struct AngleRec {
    double ct, st;
    float inv_major;  // 1 / major,  major = xdom ? ct : st
    float slope;      // minor / major
    float wgt;        // h / |a|, a = xdom ? cos : sin   (Joseph step length)
    float inv_om;     // 1 / |major| = 1 / omega  (hat half-width in bins is omega)
    int xdom;         // 1: |cos| > |sin| -> step along iy, interpolate along ix
    float inv_slope;  // major / minor, 0 when |slope| <= 1e-6 (ray parallel to the step axis)
};
static_assert(sizeof(AngleRec) == 40, "AngleRec layout is part of the C ABI");

// Per-node control word of the a14 accept / tighten-and-retry rule (block_6_admm_loop_ver2.py:100-176), kept on the
// device so that the retry passes need no host round trip: kernels launched with `masked` skip nodes whose
// `active` flag is 0.  `wpar` is the node's TV-multiplier ping-pong parity (0: w0 current), flipped by the TV kernel
// itself, so nodes that took different numbers of passes stay consistent.
struct NodeCtl { int active, tries, wpar, pad; };
static_assert(sizeof(NodeCtl) == 16, "NodeCtl layout is part of the C ABI");

constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23 : (v + kMagic) - kMagic == rint(v) for |v| < 2^22
constexpr int kMagicBits = 0x4B400000;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of K floats per thread (1-D or 2-D blocks); result valid in linear thread 0.  red: K * 32 floats.
template <int K>
__device__ __forceinline__ void block_sum(float (&v)[K], float* red) {
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5, nw = (blockDim.x * blockDim.y + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) red[k * 32 + wid] = v[k];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float t = (lane < nw) ? red[k * 32 + lane] : 0.f;
            v[k] = warp_sum(t);
        }
    }
}

// Deterministic grid reduction ("last block done"): every block stores K partials, the last block to
// arrive sums all partials of its group in a fixed order in fp64 and writes out[0..K).  The counter is
// reset by the last block so the workspace is reusable by the next launch on the same stream.
//   part:    [nblk][K] floats for this group;  counter: one unsigned for this group.
// Must be called by all threads of the block; v[] valid in thread 0 (output of block_sum).
// Returns true in every thread of the last block (after the result is stored), false elsewhere.
// `slots` (optional): out[slots[k]] receives value k instead of out[k].
template <int K>
__device__ __forceinline__ bool grid_reduce_store(const float (&v)[K], float* part, unsigned* counter,
                                                  int blk, int nblk, double* out, float* red,
                                                  const int* slots = nullptr) {
    __shared__ int s_last;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) part[(size_t)blk * K + k] = v[k];
        __threadfence();
        const unsigned t = atomicAdd(counter, 1u);
        s_last = (t == (unsigned)nblk - 1u);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        // fixed-order fp64 accumulation: thread t sums partials t, t+T, ... then a fixed tree
        double acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.0;
        for (int b = tid; b < nblk; b += nthr) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += (double)__ldcg(&part[(size_t)b * K + k]);
        }
        double* dred = reinterpret_cast<double*>(red);  // red holds >= 32*K floats -> reuse per k
        const int lane = tid & 31, wid = tid >> 5, nw = (nthr + 31) >> 5;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double a = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            __syncthreads();
            if (lane == 0) dred[wid] = a;
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int w = 0; w < nw; ++w) s += dred[w];
                out[slots ? slots[k] : k] = s;
            }
        }
        if (tid == 0) *counter = 0u;
    }
    return s_last != 0;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f32_4(unsigned addr) {  // [addr + 4]
    float v;
    asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(v) : "r"(addr));
    return v;
}

// both taps of a linear interpolation through ONE address register: a = [addr], b = [addr + 4]
__device__ __forceinline__ void lds_pair(unsigned addr, float& a, float& b) {
    asm volatile("ld.shared.f32 %0, [%2];\n\tld.shared.f32 %1, [%2+4];" : "=f"(a), "=f"(b) : "r"(addr));
}
// ---- Blackwell packed fp32 (sm_100+): two IEEE fp32 operations per issue slot (SASS FFMA2 / FADD2 / FMUL2) on a
// 64-bit register pair {lo, hi}.  Each lane rounds exactly like the scalar instruction, so a packed loop is
// bit-identical to its scalar form.  ptxas folds a pair built from one scalar (pack2(s, s)) into the broadcast
// operand form `R.F32`, and constant pairs into uniform-register / immediate operands.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 splat2(float s) { return pack2(s, s); }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2_rm(f32x2 a, f32x2 b) {   // round towards -inf (floor through the magic number)
    f32x2 d;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// make a value opaque to the optimiser (stops it from re-deriving / re-associating loop invariants per use)
__device__ __forceinline__ void opaque(unsigned& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void opaque(float& v) { asm volatile("" : "+f"(v)); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace admm
