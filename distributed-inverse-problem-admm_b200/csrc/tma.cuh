// tma.cuh -- Blackwell bulk-tensor copy (TMA) plumbing for the projector kernels: tensor-map encoding on the host
// (cuTensorMapEncodeTiled reached through the runtime's driver entry point -- the library does not link libcuda) and
// the device-side mbarrier / cp.async.bulk.tensor wrappers (inline PTX, sm_100a).
//
// Uses (projector.cu):
//   * K1 forward: the y-dominant image tile of a slab lands in shared memory by ONE cp.async.bulk.tensor.3d issued by
//     one thread (out-of-bounds columns / rows are zero-filled by the copy unit: no border tests, no halo stores at
//     the image border), and the NEXT slab's boxes of every operand stream are pulled towards L2 by
//     cp.async.bulk.prefetch.tensor while the current slab is sampled (one instruction per stream and slab).
//   * K2 back-projection: every angle's detector window of a pixel tile is one 2-D box copy from the sinogram
//     (zero-filled outside [0, D)), issued by the thread that computed the window; the tile's threads wait on the
//     mbarrier instead of running a staging loop + barrier.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace admm {

// ---- host: tensor maps ------------------------------------------------------------------------------
// fp32 tensor of rank 2 or 3, dims[0] fastest; strides_bytes[k] = byte stride of dims[k+1] (multiples of 16); box[k]
// = elements per box along dims[k] (<= 256, box[0]*4 a multiple of 16).  No swizzle, no interleave, zero OOB fill.
// Returns false (and leaves `map` untouched) when the driver entry point is missing or the arguments are not
// encodable (unaligned base / strides): callers then take the non-TMA path.
bool tma_encode_f32(CUtensorMap* map, const void* base, int rank, const unsigned long long* dims,
                    const unsigned long long* strides_bytes, const unsigned* box);

// ---- device -------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned tma_smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(count) : "memory");
}
// make the barrier initialisation visible to the async proxy (the copy unit arrives on it)
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// order this thread's earlier generic-proxy shared-memory accesses before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(tma_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            tma_smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(tma_smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            tma_smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(tma_smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace admm
