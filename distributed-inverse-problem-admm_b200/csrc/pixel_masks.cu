// pixel_masks.cu -- (f)-1: the per-pixel graph masks of block_3_graph_and_precisions.py:62-187 on the device, bit-packed.
//
// For every pixel p the reference builds a graph on the V nodes from the V x V weights q_ij[p] (strategy "knn": k largest
// neighbours per node, symmetrised, plus the edges of a maximum spanning tree if that is not connected, :62-110; "mst":
// the maximum spanning tree, :114-130; "chain": a random permutation chain, :134-146) in a Python loop over pixels with
// networkx.  Here one thread owns one pixel: q_ij[p] is formed from the W vectors exactly as make_precisions does in
// float32 (:27-39), the graph work runs on bitmasks in registers / local memory, and the result is keep_bits[i][p], a
// 32-bit word whose bit j is keep[i, j, p] (V <= 32).  Integer work: bit-exact against the reference.
//   * spanning tree = Kruskal as networkx runs it: edges of the complete graph in (0,1),(0,2),...,(V-2,V-1) order, stable
//     sort by descending weight (ties keep that order), union-find;
//   * k-NN: np.argpartition's choice among EQUAL weights is unspecified; here the smaller index wins.  With distinct
//     weights (the generic case) the selection is unique;
//   * chain: the permutations come from the host (np.random.default_rng(seed).permutation(V) per pixel, :136) -- the
//     PCG64 stream is not restated -- and are scattered here.
#include "solver_kernels.cuh"

namespace admm {

constexpr int MAXV = 32;

struct MaskParams {
    const float* W;          // [V][n] column norms^2 (make_precisions' Wi_list, floored at 1e-12)
    const unsigned char* perm;  // [n][V] chain permutations (strategy 2) or nullptr
    unsigned* keep;          // [V][n] bit-packed rows
    long long n;
    int V, strategy, k, harmonic;   // strategy 0 knn, 1 mst, 2 chain
};

__device__ __forceinline__ float q_weight(float wi, float wj, int harmonic) {
    // block_3_graph_and_precisions.py:27-39 in float32 like NumPy evaluates it (no fused multiply-add)
    float q = harmonic ? __fdiv_rn(__fmul_rn(wi, wj), __fadd_rn(wi, wj)) : __fmul_rn(0.5f, __fadd_rn(wi, wj));
    return fmaxf(q, 1e-12f);
}

__device__ __forceinline__ int uf_find(unsigned char* parent, int a) {
    while (parent[a] != a) a = parent[a];
    return a;
}

// maximum spanning tree of the complete graph (Kruskal, networkx order) -> adjacency bit rows added to adj[]
__device__ void add_max_spanning_tree(const float* w, int V, int harmonic, unsigned* adj) {
    unsigned char parent[MAXV];
    for (int i = 0; i < V; ++i) parent[i] = (unsigned char)i;
    // "used" bitmask per i over j > i: edges already taken out of the sorted order
    unsigned done[MAXV];
    for (int i = 0; i < V; ++i) done[i] = 0u;
    int taken = 0;
    const int E = V * (V - 1) / 2;
    for (int step = 0; step < E && taken < V - 1; ++step) {
        // next edge of the stable descending sort: the largest remaining weight, first in (i, j) order among equals
        float best = -1.f;
        int bi = -1, bj = -1;
        for (int i = 0; i < V; ++i)
            for (int j = i + 1; j < V; ++j) {
                if (done[i] & (1u << j)) continue;
                const float q = q_weight(w[i], w[j], harmonic);
                if (q > best) { best = q; bi = i; bj = j; }
            }
        done[bi] |= 1u << bj;
        const int ra = uf_find(parent, bi), rb = uf_find(parent, bj);
        if (ra != rb) {
            parent[ra] = (unsigned char)rb;
            adj[bi] |= 1u << bj;
            adj[bj] |= 1u << bi;
            ++taken;
        }
    }
}

__device__ bool connected(const unsigned* adj, int V) {
    unsigned seen = 1u, frontier = 1u;
    const unsigned all = (V == 32) ? 0xffffffffu : ((1u << V) - 1u);
    while (frontier) {
        unsigned next = 0u;
        for (int i = 0; i < V; ++i)
            if (frontier & (1u << i)) next |= adj[i];
        frontier = next & ~seen;
        seen |= next;
    }
    return (seen & all) == all;
}

__global__ void __launch_bounds__(128) pixel_mask_kernel(const MaskParams P) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n) return;
    const int V = P.V;
    unsigned adj[MAXV];
    for (int i = 0; i < V; ++i) adj[i] = 0u;
    if (P.strategy == 2) {
        const unsigned char* o = P.perm + p * V;
        for (int t = 0; t + 1 < V; ++t) {
            const int u = o[t], v = o[t + 1];
            adj[u] |= 1u << v;
            adj[v] |= 1u << u;
        }
    } else {
        float w[MAXV];
        for (int i = 0; i < V; ++i) w[i] = P.W[(long long)i * P.n + p];
        if (P.strategy == 1) {
            add_max_spanning_tree(w, V, P.harmonic, adj);
        } else {
            const int keff = min(P.k, V - 1);
            for (int i = 0; i < V; ++i) {           // k largest neighbours of node i (:74-80)
                unsigned sel = 0u;
                for (int t = 0; t < keff; ++t) {
                    float best = -1.f;
                    int bj = -1;
                    for (int j = 0; j < V; ++j) {
                        if (j == i || (sel & (1u << j))) continue;
                        const float q = q_weight(w[i], w[j], P.harmonic);
                        if (q > best) { best = q; bj = j; }
                    }
                    sel |= 1u << bj;
                }
                adj[i] |= sel;
            }
            for (int i = 0; i < V; ++i)              // symmetrise (:82)
                for (int j = 0; j < V; ++j)
                    if (adj[i] & (1u << j)) adj[j] |= 1u << i;
            if (!connected(adj, V)) add_max_spanning_tree(w, V, P.harmonic, adj);   // :93-103
        }
    }
    for (int i = 0; i < V; ++i) P.keep[(long long)i * P.n + p] = adj[i];
}

cudaError_t launch_pixel_masks(const MaskParams& P, cudaStream_t st) {
    if (P.n <= 0) return cudaSuccess;
    pixel_mask_kernel<<<(unsigned)((P.n + 127) / 128), 128, 0, st>>>(P);
    ++g_launch_count;
    return cudaGetLastError();
}

}  // namespace admm

// ---- C ABI ---------------------------------------------------------------------------------------------------------
#include "../../include/admm_b200.h"

extern "C" int admm_pixel_masks(int V, long long n, int strategy, int k, int harmonic, const float* d_W,
                                const unsigned char* d_perm, unsigned* d_keep_bits, void* stream) {
    if (V < 1 || V > admm::MAXV || n < 1 || strategy < 0 || strategy > 2 || !d_keep_bits) return ADMM_ERR_ARG;
    if (strategy == 2 ? (d_perm == nullptr) : (d_W == nullptr)) return ADMM_ERR_ARG;
    admm::MaskParams P{};
    P.W = d_W; P.perm = d_perm; P.keep = d_keep_bits; P.n = n; P.V = V; P.strategy = strategy; P.k = k; P.harmonic = harmonic;
    return admm::launch_pixel_masks(P, (cudaStream_t)stream) == cudaSuccess ? ADMM_OK : ADMM_ERR_CUDA;
}
