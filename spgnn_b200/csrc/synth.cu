// Synthetic airway trees generated on device (bench input for the 4096-tree and 1M-tree configs).
// Device twin of spgnn_b200/synth.py: Philox4x32-10, counter (i0, i1, stream, tree), key (seed, "SPGN").
// The integer part (tree shape, labels) is bit-identical to the host generator; the normals agree to rounding.
#include "common.cuh"

namespace spgnn {

constexpr uint32_t kKey1 = 0x5350474Eu;

__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ int64_t mulhi_pick(uint32_t r, int64_t n) { return (int64_t)(((uint64_t)r * (uint64_t)n) >> 32); }

__device__ __forceinline__ int64_t tree_k(int64_t tree, uint32_t seed, int ragged, int64_t k_fixed) {
    if (!ragged) return k_fixed;
    const uint32_t r = philox4x32(0u, 0u, 1u, (uint32_t)tree, seed, kKey1).x;
    return 120 + mulhi_pick(r, 61);
}

__global__ void synth_sizes_kernel(int64_t first, int64_t B, uint32_t seed, int ragged, int64_t k_fixed,
                                   int64_t* __restrict__ n_nodes, int64_t* __restrict__ n_edges) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < B; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = 2 * tree_k(first + t, seed, ragged, k_fixed) + 1;
        n_nodes[t] = n;
        n_edges[t] = 3 * n - 2;   // 2(n-1) tree edges + n self loops
    }
}

// one thread per tree; the labels slice doubles as the leaf list while the tree grows
__global__ void synth_trees_kernel(int64_t first, int64_t B, uint32_t seed, int ragged, int64_t k_fixed,
                                   const int64_t* __restrict__ node_off, int64_t* __restrict__ parent,
                                   int64_t* __restrict__ labels) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < B; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tree = first + t, base = node_off[t];
        const int64_t n = node_off[t + 1] - base, k = (n - 1) / 2;
        int64_t* par = parent + base;
        int64_t* leaves = labels + base;
        par[0] = -1;
        leaves[0] = 0;
        for (int64_t j = 0; j < k; ++j) {
            const uint32_t r = philox4x32((uint32_t)j, 0u, 0u, (uint32_t)tree, seed, kKey1).x;
            const int64_t p = mulhi_pick(r, j + 1);
            const int64_t node = leaves[p];
            par[2 * j + 1] = node;
            par[2 * j + 2] = node;
            leaves[p] = 2 * j + 1;
            leaves[j + 1] = 2 * j + 2;
        }
        for (int64_t i = 0; i < n; ++i) leaves[i] = 0;
        // labels 1..21: partial Fisher-Yates with a sparse record of the touched positions
        int64_t pos[42], val[42];
        int cnt = 0;
        auto get = [&](int64_t i) { for (int q = cnt - 1; q >= 0; --q) if (pos[q] == i) return val[q]; return i; };
        auto put = [&](int64_t i, int64_t v) { for (int q = 0; q < cnt; ++q) if (pos[q] == i) { val[q] = v; return; } pos[cnt] = i; val[cnt] = v; ++cnt; };
        for (int j = 0; j < 21 && j < n; ++j) {
            const uint32_t r = philox4x32((uint32_t)j, 0u, 2u, (uint32_t)tree, seed, kKey1).x;
            const int64_t p = j + mulhi_pick(r, n - j);
            const int64_t a = get(j), b = get(p);
            put(j, b);
            put(p, a);
            labels[base + b] = j + 1;
        }
    }
}

// DGL-order edge list straight from the parent array: row r lists parent(r) < r, then its two children > r
__global__ void synth_edges_kernel(const int64_t* __restrict__ node_off, const int64_t* __restrict__ edge_off,
                                   const int64_t* __restrict__ parent, int64_t B, int64_t* __restrict__ sl,
                                   int64_t* __restrict__ dl) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < B; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t base = node_off[t], n = node_off[t + 1] - base, eb = edge_off[t];
        const int64_t* par = parent + base;
        int64_t* child0 = dl + eb + 2 * (n - 1);   // scratch in the self-loop tail, overwritten last
        for (int64_t i = 0; i < n; ++i) child0[i] = -1;
        for (int64_t c = 1; c < n; c += 2) child0[par[c]] = c;
        int64_t w = eb;
        for (int64_t r = 0; r < n; ++r) {
            if (r > 0) { sl[w] = r; dl[w] = par[r]; ++w; }
            const int64_t c = child0[r];
            if (c >= 0) { sl[w] = r; dl[w] = c; ++w; sl[w] = r; dl[w] = c + 1; ++w; }
        }
        for (int64_t r = 0; r < n; ++r) { sl[w + r] = r; dl[w + r] = r; }
    }
}

__device__ __forceinline__ float u32_to_unit(uint32_t x) { return ((float)x + 0.5f) * (1.0f / 4294967296.0f); }
__device__ __forceinline__ void normals4(uint4 w, float (&z)[4]) {
    const float r0 = sqrtf(-2.f * logf(u32_to_unit(w.x))), r1 = sqrtf(-2.f * logf(u32_to_unit(w.z)));
    float s0, c0, s1, c1;
    sincospif(2.f * u32_to_unit(w.y), &s0, &c0);
    sincospif(2.f * u32_to_unit(w.w), &s1, &c1);
    z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
}

__device__ __forceinline__ int64_t seg_of(const int64_t* __restrict__ off, int64_t nseg, int64_t x) {
    int64_t lo = 0, hi = nseg;
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (off[mid] <= x) lo = mid; else hi = mid; }
    return lo;
}

// stream 3: fvs = max(0, N(0,1)); one thread per 4 consecutive columns (fv_dim % 4 == 0)
__global__ void synth_fvs_kernel(int64_t first, int64_t B, uint32_t seed, const int64_t* __restrict__ node_off,
                                 float* __restrict__ fvs, int64_t ldf, int64_t fv_dim) {
    const int64_t N = node_off[B], q = fv_dim / 4, total = N * q;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / q, c4 = i - row * q;
        const int64_t t = seg_of(node_off, B, row);
        const int64_t blk = (row - node_off[t]) * q + c4;
        float z[4];
        normals4(philox4x32((uint32_t)blk, 0u, 3u, (uint32_t)(first + t), seed, kKey1), z);
        st4(fvs + row * ldf + c4 * 4, make_float4(fmaxf(z[0], 0.f), fmaxf(z[1], 0.f), fmaxf(z[2], 0.f), fmaxf(z[3], 0.f)));
    }
}

// stream 4: fvs_out = 3 N(0,1); one thread per node row (n_class values, blocks may straddle rows)
__global__ void synth_fvs_out_kernel(int64_t first, int64_t B, uint32_t seed, const int64_t* __restrict__ node_off,
                                     float* __restrict__ fo, int64_t ldo, int64_t C) {
    const int64_t N = node_off[B];
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < N; row += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = seg_of(node_off, B, row);
        const int64_t e0 = (row - node_off[t]) * C;
        int64_t cur_blk = -1;
        float z[4];
        for (int64_t c = 0; c < C; ++c) {
            const int64_t e = e0 + c, blk = e >> 2;
            if (blk != cur_blk) {
                normals4(philox4x32((uint32_t)blk, 0u, 4u, (uint32_t)(first + t), seed, kKey1), z);
                cur_blk = blk;
            }
            fo[row * ldo + c] = 3.f * z[e & 3];
        }
    }
}

static inline unsigned tgrid(int64_t n, int threads) {
    int64_t want = ceil_div(n, threads), cap = (int64_t)sm_count() * 16;
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

}  // namespace spgnn

using namespace spgnn;

extern "C" int spgnn_synth_sizes(int64_t first_tree, int64_t B, uint32_t seed, int ragged, int64_t k_fixed,
                                 int64_t* n_nodes, int64_t* n_edges, void* stream) {
    SPGNN_REQUIRE(n_nodes && n_edges && B > 0 && k_fixed >= 10, "synth_sizes: bad argument");
    synth_sizes_kernel<<<tgrid(B, 128), 128, 0, as_stream(stream)>>>(first_tree, B, seed, ragged, k_fixed, n_nodes, n_edges);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

extern "C" int spgnn_synth_trees(int64_t first_tree, int64_t B, uint32_t seed, int ragged, int64_t k_fixed,
                                 const int64_t* node_off, int64_t* parent_local, int64_t* labels, void* stream) {
    SPGNN_REQUIRE(node_off && parent_local && labels && B > 0, "synth_trees: bad argument");
    synth_trees_kernel<<<tgrid(B, 64), 64, 0, as_stream(stream)>>>(first_tree, B, seed, ragged, k_fixed, node_off,
                                                                  parent_local, labels);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

extern "C" int spgnn_synth_edges(const int64_t* node_off, const int64_t* edge_off, const int64_t* parent_local,
                                 int64_t B, int64_t* src_local, int64_t* dst_local, void* stream) {
    SPGNN_REQUIRE(node_off && edge_off && parent_local && src_local && dst_local && B > 0, "synth_edges: bad argument");
    synth_edges_kernel<<<tgrid(B, 64), 64, 0, as_stream(stream)>>>(node_off, edge_off, parent_local, B, src_local, dst_local);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

extern "C" int spgnn_synth_features(int64_t first_tree, int64_t B, uint32_t seed, const int64_t* node_off, int64_t N,
                                    float* fvs, int64_t ldf, int64_t fv_dim, float* fvs_out, int64_t ldo,
                                    int64_t n_class, void* stream) {
    SPGNN_REQUIRE(node_off && B > 0 && N > 0, "synth_features: bad argument");
    cudaStream_t st = as_stream(stream);
    if (fvs) {
        SPGNN_REQUIRE(fv_dim % 4 == 0 && ldf % 4 == 0 && ((uintptr_t)fvs & 15) == 0, "synth_features: fvs alignment");
        synth_fvs_kernel<<<tgrid(N * (fv_dim / 4), 256), 256, 0, st>>>(first_tree, B, seed, node_off, fvs, ldf, fv_dim);
        SPGNN_LAUNCH_OK();
    }
    if (fvs_out) {
        synth_fvs_out_kernel<<<tgrid(N, 128), 128, 0, st>>>(first_tree, B, seed, node_off, fvs_out, ldo, n_class);
        SPGNN_LAUNCH_OK();
    }
    return SPGNN_OK;
}

SPGNN_REGISTER_SALT(synth)
