// Positional-encoding inits on device (one CTA per graph):
//   anchor_select   job_runner.py:1727-1757 (+ add_distal_leafs :1712-1725)
//   pe_dist_init    job_runner.py:1759-1777   all-pairs bit-parallel BFS -> diameter, hops to the anchors
//   pe_rw_init      job_runner.py:1684-1702   diag((A D^-1)^k), k = 1..pos_dim, fp64 accumulate
// Integer BFS + one IEEE fp32 divide: pos_enc is bit-identical to the reference's float32(hops / float(diameter)).
#include "common.cuh"

namespace spgnn {

constexpr int kPeThreads = 256;

struct ArgMax {
    float v;
    int i;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {   // first max: larger value, then smaller index
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}
__device__ __forceinline__ ArgMax block_argmax(ArgMax x, ArgMax* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax y;
        y.v = __shfl_xor_sync(kFull, x.v, o);
        y.i = __shfl_xor_sync(kFull, x.i, o);
        x = better(x, y);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = x;
    __syncthreads();
    ArgMax r = sh[0];
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) r = better(r, sh[k]);
    return r;
}

// dynamic smem: float mx[n], float inv_sum[n], uint8 mask[n], uint8 leaf[n], int16 dist[18][n]
__global__ void __launch_bounds__(kPeThreads) anchor_select_kernel(const float* __restrict__ fo, int64_t ld, int C,
                                                                   const int64_t* __restrict__ node_off,
                                                                   const int32_t* __restrict__ in_ptr,
                                                                   const int32_t* __restrict__ in_src, int pos_dim,
                                                                   int max_nodes, int32_t* __restrict__ anchors) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ ArgMax red[kPeThreads / 32];
    __shared__ int s_anchor[64];
    const int64_t base = node_off[blockIdx.x];
    const int n = (int)(node_off[blockIdx.x + 1] - base);
    float* mx = reinterpret_cast<float*>(smem);
    float* inv = mx + max_nodes;
    unsigned char* mask = reinterpret_cast<unsigned char*>(inv + max_nodes);
    unsigned char* leaf = mask + max_nodes;
    int16_t* dist = reinterpret_cast<int16_t*>(leaf + max_nodes);   // byte offset 10*max_nodes: even
    const int n_labels = C - 1;   // 21

    for (int v = threadIdx.x; v < n; v += blockDim.x) {
        const float* row = fo + (base + v) * ld;
        float m = -INFINITY;
        for (int c = 0; c < C; ++c) m = fmaxf(m, __ldg(row + c));
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += expf(__ldg(row + c) - m);
        mx[v] = m;
        inv[v] = s;
        mask[v] = 1;
        // leaf of the DAG {u->v : u<v}: no neighbour with a larger index
        bool has_child = false;
        for (int e = in_ptr[base + v]; e < in_ptr[base + v + 1]; ++e) has_child |= (in_src[e] - base) > v;
        leaf[v] = has_child ? 0 : 1;
    }
    __syncthreads();
    for (int label = 1; label <= n_labels; ++label) {
        ArgMax best{-INFINITY, 0x7fffffff};
        for (int v = threadIdx.x; v < n; v += blockDim.x) {
            // np.argmax(P[:, label] * mask): masked rows contribute exactly 0.0
            const float p = mask[v] ? expf(__ldg(fo + (base + v) * ld + label) - mx[v]) / inv[v] : 0.f;
            best = better(best, ArgMax{p, v});
        }
        best = block_argmax(best, red);
        if (threadIdx.x == 0) {
            mask[best.i] = 0;
            s_anchor[label - 1] = best.i;
            anchors[(int64_t)blockIdx.x * pos_dim + (label - 1)] = best.i;
        }
        __syncthreads();
    }
    if (pos_dim <= n_labels) return;
    // distal leaves of the first n_labels-3 anchors: one thread per anchor, single pass in index order
    const int n_extra = pos_dim - n_labels;   // 18
    if ((int)threadIdx.x < n_extra) {
        const int a = s_anchor[threadIdx.x];
        int16_t* d = dist + (int64_t)threadIdx.x * max_nodes;
        int best = a, best_d = -1;
        d[a] = 0;
        for (int v = a + 1; v < n; ++v) {
            int dv = 0x7fff;
            for (int e = in_ptr[base + v]; e < in_ptr[base + v + 1]; ++e) {
                const int u = (int)(in_src[e] - base);
                if (u >= a && u < v && d[u] != 0x7fff && d[u] + 1 < dv) dv = d[u] + 1;
            }
            d[v] = (int16_t)dv;
            if (dv != 0x7fff && leaf[v] && dv >= best_d) { best_d = dv; best = v; }   // ties -> largest index
        }
        anchors[(int64_t)blockIdx.x * pos_dim + n_labels + threadIdx.x] = best;
    }
}

// Fast path for TREES (what airway graphs are: dataset.py:409-417): the encoding needs hop counts from the pos_dim
// anchors only, and the diameter of a tree is the eccentricity of the node farthest from any start node (two BFS).
// One 64-bit reach mask per node — bit k: anchor k's wave has arrived, bit pos_dim: the wave of node 0 — instead
// of the n x n bitsets of the general kernel below: ~5x less shared-memory work and 10 KB instead of 24 KB per CTA.
// A graph is taken here iff it is symmetric, has exactly 2 (n - 1) non-self edges and node 0 reaches every node;
// everything else (cycles, directed or disconnected graphs) is left to the all-pairs kernel (done[g] = 0).
__global__ void __launch_bounds__(kPeThreads) pe_dist_tree_kernel(const int64_t* __restrict__ node_off,
                                                                  const int32_t* __restrict__ ptr,
                                                                  const int32_t* __restrict__ nbr,
                                                                  const int32_t* __restrict__ anchors, int pos_dim,
                                                                  float* __restrict__ pe, int64_t ldp,
                                                                  int32_t* __restrict__ diam, int32_t* __restrict__ done) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_flag, s_edges, s_far, s_lvl_far;
    __shared__ int s_anchor[64];
    const int64_t base = node_off[blockIdx.x];
    const int n = (int)(node_off[blockIdx.x + 1] - base);
    unsigned long long* cur = reinterpret_cast<unsigned long long*>(smem);
    unsigned long long* nxt = cur + n;
    int* d2 = reinterpret_cast<int*>(nxt + n);              // second BFS: hops from the far node
    const int tid = threadIdx.x, nt = blockDim.x;

    // ---- a symmetric graph with 2 (n - 1) non-self edges?
    if (tid == 0) { s_flag = 0; s_edges = 0; s_far = 0; }
    if (tid < pos_dim) s_anchor[tid] = anchors[(int64_t)blockIdx.x * pos_dim + tid];
    __syncthreads();
    int nonself = 0, asym = 0;
    for (int v = tid; v < n; v += nt) {
        for (int e = ptr[base + v]; e < ptr[base + v + 1]; ++e) {
            const int u = nbr[e] - (int)base;
            if (u == v) continue;
            ++nonself;
            bool back = false;
            for (int f = ptr[base + u]; f < ptr[base + u + 1]; ++f) back |= (nbr[f] - (int)base) == v;
            asym |= !back;
        }
    }
    if (nonself) atomicAdd(&s_edges, nonself);
    if (asym) s_flag = 1;
    __syncthreads();
    if (s_flag != 0 || s_edges != 2 * (n - 1)) {             // uniform: shared values after the barrier
        if (tid == 0) done[blockIdx.x] = 0;
        return;
    }
    // ---- BFS 1: waves of the anchors (bits 0..pos_dim-1) and of node 0 (bit pos_dim), level-synchronous
    const unsigned long long root_bit = 1ull << pos_dim;
    for (int v = tid; v < n; v += nt) {
        unsigned long long m = v == 0 ? root_bit : 0ull;
        for (int k = 0; k < pos_dim; ++k) m |= (s_anchor[k] == v) ? (1ull << k) : 0ull;
        cur[v] = m;
        for (int k = 0; k < pos_dim; ++k) pe[(base + v) * ldp + k] = ((m >> k) & 1ull) ? 0.f : -1.f;
    }
    __syncthreads();
    for (int level = 1;; ++level) {
        if (tid == 0) { s_flag = 0; s_lvl_far = -1; }
        __syncthreads();
        for (int v = tid; v < n; v += nt) {
            const unsigned long long old = cur[v];
            unsigned long long m = old;
            for (int e = ptr[base + v]; e < ptr[base + v + 1]; ++e) m |= cur[nbr[e] - (int)base];
            nxt[v] = m;
            unsigned long long fresh = m & ~old;
            if (fresh) {
                s_flag = 1;
                if (fresh & root_bit) {
                    atomicMax(&s_lvl_far, v);                 // the deepest level reached from node 0 wins (below)
                    fresh &= ~root_bit;
                }
                while (fresh) {
                    const int k = __ffsll((long long)fresh) - 1;
                    fresh &= fresh - 1;
                    pe[(base + v) * ldp + k] = (float)level;
                }
            }
        }
        __syncthreads();
        if (!s_flag) break;                                   // uniform
        if (tid == 0 && s_lvl_far >= 0) s_far = s_lvl_far;
        unsigned long long* t = cur; cur = nxt; nxt = t;
        __syncthreads();
    }
    // node 0 must have reached every node (a forest plus a cycle has the edge count of a tree)
    __syncthreads();                                          // every thread has read s_flag and left the loop
    if (tid == 0) s_flag = 0;
    __syncthreads();
    int unreached = 0;
    for (int v = tid; v < n; v += nt) unreached |= !(cur[v] & root_bit);
    if (unreached) s_flag = 1;
    __syncthreads();
    if (s_flag) {
        if (tid == 0) done[blockIdx.x] = 0;
        return;
    }
    // ---- BFS 2 from the node farthest from node 0: its eccentricity is the diameter of the tree
    const int far = s_far;
    int* mark = reinterpret_cast<int*>(nxt);                  // BFS 1 is over: its spare wave buffer holds the join marks
    for (int v = tid; v < n; v += nt) { d2[v] = v == far ? 0 : -1; mark[v] = 0; }
    __syncthreads();
    int ecc = 0;
    while (true) {
        if (tid == 0) s_flag = 0;
        __syncthreads();
        for (int v = tid; v < n; v += nt) {
            if (d2[v] != -1) continue;
            bool hit = false;
            for (int e = ptr[base + v]; e < ptr[base + v + 1]; ++e) hit |= d2[nbr[e] - (int)base] == ecc;
            if (hit) { mark[v] = 1; s_flag = 1; }             // joins level ecc + 1 after the barrier (d2 is only read here)
        }
        __syncthreads();
        if (!s_flag) break;                                   // uniform
        ++ecc;
        for (int v = tid; v < n; v += nt)
            if (mark[v]) { d2[v] = ecc; mark[v] = 0; }
        __syncthreads();
    }
    if (tid == 0) { diam[blockIdx.x] = ecc; done[blockIdx.x] = 1; }
    const float fd = (float)ecc;
    for (int i = tid; i < n * pos_dim; i += nt) {
        const int v = i / pos_dim, k = i - v * pos_dim;
        float* q = pe + (base + v) * ldp + k;
        *q = ecc > 0 ? __fdiv_rn(*q, fd) : 0.f;
    }
}

// reach[2][n][W] bitsets in shared memory (or in `ws` when too large)
__global__ void __launch_bounds__(kPeThreads) pe_dist_kernel(const int64_t* __restrict__ node_off,
                                                             const int32_t* __restrict__ in_ptr,
                                                             const int32_t* __restrict__ in_src,
                                                             const int32_t* __restrict__ anchors, int pos_dim,
                                                             int max_nodes, float* __restrict__ pe, int64_t ldp,
                                                             int32_t* __restrict__ diam, int32_t* __restrict__ flags,
                                                             uint32_t* __restrict__ ws, const int32_t* __restrict__ done) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_changed;
    __shared__ int s_anchor[64];
    if (done && done[blockIdx.x]) return;                     // a tree: pe_dist_tree_kernel has written it
    const int64_t base = node_off[blockIdx.x];
    const int n = (int)(node_off[blockIdx.x + 1] - base);
    const int W = (max_nodes + 31) >> 5;
    uint32_t* reach = ws ? ws + (int64_t)blockIdx.x * 2 * max_nodes * W : reinterpret_cast<uint32_t*>(smem);
    uint32_t* cur = reach;
    uint32_t* nxt = reach + (int64_t)max_nodes * W;
    const int Wn = (n + 31) >> 5;

    if ((int)threadIdx.x < pos_dim) s_anchor[threadIdx.x] = anchors[(int64_t)blockIdx.x * pos_dim + threadIdx.x];
    for (int i = threadIdx.x; i < n * Wn; i += blockDim.x) {
        const int v = i / Wn, w = i - v * Wn;
        cur[v * W + w] = (v >> 5) == w ? (1u << (v & 31)) : 0u;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * pos_dim; i += blockDim.x) {
        const int v = i / pos_dim, k = i - v * pos_dim;
        pe[(base + v) * ldp + k] = s_anchor[k] == v ? 0.f : -1.f;   // -1: not reached (disconnected graph)
    }
    int level = 0;
    while (true) {
        if (threadIdx.x == 0) s_changed = 0;
        __syncthreads();
        int changed = 0;
        for (int i = threadIdx.x; i < n * Wn; i += blockDim.x) {
            const int v = i / Wn, w = i - v * Wn;
            uint32_t r = cur[v * W + w];
            const uint32_t old = r;
            for (int e = in_ptr[base + v]; e < in_ptr[base + v + 1]; ++e) r |= cur[(in_src[e] - base) * W + w];
            nxt[v * W + w] = r;
            changed |= (r != old);
        }
        if (changed) s_changed = 1;
        __syncthreads();
        if (!s_changed) break;
        ++level;
        // hops to the anchors: the level at which anchor a's bit first shows up at v  (stored as a float count)
        for (int i = threadIdx.x; i < n * pos_dim; i += blockDim.x) {
            const int v = i / pos_dim, k = i - v * pos_dim;
            const int a = s_anchor[k];
            const uint32_t bit = 1u << (a & 31);
            if ((nxt[v * W + (a >> 5)] & bit) && !(cur[v * W + (a >> 5)] & bit)) pe[(base + v) * ldp + k] = (float)level;
        }
        __syncthreads();
        uint32_t* t = cur; cur = nxt; nxt = t;
    }
    __syncthreads();   // everyone has left the loop before s_changed is reused
    // connected <=> every row saturated
    int bad = 0;
    for (int i = threadIdx.x; i < n * Wn; i += blockDim.x) {
        const int v = i / Wn, w = i - v * Wn;
        const uint32_t full = (w == Wn - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
        bad |= (cur[v * W + w] != full);
    }
    if (bad) s_changed = 2;
    __syncthreads();
    if (threadIdx.x == 0) {
        diam[blockIdx.x] = level;
        if (s_changed == 2) atomicAdd(&flags[0], 1);
    }
    const float fd = (float)level;
    for (int i = threadIdx.x; i < n * pos_dim; i += blockDim.x) {
        const int v = i / pos_dim, k = i - v * pos_dim;
        float* q = pe + (base + v) * ldp + k;
        *q = level > 0 ? __fdiv_rn(*q, fd) : 0.f;
    }
}

// dynamic smem: double dinv[max_nodes]; per warp 2 x double[max_nodes]
__global__ void __launch_bounds__(kPeThreads) pe_rw_kernel(const int64_t* __restrict__ node_off,
                                                           const int32_t* __restrict__ in_ptr,
                                                           const int32_t* __restrict__ in_src, int pos_dim,
                                                           int max_nodes, float* __restrict__ pe, int64_t ldp) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t base = node_off[blockIdx.x];
    const int n = (int)(node_off[blockIdx.x + 1] - base);
    double* dinv = reinterpret_cast<double*>(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* xa = dinv + max_nodes + (int64_t)warp * 2 * max_nodes;
    double* xb = xa + max_nodes;
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
        int deg = 0;
        for (int e = in_ptr[base + v]; e < in_ptr[base + v + 1]; ++e) deg += (in_src[e] - base) != v;
        dinv[v] = 1.0 / (double)(deg < 1 ? 1 : deg);
    }
    __syncthreads();
    for (int src = warp; src < n; src += nw) {
        for (int r = lane; r < n; r += 32) xa[r] = r == src ? 1.0 : 0.0;
        __syncwarp();
        double* x = xa;
        double* y = xb;
        for (int k = 0; k < pos_dim; ++k) {
            for (int r = lane; r < n; r += 32) {
                double acc = 0.0;
                for (int e = in_ptr[base + r]; e < in_ptr[base + r + 1]; ++e) {
                    const int c = (int)(in_src[e] - base);
                    if (c != r) acc += x[c] * dinv[c];
                }
                y[r] = acc;
            }
            __syncwarp();
            if (lane == 0) pe[(base + src) * ldp + k] = (float)y[src];
            double* t = x; x = y; y = t;
        }
        __syncwarp();
    }
}

// per graph, per class c in [1,C): node with the highest softmax probability (first max)
__global__ void __launch_bounds__(kPeThreads) segmented_argmax_kernel(const float* __restrict__ logits, int64_t ld,
                                                                      int C, const int64_t* __restrict__ node_off,
                                                                      int64_t* __restrict__ out) {
    __shared__ ArgMax red[kPeThreads / 32];
    const int64_t base = node_off[blockIdx.x];
    const int n = (int)(node_off[blockIdx.x + 1] - base);
    for (int c = 1; c < C; ++c) {
        ArgMax best{-INFINITY, 0x7fffffff};
        for (int v = threadIdx.x; v < n; v += blockDim.x) {
            const float* row = logits + (base + v) * ld;
            float m = -INFINITY;
            for (int j = 0; j < C; ++j) m = fmaxf(m, __ldg(row + j));
            float s = 0.f;
            for (int j = 0; j < C; ++j) s += expf(__ldg(row + j) - m);
            best = better(best, ArgMax{expf(__ldg(row + c) - m) / s, v});
        }
        best = block_argmax(best, red);
        if (threadIdx.x == 0) out[(int64_t)blockIdx.x * (C - 1) + (c - 1)] = base + best.i;
        __syncthreads();
    }
}

}  // namespace spgnn

using namespace spgnn;

extern "C" int spgnn_anchor_select(const float* fvs_out, int64_t ld, int64_t n_class, const int64_t* node_off,
                                   const int32_t* in_ptr, const int32_t* in_src, int64_t B, int64_t pos_dim,
                                   int64_t max_nodes, int32_t* anchors, void* stream) {
    SPGNN_REQUIRE(fvs_out && node_off && in_ptr && in_src && anchors && B > 0, "anchor_select: bad argument");
    SPGNN_REQUIRE(n_class >= 2 && n_class - 1 <= 61, "anchor_select: n_class out of range");
    SPGNN_REQUIRE(pos_dim == n_class - 1 || pos_dim == 2 * (n_class - 1) - 3,
                  "pos enc dim : %lld! (21 or 39 for 22 classes, job_runner.py:1746-1755)", (long long)pos_dim);
    SPGNN_REQUIRE(max_nodes >= n_class - 1 && max_nodes < 32767, "anchor_select: max_nodes %lld out of range",
                  (long long)max_nodes);
    const int n_extra = (int)(pos_dim - (n_class - 1));
    size_t smem = (size_t)max_nodes * 8 + 2 * (size_t)max_nodes + 2 + (size_t)n_extra * max_nodes * 2;
    if (smem > 200 * 1024) {
        set_error("anchor_select: graph of %lld nodes needs %zu B shared memory", (long long)max_nodes, smem);
        return SPGNN_E_UNSUPPORTED;
    }
    SPGNN_CUDA_OK(cudaFuncSetAttribute(anchor_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    anchor_select_kernel<<<(unsigned)B, kPeThreads, smem, as_stream(stream)>>>(fvs_out, ld, (int)n_class, node_off,
                                                                               in_ptr, in_src, (int)pos_dim,
                                                                               (int)max_nodes, anchors);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

static const size_t kPeDistSmemMax = 160 * 1024;
static size_t pe_dist_bits_bytes(int64_t max_nodes) { return (size_t)2 * max_nodes * ((max_nodes + 31) / 32) * 4; }

static int64_t pe_done_bytes(int64_t B) { return (B * 4 + 255) & ~(int64_t)255; }

extern "C" int64_t spgnn_pe_dist_ws_bytes(int64_t B, int64_t max_nodes) {
    size_t per = pe_dist_bits_bytes(max_nodes);
    return pe_done_bytes(B) + (per <= kPeDistSmemMax ? 0 : (int64_t)(per * B));     // done flags [+ bitsets]
}

extern "C" int spgnn_pe_dist_init(const int64_t* node_off, const int32_t* in_ptr, const int32_t* in_src,
                                  const int32_t* anchors, int64_t B, int64_t pos_dim, int64_t max_nodes,
                                  float* pos_enc, int64_t ldp, int32_t* diam, int32_t* flags, void* ws, void* stream) {
    SPGNN_REQUIRE(node_off && in_ptr && in_src && anchors && pos_enc && diam && flags && B > 0, "pe_dist: bad argument");
    SPGNN_REQUIRE(pos_dim > 0 && pos_dim <= 64 && ldp >= pos_dim && max_nodes > 0, "pe_dist: bad shape");
    size_t per = pe_dist_bits_bytes(max_nodes);
    const bool use_ws = per > kPeDistSmemMax;
    SPGNN_REQUIRE(ws, "pe_dist: workspace of spgnn_pe_dist_ws_bytes(B, max_nodes) bytes required");
    cudaStream_t st = as_stream(stream);
    SPGNN_CUDA_OK(cudaMemsetAsync(flags, 0, sizeof(int32_t), st));
    int32_t* done = reinterpret_cast<int32_t*>(ws);
    uint32_t* bits = use_ws ? reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ws) + pe_done_bytes(B)) : nullptr;
    // trees (every airway graph) take the anchor-wave kernel; whatever it declines goes to the all-pairs kernel
    const bool fast = pos_dim < 63 && getenv("SPGNN_PE_ALL_PAIRS") == nullptr;
    if (fast) {
        const size_t tsmem = (size_t)max_nodes * (8 + 8 + 4) + 16;
        SPGNN_REQUIRE(tsmem <= 200 * 1024, "pe_dist: graph of %lld nodes is too large", (long long)max_nodes);
        static DeviceOnce tattr;
        if (tsmem > 48 * 1024 && tattr.pending()) {
            SPGNN_CUDA_OK(cudaFuncSetAttribute(pe_dist_tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            tattr.done();
        }
        pe_dist_tree_kernel<<<(unsigned)B, kPeThreads, tsmem, st>>>(node_off, in_ptr, in_src, anchors, (int)pos_dim,
                                                                    pos_enc, ldp, diam, done);
        SPGNN_LAUNCH_OK();
    }
    size_t smem = use_ws ? 0 : per;
    static DeviceOnce attr;
    if (attr.pending()) {
        SPGNN_CUDA_OK(cudaFuncSetAttribute(pe_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)kPeDistSmemMax));
        attr.done();
    }
    pe_dist_kernel<<<(unsigned)B, kPeThreads, smem, st>>>(node_off, in_ptr, in_src, anchors, (int)pos_dim,
                                                          (int)max_nodes, pos_enc, ldp, diam, flags, bits,
                                                          fast ? done : nullptr);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

extern "C" int spgnn_pe_rw_init(const int64_t* node_off, const int32_t* in_ptr, const int32_t* in_src, int64_t B,
                                int64_t pos_dim, int64_t max_nodes, float* pos_enc, int64_t ldp, void* stream) {
    SPGNN_REQUIRE(node_off && in_ptr && in_src && pos_enc && B > 0 && pos_dim > 0 && ldp >= pos_dim && max_nodes > 0,
                  "pe_rw: bad argument");
    int threads = kPeThreads;
    size_t smem = (size_t)max_nodes * 8 * (1 + 2 * (threads / 32));
    while (smem > 200 * 1024 && threads > 32) {
        threads >>= 1;
        smem = (size_t)max_nodes * 8 * (1 + 2 * (threads / 32));
    }
    if (smem > 200 * 1024) {
        set_error("pe_rw: graph of %lld nodes needs %zu B shared memory", (long long)max_nodes, smem);
        return SPGNN_E_UNSUPPORTED;
    }
    SPGNN_CUDA_OK(cudaFuncSetAttribute(pe_rw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pe_rw_kernel<<<(unsigned)B, threads, smem, as_stream(stream)>>>(node_off, in_ptr, in_src, (int)pos_dim,
                                                                    (int)max_nodes, pos_enc, ldp);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

extern "C" int spgnn_segmented_argmax(const float* logits, int64_t ld, int64_t n_class, const int64_t* node_off,
                                      int64_t B, int64_t* out, void* stream) {
    SPGNN_REQUIRE(logits && node_off && out && B > 0 && n_class >= 2 && ld >= n_class, "segmented_argmax: bad argument");
    segmented_argmax_kernel<<<(unsigned)B, kPeThreads, 0, as_stream(stream)>>>(logits, ld, (int)n_class, node_off, out);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}
