// Device graph batch builder: prefix scans, dgl.batch id arithmetic, in-CSC / out-CSR, dense adj -> COO.
// Integer work only; every result is bit-identical to what DGL produces (see include/spgnn_b200.h).
//
// HBM-bound byte/integer kernels: coalesced int64/int32 streams, grids sized to a multiple of the SM count.
#include "common.cuh"

namespace spgnn {

// ------------------------------------------------------------------------------------------------
// 3-phase exclusive scan (block reduce -> recursive scan of block sums -> block scan + offset)
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;   // 1024

template <typename Tin>
__global__ void scan_block_sums(const Tin* __restrict__ in, int64_t n, int64_t* __restrict__ block_sums) {
    __shared__ int64_t warp_part[kScanThreads / 32];
    int64_t base = (int64_t)blockIdx.x * kScanTile;
    int64_t v = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        int64_t idx = base + (int64_t)i * kScanThreads + threadIdx.x;
        if (idx < n) v += (int64_t)in[idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t s = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) s += warp_part[w];
        block_sums[blockIdx.x] = s;
    }
}

// Each thread owns kScanItems CONSECUTIVE elements so the scan order is the memory order.
template <typename Tin, typename Tout>
__global__ void scan_blocks(const Tin* __restrict__ in, int64_t n, const int64_t* __restrict__ block_offs,
                            Tout* __restrict__ out) {
    __shared__ int64_t warp_part[kScanThreads / 32];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int64_t x[kScanItems];
    int64_t tsum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        x[i] = (base + i < n) ? (int64_t)in[base + i] : 0;
        tsum += x[i];
    }
    // inclusive warp scan of thread sums
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_part[warp] = inc;
    __syncthreads();
    int64_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += warp_part[w];
    int64_t excl = (block_offs ? block_offs[blockIdx.x] : 0) + woff + inc - tsum;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = (Tout)excl;
        excl += x[i];
        if (base + i == n - 1) out[n] = (Tout)excl;   // grand total
    }
}

__global__ void scan_write_zero(int64_t* out) { out[0] = 0; }

static int64_t scan_ws_elems(int64_t n) {
    // block sums + their scanned offsets, recursively
    int64_t total = 0;
    while (n > kScanTile) {
        int64_t nb = ceil_div(n, kScanTile);
        total += 2 * (nb + 1);
        n = nb;
    }
    return total + 4;
}

template <typename Tin, typename Tout>
static int scan_impl(const Tin* in, Tout* out, int64_t n, int64_t* ws, cudaStream_t st) {
    if (n <= 0) {
        SPGNN_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(Tout), st));
        return SPGNN_OK;
    }
    int64_t nb = ceil_div(n, kScanTile);
    int64_t* offs = nullptr;
    if (nb > 1) {
        int64_t* sums = ws;
        offs = ws + (nb + 1);
        scan_block_sums<Tin><<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, sums);
        SPGNN_LAUNCH_OK();
        int rc = scan_impl<int64_t, int64_t>(sums, offs, nb, ws + 2 * (nb + 1), st);
        if (rc) return rc;
    }
    scan_blocks<Tin, Tout><<<(unsigned)nb, kScanThreads, 0, st>>>(in, n, offs, out);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

// ------------------------------------------------------------------------------------------------
// dgl.batch
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t find_segment(const int64_t* __restrict__ off, int64_t nseg, int64_t x) {
    // largest g with off[g] <= x   (off has nseg+1 entries, off[0] = 0); skips empty segments correctly
    int64_t lo = 0, hi = nseg;
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (off[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void k_node_gid(const int64_t* __restrict__ node_off, int64_t B, int64_t N, int32_t* __restrict__ gid) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
        gid[i] = (int32_t)find_segment(node_off, B, i);
}

__global__ void k_globalize(const int64_t* __restrict__ node_off, const int64_t* __restrict__ edge_off, int64_t B,
                            int64_t E, const int64_t* __restrict__ sl, const int64_t* __restrict__ dl,
                            int64_t* __restrict__ src, int64_t* __restrict__ dst,
                            int32_t* __restrict__ deg_in, int32_t* __restrict__ deg_out, int32_t* __restrict__ flags) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t g = find_segment(edge_off, B, e);
        int64_t base = node_off[g], n = node_off[g + 1] - base;
        int64_t s = sl[e], d = dl[e];
        if (s < 0 || s >= n || d < 0 || d >= n) {
            atomicAdd(&flags[1], 1);
            s = d = 0;
        }
        s += base;
        d += base;
        src[e] = s;
        dst[e] = d;
        atomicAdd(&deg_in[d], 1);
        atomicAdd(&deg_out[s], 1);
    }
}

__global__ void k_fill_in(const int64_t* __restrict__ dst, int64_t E, const int32_t* __restrict__ in_ptr,
                          int32_t* __restrict__ cursor, int32_t* __restrict__ in_eid) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t d = dst[e];
        int32_t p = in_ptr[d] + atomicAdd(&cursor[d], 1);
        in_eid[p] = (int32_t)e;
    }
}

// One thread per node: insertion-sort the (short) segment; airway trees have degree <= 4.
__global__ void k_sort_in(const int64_t* __restrict__ src, const int32_t* __restrict__ in_ptr, int64_t N,
                          int32_t* __restrict__ in_eid, int32_t* __restrict__ in_src, int32_t* __restrict__ flags) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < N; v += (int64_t)gridDim.x * blockDim.x) {
        int32_t b = in_ptr[v], e = in_ptr[v + 1];
        if (e == b) atomicAdd(&flags[0], 1);
        if (e - b > *reinterpret_cast<volatile int32_t*>(flags + 2)) atomicMax(&flags[2], e - b);     // monotone: mostly skipped
        for (int32_t i = b + 1; i < e; ++i) {
            int32_t key = in_eid[i];
            int32_t j = i - 1;
            while (j >= b && in_eid[j] > key) { in_eid[j + 1] = in_eid[j]; --j; }
            in_eid[j + 1] = key;
        }
        for (int32_t i = b; i < e; ++i) in_src[i] = (int32_t)src[in_eid[i]];
    }
}

__global__ void k_fill_out(const int32_t* __restrict__ in_src, int64_t E, const int32_t* __restrict__ out_ptr,
                           int32_t* __restrict__ cursor, int32_t* __restrict__ out_slot) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < E; p += (int64_t)gridDim.x * blockDim.x) {
        int32_t s = in_src[p];
        int32_t q = out_ptr[s] + atomicAdd(&cursor[s], 1);
        out_slot[q] = (int32_t)p;
    }
}

__global__ void k_sort_out(const int64_t* __restrict__ dst, const int32_t* __restrict__ in_eid,
                           const int32_t* __restrict__ out_ptr, int64_t N, int32_t* __restrict__ out_slot,
                           int32_t* __restrict__ out_dst, int32_t* __restrict__ flags) {
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < N; u += (int64_t)gridDim.x * blockDim.x) {
        int32_t b = out_ptr[u], e = out_ptr[u + 1];
        if (e - b > *reinterpret_cast<volatile int32_t*>(flags + 3)) atomicMax(&flags[3], e - b);
        for (int32_t i = b + 1; i < e; ++i) {
            int32_t key = out_slot[i];
            int32_t j = i - 1;
            while (j >= b && out_slot[j] > key) { out_slot[j + 1] = out_slot[j]; --j; }
            out_slot[j + 1] = key;
        }
        for (int32_t i = b; i < e; ++i) out_dst[i] = (int32_t)dst[in_eid[out_slot[i]]];
    }
}

// ------------------------------------------------------------------------------------------------
// dense adjacency -> edge list (one warp per adjacency row)
// ------------------------------------------------------------------------------------------------
__global__ void k_adj_count(const uint8_t* __restrict__ adj, const int64_t* __restrict__ adj_off,
                            const int64_t* __restrict__ node_off, int64_t B, int64_t N,
                            int64_t* __restrict__ row_cnt, unsigned long long* __restrict__ n_edges) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < N; row += nwarp) {
        int64_t g = find_segment(node_off, B, row);
        int64_t r = row - node_off[g], n = node_off[g + 1] - node_off[g];
        const uint8_t* p = adj + adj_off[g] + r * n;
        int cnt = 0;
        for (int64_t c = lane; c < n; c += 32) cnt += (p[c] != 0 && c != r);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
        if (lane == 0) {
            row_cnt[row] = cnt;
            atomicAdd(&n_edges[g], (unsigned long long)cnt + 1ull);   // +1: this row's self loop
        }
    }
}

__global__ void k_adj_fill(const uint8_t* __restrict__ adj, const int64_t* __restrict__ adj_off,
                           const int64_t* __restrict__ node_off, const int64_t* __restrict__ edge_off,
                           const int64_t* __restrict__ row_off, int64_t B, int64_t N,
                           int64_t* __restrict__ sl, int64_t* __restrict__ dl) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < N; row += nwarp) {
        int64_t g = find_segment(node_off, B, row);
        int64_t nb = node_off[g], r = row - nb, n = node_off[g + 1] - nb;
        const uint8_t* p = adj + adj_off[g] + r * n;
        int64_t w = edge_off[g] + (row_off[row] - row_off[nb]);
        for (int64_t c0 = 0; c0 < n; c0 += 32) {
            int64_t c = c0 + lane;
            bool nz = (c < n) && p[c] != 0 && c != r;
            unsigned m = __ballot_sync(kFull, nz);
            if (nz) {
                int64_t pos = w + __popc(m & ((1u << lane) - 1u));
                sl[pos] = r;
                dl[pos] = c;
            }
            w += __popc(m);
        }
        if (lane == 0) {   // self loops go last: ids E_g - n_g + r  (g.add_edges(g.nodes(), g.nodes()))
            int64_t pos = edge_off[g + 1] - n + r;
            sl[pos] = r;
            dl[pos] = r;
        }
    }
}

static inline unsigned grid_for(int64_t work, int threads, int per_sm = 8) {
    int64_t want = ceil_div(work, threads);
    int64_t cap = (int64_t)sm_count() * per_sm;
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

}  // namespace spgnn

using namespace spgnn;

extern "C" int64_t spgnn_scan_ws_bytes(int64_t n) { return scan_ws_elems(n) * (int64_t)sizeof(int64_t); }

extern "C" int spgnn_scan_i64(const int64_t* in, int64_t* out, int64_t n, void* ws, void* stream) {
    SPGNN_REQUIRE(out && (n == 0 || in), "scan: null pointer");
    SPGNN_REQUIRE(n <= kScanTile || ws, "scan: workspace required for n > %d", kScanTile);
    return scan_impl<int64_t, int64_t>(in, out, n, (int64_t*)ws, as_stream(stream));
}

extern "C" int spgnn_adj_count(const uint8_t* adj, const int64_t* adj_off, const int64_t* n_nodes,
                               const int64_t* node_off, int64_t B, int64_t N, int64_t* row_cnt, int64_t* n_edges,
                               void* stream) {
    (void)n_nodes;
    SPGNN_REQUIRE(adj && adj_off && node_off && row_cnt && n_edges && B > 0 && N > 0, "adj_count: bad argument");
    cudaStream_t st = as_stream(stream);
    SPGNN_CUDA_OK(cudaMemsetAsync(n_edges, 0, sizeof(int64_t) * B, st));
    k_adj_count<<<grid_for(N * 32, 256), 256, 0, st>>>(adj, adj_off, node_off, B, N, row_cnt,
                                                        (unsigned long long*)n_edges);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

extern "C" int spgnn_adj_fill(const uint8_t* adj, const int64_t* adj_off, const int64_t* n_nodes,
                              const int64_t* node_off, const int64_t* edge_off, const int64_t* row_off, int64_t B,
                              int64_t N, int64_t* src_local, int64_t* dst_local, void* stream) {
    (void)n_nodes;
    SPGNN_REQUIRE(adj && adj_off && node_off && edge_off && row_off && src_local && dst_local, "adj_fill: null");
    k_adj_fill<<<grid_for(N * 32, 256), 256, 0, as_stream(stream)>>>(adj, adj_off, node_off, edge_off, row_off, B, N,
                                                                     src_local, dst_local);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

static int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

extern "C" int64_t spgnn_batch_ws_bytes(int64_t N, int64_t E) {
    (void)E;
    // deg_in, deg_out, cursor_in, cursor_out (int32 [N]) + scan scratch
    return 4 * align256((N + 1) * 4) + align256(spgnn_scan_ws_bytes(N + 1)) + 256;
}

extern "C" int spgnn_batch_build(const int64_t* node_off, const int64_t* edge_off, int64_t B, int64_t N, int64_t E,
                                 const int64_t* src_local, const int64_t* dst_local, int64_t* src, int64_t* dst,
                                 int32_t* node_gid, int32_t* in_ptr, int32_t* in_src, int32_t* in_eid,
                                 int32_t* out_ptr, int32_t* out_dst, int32_t* out_slot, int32_t* flags, void* ws,
                                 void* stream) {
    SPGNN_REQUIRE(node_off && edge_off && src_local && dst_local && src && dst && node_gid && in_ptr && in_src &&
                      in_eid && out_ptr && out_dst && out_slot && flags && ws,
                  "batch_build: null pointer");
    SPGNN_REQUIRE(B > 0 && N > 0 && E > 0, "batch_build: empty batch (B=%lld N=%lld E=%lld)", (long long)B,
                  (long long)N, (long long)E);
    SPGNN_REQUIRE(N < (1ll << 31) - 1 && E < (1ll << 31) - 1, "batch_build: N/E exceed int32 index range");
    cudaStream_t st = as_stream(stream);
    char* w = (char*)ws;
    int64_t seg = align256((N + 1) * 4);
    int32_t* deg_in = (int32_t*)w;
    int32_t* deg_out = (int32_t*)(w + seg);
    int32_t* cur_in = (int32_t*)(w + 2 * seg);
    int32_t* cur_out = (int32_t*)(w + 3 * seg);
    int64_t* scan_ws = (int64_t*)(w + 4 * seg);
    SPGNN_CUDA_OK(cudaMemsetAsync(w, 0, 4 * seg, st));
    SPGNN_CUDA_OK(cudaMemsetAsync(flags, 0, 4 * sizeof(int32_t), st));

    k_node_gid<<<grid_for(N, 256), 256, 0, st>>>(node_off, B, N, node_gid);
    SPGNN_LAUNCH_OK();
    k_globalize<<<grid_for(E, 256), 256, 0, st>>>(node_off, edge_off, B, E, src_local, dst_local, src, dst, deg_in,
                                                  deg_out, flags);
    SPGNN_LAUNCH_OK();
    int rc = scan_impl<int32_t, int32_t>(deg_in, in_ptr, N, scan_ws, st);
    if (rc) return rc;
    rc = scan_impl<int32_t, int32_t>(deg_out, out_ptr, N, scan_ws, st);
    if (rc) return rc;
    k_fill_in<<<grid_for(E, 256), 256, 0, st>>>(dst, E, in_ptr, cur_in, in_eid);
    SPGNN_LAUNCH_OK();
    k_sort_in<<<grid_for(N, 128), 128, 0, st>>>(src, in_ptr, N, in_eid, in_src, flags);
    SPGNN_LAUNCH_OK();
    k_fill_out<<<grid_for(E, 256), 256, 0, st>>>(in_src, E, out_ptr, cur_out, out_slot);
    SPGNN_LAUNCH_OK();
    k_sort_out<<<grid_for(N, 128), 128, 0, st>>>(dst, in_eid, out_ptr, N, out_slot, out_dst, flags);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}
