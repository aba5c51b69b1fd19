// Wire format of a scan batch between the host loader and the device batch builder (the H2D boundary the reference
// crosses at job_runner.py:1872-1875 with one dense float tensor and one dense [n, n] adjacency per scan).
//
// At 4096 trees per step that copy is 5.5 GB and the end-to-end step is PCIe-bound (104 ms of copy vs 80 ms of
// compute on one GPU; 8 GPUs share the host path and fall to 22 GB/s each).  Both big tensors are mostly zeros for
// structural reasons: the CNN embedding `fvs` is the output of a ReLU (models.py:1100-1107) and the adjacency of a
// tree has 3n - 2 non-zeros out of n^2.  The loader therefore ships them LOSSLESSLY compacted:
//
//   * zero-suppressed rows (fvs): mask [rows, ceil(cols/32)] uint32 — bit b of word w set iff the fp32 BIT PATTERN of
//     column 32w + b is non-zero (so -0.0 survives) — the non-zero values in row-major order, and row_off [rows + 1];
//   * edge lists (adj): the off-diagonal non-zeros of every scan's adjacency in row-major order (= DGL edge order,
//     SURVEY.md 8a A7) as int32 local (src, dst) pairs + the per-scan counts; the self loops DGL appends last are
//     generated on the device.
//
// Host side (spgnn_host_*: plain C++ threads, no CUDA calls) packs a batch once when it is read; the device side
// decodes into exactly the tensors the dense path builds (bit-identical: tests/test_gpu_runners.py).
#include "common.cuh"
#include <thread>
#include <vector>
#include <string.h>

namespace spgnn {
namespace wire {

static inline int clamp_threads(int threads, int64_t rows) {
    if (threads < 1) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if ((int64_t)threads > rows) threads = (int)(rows > 0 ? rows : 1);
    return threads;
}

template <typename F>
static void parallel_rows(int64_t rows, int threads, F fn) {
    threads = clamp_threads(threads, rows);
    if (threads == 1) { fn(0, rows); return; }
    std::vector<std::thread> pool;
    const int64_t per = (rows + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        const int64_t r0 = t * per, r1 = r0 + per < rows ? r0 + per : rows;
        if (r0 >= r1) break;
        pool.emplace_back([=] { fn(r0, r1); });
    }
    for (auto& th : pool) th.join();
}

// One warp per row.  Lane l first owns mask word l of a group of 32 words (one coalesced load, one warp scan of the
// pop-counts); the group is then decoded 4 words = 128 columns per step: lane l takes the 4-bit nibble n = l % 8 of word
// 4 * step + l / 8, i.e. 4 consecutive columns, so a warp instruction stores 512 contiguous bytes (float4 per lane)
// and the value loads of neighbouring lanes touch consecutive addresses.  (The first version decoded one word per
// step with one column per lane: 2 shuffles and a 128-byte store per 32 columns, 3.5 TB/s.)
__global__ void __launch_bounds__(256) unpack_rows_kernel(const uint32_t* __restrict__ mask, const float* __restrict__ vals,
                                                          const int64_t* __restrict__ row_off, int64_t rows, int cols,
                                                          int words, float* __restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int q = lane >> 3, sh = (lane & 7) * 4;
    const uint32_t below = (1u << sh) - 1u;
    const bool vec_ok = (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const float* v = vals + __ldg(row_off + r);
        const uint32_t* m = mask + r * words;
        float* o = out + r * ldo;
        int base = 0;
        for (int w0 = 0; w0 < words; w0 += 32) {
            const uint32_t mine = (w0 + lane < words) ? __ldg(m + w0 + lane) : 0u;
            int incl = __popc(mine);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += t;
            }
            const int excl = incl - __popc(mine);
            const int nw = min(32, words - w0);
#pragma unroll 2
            for (int j0 = 0; j0 < nw; j0 += 4) {
                const int j = j0 + q;                              // this lane's word of the step (< 32)
                const uint32_t wj = __shfl_sync(kFull, mine, j & 31);
                const int pj = __shfl_sync(kFull, excl, j & 31);
                const uint32_t w = j < nw ? wj : 0u;               // no value loads behind the last word of the row
                const uint32_t bits = (w >> sh) & 0xFu;
                int off = base + pj + __popc(w & below);
                float4 x;
                x.x = (bits & 1u) ? __ldg(v + off) : 0.f; off += bits & 1u;
                x.y = (bits & 2u) ? __ldg(v + off) : 0.f; off += (bits >> 1) & 1u;
                x.z = (bits & 4u) ? __ldg(v + off) : 0.f; off += (bits >> 2) & 1u;
                x.w = (bits & 8u) ? __ldg(v + off) : 0.f;
                const int col = (w0 + j) * 32 + sh;
                if (j < nw) {
                    if (vec_ok && col + 3 < cols) {
                        *reinterpret_cast<float4*>(o + col) = x;
                    } else {
                        if (col < cols) o[col] = x.x;
                        if (col + 1 < cols) o[col + 1] = x.y;
                        if (col + 2 < cols) o[col + 2] = x.z;
                        if (col + 3 < cols) o[col + 3] = x.w;
                    }
                }
            }
            base += __shfl_sync(kFull, incl, 31);
        }
    }
}

// compact edge lists -> the per-graph LOCAL edge lists of DGL (int64), self loops appended last
__global__ void edges_expand_kernel(const int32_t* __restrict__ src32, const int32_t* __restrict__ dst32,
                                    const int64_t* __restrict__ ne_off, const int64_t* __restrict__ node_off, int64_t B,
                                    int64_t NE, int64_t N, int64_t* __restrict__ sl, int64_t* __restrict__ dl,
                                    int64_t* __restrict__ n_edges) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < NE + N; i += nth) {
        if (i < NE) {
            int64_t lo = 0, hi = B;                         // largest g with ne_off[g] <= i
            while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (ne_off[mid] <= i) lo = mid; else hi = mid; }
            const int64_t pos = i + node_off[lo];
            sl[pos] = src32[i];
            dl[pos] = dst32[i];
        } else {
            const int64_t v = i - NE;
            int64_t lo = 0, hi = B;
            while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (node_off[mid] <= v) lo = mid; else hi = mid; }
            const int64_t r = v - node_off[lo];
            const int64_t pos = ne_off[lo + 1] + node_off[lo] + r;
            sl[pos] = r;
            dl[pos] = r;
        }
    }
    for (int64_t g = tid; g < B; g += nth) n_edges[g] = (ne_off[g + 1] - ne_off[g]) + (node_off[g + 1] - node_off[g]);
}

}  // namespace wire
}  // namespace spgnn

using namespace spgnn;
using namespace spgnn::wire;

extern "C" int64_t spgnn_host_pack_rows_count(const float* x, int64_t ldx, int64_t rows, int64_t cols, int64_t* row_off,
                                              int threads) {
    if (!x || !row_off || rows < 0 || cols <= 0 || ldx < cols) { set_error("host_pack_rows_count: bad argument"); return -1; }
    parallel_rows(rows, threads, [=](int64_t r0, int64_t r1) {
        for (int64_t r = r0; r < r1; ++r) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(x + r * ldx);
            int64_t c = 0;
            for (int64_t k = 0; k < cols; ++k) c += p[k] != 0u;
            row_off[r + 1] = c;
        }
    });
    row_off[0] = 0;
    for (int64_t r = 0; r < rows; ++r) row_off[r + 1] += row_off[r];
    return row_off[rows];
}

extern "C" int spgnn_host_pack_rows_fill(const float* x, int64_t ldx, int64_t rows, int64_t cols, const int64_t* row_off,
                                         uint32_t* mask, float* vals, int threads) {
    SPGNN_REQUIRE(x && row_off && mask && (vals || row_off[rows] == 0) && rows >= 0 && cols > 0 && ldx >= cols,
                  "host_pack_rows_fill: bad argument");
    const int64_t words = (cols + 31) / 32;
    parallel_rows(rows, threads, [=](int64_t r0, int64_t r1) {
        for (int64_t r = r0; r < r1; ++r) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(x + r * ldx);
            uint32_t* v = reinterpret_cast<uint32_t*>(vals) + row_off[r];
            uint32_t* m = mask + r * words;
            for (int64_t w = 0; w < words; ++w) {
                const int64_t k0 = w * 32, k1 = k0 + 32 < cols ? k0 + 32 : cols;
                uint32_t bits = 0u;
                for (int64_t k = k0; k < k1; ++k) {
                    const uint32_t b = p[k];
                    if (b != 0u) *v++ = b;                // (an unconditional store would touch the next row's slot)
                    bits |= (uint32_t)(b != 0u) << (k - k0);
                }
                m[w] = bits;
            }
        }
    });
    return SPGNN_OK;
}

// dense adjacency blocks (uint8 [n_g, n_g] concatenated) -> int32 edge lists of the off-diagonal non-zeros, row-major
extern "C" int64_t spgnn_host_adj_edges_count(const uint8_t* adj_cat, const int64_t* n_nodes, int64_t B, int64_t* ne_off,
                                              int threads) {
    if (!adj_cat || !n_nodes || !ne_off || B <= 0) { set_error("host_adj_edges_count: bad argument"); return -1; }
    std::vector<int64_t> aoff(B + 1, 0);
    for (int64_t g = 0; g < B; ++g) aoff[g + 1] = aoff[g] + n_nodes[g] * n_nodes[g];
    const int64_t* ao = aoff.data();
    parallel_rows(B, threads, [=](int64_t g0, int64_t g1) {
        for (int64_t g = g0; g < g1; ++g) {
            const int64_t n = n_nodes[g];
            const uint8_t* a = adj_cat + ao[g];
            int64_t c = 0;
            for (int64_t i = 0; i < n * n; ++i) c += a[i] != 0;
            for (int64_t r = 0; r < n; ++r) c -= a[r * n + r] != 0;
            ne_off[g + 1] = c;
        }
    });
    ne_off[0] = 0;
    for (int64_t g = 0; g < B; ++g) ne_off[g + 1] += ne_off[g];
    return ne_off[B];
}

extern "C" int spgnn_host_adj_edges_fill(const uint8_t* adj_cat, const int64_t* n_nodes, int64_t B, const int64_t* ne_off,
                                         int32_t* src, int32_t* dst, int32_t* max_degree, int threads) {
    SPGNN_REQUIRE(adj_cat && n_nodes && ne_off && B > 0 && (ne_off[B] == 0 || (src && dst)), "host_adj_edges_fill: bad argument");
    std::vector<int64_t> aoff(B + 1, 0);
    for (int64_t g = 0; g < B; ++g) aoff[g + 1] = aoff[g] + n_nodes[g] * n_nodes[g];
    const int64_t* ao = aoff.data();
    std::vector<int32_t> gmax(B, 0);
    int32_t* gm = gmax.data();
    parallel_rows(B, threads, [=](int64_t g0, int64_t g1) {
        std::vector<int32_t> indeg;
        for (int64_t g = g0; g < g1; ++g) {
            const int64_t n = n_nodes[g];
            const uint8_t* a = adj_cat + ao[g];
            int64_t o = ne_off[g];
            indeg.assign((size_t)n, 0);
            int32_t best = 0;
            for (int64_t r = 0; r < n; ++r) {
                int32_t outdeg = 0;
                const uint8_t* row = a + r * n;
                for (int64_t c = 0; c < n; ++c)
                    if (row[c] != 0 && c != r) {
                        src[o] = (int32_t)r;
                        dst[o] = (int32_t)c;
                        ++o;
                        ++outdeg;
                        ++indeg[(size_t)c];
                    }
                best = outdeg > best ? outdeg : best;
            }
            for (int64_t c = 0; c < n; ++c) best = indeg[(size_t)c] > best ? indeg[(size_t)c] : best;
            gm[g] = best;
        }
    });
    if (max_degree) {
        int32_t m = 0;
        for (int64_t g = 0; g < B; ++g) m = gmax[g] > m ? gmax[g] : m;
        *max_degree = m + 1;                          // + the self loop appended to every node on the device
    }
    return SPGNN_OK;
}

extern "C" int spgnn_unpack_rows(const uint32_t* mask, const float* vals, const int64_t* row_off, int64_t rows, int64_t cols,
                                 float* out, int64_t ldo, void* stream) {
    SPGNN_REQUIRE(mask && row_off && out && rows > 0 && cols > 0 && ldo >= cols, "unpack_rows: bad argument");
    const int64_t warps = rows, blocks = ceil_div(warps * 32, 256), cap = (int64_t)sm_count() * 16;
    unpack_rows_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(
        mask, vals, row_off, rows, (int)cols, (int)((cols + 31) / 32), out, ldo);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

extern "C" int spgnn_edges_expand(const int32_t* src32, const int32_t* dst32, const int64_t* ne_off, const int64_t* node_off,
                                  int64_t B, int64_t NE, int64_t N, int64_t* src_local, int64_t* dst_local,
                                  int64_t* n_edges, void* stream) {
    SPGNN_REQUIRE(ne_off && node_off && src_local && dst_local && n_edges && B > 0 && N > 0 && NE >= 0 &&
                      (NE == 0 || (src32 && dst32)),
                  "edges_expand: bad argument");
    const int64_t blocks = ceil_div(NE + N, 256), cap = (int64_t)sm_count() * 8;
    edges_expand_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(
        src32, dst32, ne_off, node_off, B, NE, N, src_local, dst_local, n_edges);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}
