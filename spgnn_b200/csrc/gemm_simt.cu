// fp32 SIMT GEMM family for the dense projections (mode 0).  128x128x16 CTA tile, 8x8 register tile per
// thread, register-prefetched double-buffered shared memory, 128-bit global loads where alignment allows.
//
//   fwd        C[M,N]  = [A1|A2][M,K] * W[N,K]^T (+bias)(act)        A: k-contiguous, B: k-contiguous
//   bwd input  dA[M,K] = dC[M,N]     * W[N,K]                          A: k-contiguous, B: n-contiguous
//   bwd weight dW[N,K] = dC[M,N]^T   * A[M,K]    (split over M)        A: m-contiguous, B: n-contiguous
//
// Roofline: compute (FFMA) bound; the tensor-core path in gemm_tc.cu replaces it for the large layers.
#include "common.cuh"

namespace spgnn {

constexpr int BM = 128, BN = 128, BK = 16, GT = 256;
constexpr int PADM = 4;

struct GemmArgs {
    // operand A (rows = m, reduction = k)
    const float* A1; int64_t lda1; int64_t K1;
    const float* A2; int64_t lda2; int64_t K2;      // second source (k-contiguous A only)
    // operand B
    const float* B; int64_t ldb; int64_t b_koff;     // b_koff: column offset into B's k (kcontig) or n (ncontig) axis
    float* C; int64_t ldc;
    int64_t M, N;                                    // output extent
    int64_t Kred;                                    // total reduction length (K1+K2 tile-padded handled inside)
    const float* bias; int act; float slope;
    int64_t k_chunk;                                 // split-K chunk (multiple of BK); gridDim.z splits
    int64_t split_stride;                            // elements between split outputs
    int vecA1, vecA2, vecB, vecC;
};

template <bool KC>
__device__ __forceinline__ void load_tile(const float* __restrict__ p, int64_t ld, int64_t r0, int64_t R, int64_t k0,
                                          int64_t Klim, bool vec, float4 (&reg)[2]) {
    const int tid = threadIdx.x;
    if (KC) {
        const int64_t k = k0 + (tid & 3) * 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int64_t r = r0 + (tid >> 2) + i * 64;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < R) {
                const float* q = p + r * ld + k;
                if (vec && k + 3 < Klim) {
                    v = ldg4(q);
                } else {
                    if (k + 0 < Klim) v.x = __ldg(q + 0);
                    if (k + 1 < Klim) v.y = __ldg(q + 1);
                    if (k + 2 < Klim) v.z = __ldg(q + 2);
                    if (k + 3 < Klim) v.w = __ldg(q + 3);
                }
            }
            reg[i] = v;
        }
    } else {
        const int64_t r = r0 + (tid & 31) * 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int64_t k = k0 + (tid >> 5) + i * 8;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < Klim) {
                const float* q = p + k * ld + r;
                if (vec && r + 3 < R) {
                    v = ldg4(q);
                } else {
                    if (r + 0 < R) v.x = __ldg(q + 0);
                    if (r + 1 < R) v.y = __ldg(q + 1);
                    if (r + 2 < R) v.z = __ldg(q + 2);
                    if (r + 3 < R) v.w = __ldg(q + 3);
                }
            }
            reg[i] = v;
        }
    }
}

template <bool KC>
__device__ __forceinline__ void store_tile(float (*S)[BM + PADM], const float4 (&reg)[2]) {
    const int tid = threadIdx.x;
    if (KC) {
        const int kk = (tid & 3) * 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = (tid >> 2) + i * 64;
            S[kk + 0][r] = reg[i].x;
            S[kk + 1][r] = reg[i].y;
            S[kk + 2][r] = reg[i].z;
            S[kk + 3][r] = reg[i].w;
        }
    } else {
        const int r = (tid & 31) * 4;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int kk = (tid >> 5) + i * 8;
            *reinterpret_cast<float4*>(&S[kk][r]) = reg[i];
        }
    }
}

// AKC: A is k-contiguous (A[m*lda+k]) else m-contiguous (A[k*lda+m]); BKC likewise for B[n*ldb+k] / B[k*ldb+n].
template <bool AKC, bool BKC>
__global__ void __launch_bounds__(GT, 2) gemm_simt_kernel(const GemmArgs g) {
    __shared__ __align__(16) float As[2][BK][BM + PADM];
    __shared__ __align__(16) float Bs[2][BK][BN + PADM];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    // 1-D tile index, n-tiles fastest: CTAs that run together share the same A row band (L2 reuse of the big operand)
    const int64_t ntn = (g.N + BN - 1) / BN;
    const int64_t m0 = ((int64_t)blockIdx.x / ntn) * BM, n0 = ((int64_t)blockIdx.x % ntn) * BN;

    // reduction tiles: source 1 covers tiles [0,T1), source 2 tiles [T1, T1+T2); each padded to BK
    const int64_t T1 = (g.K1 + BK - 1) / BK;
    const int64_t T2 = AKC ? (g.K2 + BK - 1) / BK : 0;
    int64_t t_begin = 0, t_end = T1 + T2;
    if (g.k_chunk > 0) {
        t_begin = (int64_t)blockIdx.z * (g.k_chunk / BK);
        int64_t te = t_begin + g.k_chunk / BK;
        t_end = te < t_end ? te : t_end;
    }

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];
    auto fetch = [&](int64_t t) {
        const bool s2 = AKC && t >= T1;
        const int64_t k0 = s2 ? (t - T1) * BK : t * BK;
        const int64_t Klim = s2 ? g.K2 : g.K1;
        const float* Ap = s2 ? g.A2 : g.A1;
        const int64_t lda = s2 ? g.lda2 : g.lda1;
        const bool va = s2 ? g.vecA2 : g.vecA1;
        load_tile<AKC>(Ap, lda, m0, g.M, k0, Klim, va, ra);
        // B's reduction index continues after K1 for the second source
        const int64_t kb0 = s2 ? g.K1 + k0 : k0;
        const int64_t Kblim = s2 ? g.K1 + g.K2 : g.K1;
        if (BKC) {
            load_tile<true>(g.B + g.b_koff, g.ldb, n0, g.N, kb0, Kblim, g.vecB && ((kb0 & 3) == 0), rb);
        } else {
            load_tile<false>(g.B + g.b_koff, g.ldb, n0, g.N, kb0, Kblim, g.vecB, rb);
        }
    };

    if (t_begin < t_end) {
        fetch(t_begin);
        store_tile<AKC>(As[0], ra);
        store_tile<BKC>(Bs[0], rb);
    }
    __syncthreads();

    int buf = 0;
    for (int64_t t = t_begin; t < t_end; ++t) {
        const bool more = t + 1 < t_end;
        if (more) fetch(t + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            store_tile<AKC>(As[buf ^ 1], ra);
            store_tile<BKC>(Bs[buf ^ 1], rb);
        }
        __syncthreads();
        buf ^= 1;
    }

    float* C = g.C + (g.k_chunk > 0 ? (int64_t)blockIdx.z * g.split_stride : 0);
    const bool epi = g.k_chunk == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int64_t n = n0 + jh * 64 + tx * 4;
            if (n >= g.N) continue;
            float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
            if (epi) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n + j < g.N) {
                        if (g.bias) v[j] += __ldg(g.bias + n + j);
                        v[j] = act_fwd(v[j], g.act, g.slope);
                    }
                }
            }
            float* q = C + m * g.ldc + n;
            if (g.vecC && n + 3 < g.N) {
                st4(q, make_float4(v[0], v[1], v[2], v[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < g.N) q[j] = v[j];
            }
        }
    }
}

__global__ void reduce_splits_kernel(const float* __restrict__ ws, int64_t splits, int64_t split_stride, int64_t N,
                                     int64_t K, float* __restrict__ out, int64_t ldo) {
    const int64_t total = N * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // fixed order: eight interleaved running sums (independent loads in flight), combined pairwise at the end
        float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int64_t z = 0;
        for (; z + 8 <= splits; z += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) s[u] += __ldg(ws + (z + u) * split_stride + i);
        }
        for (int u = 0; z < splits; ++z, ++u) s[u] += __ldg(ws + z * split_stride + i);
        out[(i / K) * ldo + (i % K)] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
    }
}

// column sums: partial per block over a row band, then a fixed-order final reduce
__global__ void colsum_partial_kernel(const float* __restrict__ X, int64_t ldx, int64_t M, int64_t N,
                                      int64_t rows_per_block, float* __restrict__ part) {
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < N; c += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int64_t r = r0; r < r1; ++r) s += __ldg(X + r * ldx + c);
        part[(int64_t)blockIdx.y * N + c] = s;
    }
}
// one warp per column: lane l adds the partial rows l, l + 32, ... then a butterfly (a fixed order, hence reproducible);
// one thread per column walked 592 dependent loads: 50 us for the 22 columns of the head's bias gradient
__global__ void colsum_final_kernel(const float* __restrict__ part, int64_t nparts, int64_t N, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp; c < N; c += nwarps) {
        float s = 0.f;
        for (int64_t p = lane; p < nparts; p += 32) s += part[p * N + c];
        s = warp_sum(s);
        if (lane == 0) out[c] = s;
    }
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

constexpr int64_t kColsumParts = 592;   // 4 x 148 row bands

int simt_linear_fwd(const float* A1, int64_t lda1, int64_t K1, const float* A2, int64_t lda2, int64_t K2,
                    const float* W, int64_t ldw, const float* bias, int act, float slope, float* C, int64_t ldc,
                    int64_t M, int64_t N, cudaStream_t st) {
    GemmArgs g{};
    g.A1 = A1; g.lda1 = lda1; g.K1 = K1; g.A2 = A2; g.lda2 = lda2; g.K2 = A2 ? K2 : 0;
    g.B = W; g.ldb = ldw; g.b_koff = 0; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.Kred = K1 + K2;
    g.bias = bias; g.act = act; g.slope = slope; g.k_chunk = 0; g.split_stride = 0;
    g.vecA1 = aligned16(A1) && (lda1 % 4 == 0);
    g.vecA2 = A2 && aligned16(A2) && (lda2 % 4 == 0);
    g.vecB = aligned16(W) && (ldw % 4 == 0);
    g.vecC = aligned16(C) && (ldc % 4 == 0);
    dim3 grid((unsigned)(ceil_div(N, BN) * ceil_div(M, BM)), 1, 1);
    gemm_simt_kernel<true, true><<<grid, GT, 0, st>>>(g);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

int simt_linear_bwd_input(const float* dC, int64_t lddc, const float* W, int64_t ldw, int64_t k_off, float* dA,
                          int64_t ldda, int64_t M, int64_t N, int64_t K, cudaStream_t st) {
    // dA[m,k] = sum_n dC[m,n] * W[n, k_off+k]: A = dC (k-contig over n), B[kred=n][col=k] n-contiguous in k
    GemmArgs g{};
    g.A1 = dC; g.lda1 = lddc; g.K1 = N; g.A2 = nullptr; g.K2 = 0;
    g.B = W; g.ldb = ldw; g.b_koff = k_off; g.C = dA; g.ldc = ldda; g.M = M; g.N = K; g.Kred = N;
    g.bias = nullptr; g.act = 0; g.slope = 0; g.k_chunk = 0;
    g.vecA1 = aligned16(dC) && (lddc % 4 == 0);
    g.vecB = aligned16(W + k_off) && (ldw % 4 == 0);
    g.vecC = aligned16(dA) && (ldda % 4 == 0);
    dim3 grid((unsigned)(ceil_div(K, BN) * ceil_div(M, BM)), 1, 1);
    gemm_simt_kernel<true, false><<<grid, GT, 0, st>>>(g);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

static void weight_split_plan(int64_t M, int64_t N, int64_t K, int64_t* splits, int64_t* k_chunk) {
    int64_t tiles = ceil_div(N, BM) * ceil_div(K, BN);
    int64_t want = ceil_div((int64_t)sm_count() * 4, tiles);
    int64_t max_by_rows = ceil_div(M, 1024);          // at least 1024 reduction rows per split
    if (want > max_by_rows) want = max_by_rows;
    if (want > 128) want = 128;
    if (want < 1) want = 1;
    int64_t chunk = ceil_div(ceil_div(M, want), BK) * BK;
    *k_chunk = chunk;
    *splits = ceil_div(M, chunk);
}

int64_t simt_linear_bwd_weight_ws(int64_t M, int64_t N, int64_t K) {
    int64_t splits, chunk;
    weight_split_plan(M, N, K, &splits, &chunk);
    return splits * N * K * (int64_t)sizeof(float);
}

int simt_linear_bwd_weight(const float* dC, int64_t lddc, const float* A, int64_t lda, float* dW, int64_t lddw,
                           int64_t k_off, int64_t M, int64_t N, int64_t K, void* ws, cudaStream_t st) {
    // dW[n,k] = sum_m dC[m,n] * A[m,k]: "A operand" = dC read m-contiguous in n (rows = n), "B operand" = A (cols = k)
    int64_t splits, chunk;
    weight_split_plan(M, N, K, &splits, &chunk);
    GemmArgs g{};
    g.A1 = dC; g.lda1 = lddc; g.K1 = M; g.A2 = nullptr; g.K2 = 0;
    g.B = A; g.ldb = lda; g.b_koff = 0; g.M = N; g.N = K; g.Kred = M;
    g.bias = nullptr; g.act = 0; g.slope = 0;
    g.vecA1 = aligned16(dC) && (lddc % 4 == 0);
    g.vecB = aligned16(A) && (lda % 4 == 0);
    if (splits == 1) {
        g.C = dW + k_off; g.ldc = lddw; g.k_chunk = 0; g.split_stride = 0;
        g.vecC = aligned16(dW + k_off) && (lddw % 4 == 0);
    } else {
        g.C = (float*)ws; g.ldc = K; g.k_chunk = chunk; g.split_stride = N * K;
        g.vecC = aligned16(ws) && (K % 4 == 0) && ((N * K) % 4 == 0);
    }
    dim3 grid((unsigned)(ceil_div(K, BN) * ceil_div(N, BM)), 1, (unsigned)splits);
    gemm_simt_kernel<false, false><<<grid, GT, 0, st>>>(g);
    SPGNN_LAUNCH_OK();
    if (splits > 1) {
        int64_t total = N * K;
        unsigned blocks = (unsigned)(ceil_div(total, 256) < (int64_t)sm_count() * 8 ? ceil_div(total, 256)
                                                                                    : (int64_t)sm_count() * 8);
        reduce_splits_kernel<<<blocks, 256, 0, st>>>((const float*)ws, splits, N * K, N, K, dW + k_off, lddw);
        SPGNN_LAUNCH_OK();
    }
    return SPGNN_OK;
}

void reduce_splits(const float* ws, int64_t splits, int64_t N, int64_t K, float* out, int64_t ldo, cudaStream_t st) {
    const int64_t total = N * K;
    const int64_t want = ceil_div(total, 256), cap = (int64_t)sm_count() * 8;
    reduce_splits_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(ws, splits, N * K, N, K, out, ldo);
    count_launch();
}

}  // namespace spgnn

using namespace spgnn;

extern "C" int64_t spgnn_colsum_ws(int64_t N) { return kColsumParts * N * (int64_t)sizeof(float); }

extern "C" int spgnn_colsum(const float* X, int64_t ldx, int64_t M, int64_t N, float* out, void* ws, void* stream) {
    SPGNN_REQUIRE(X && out && ws && M > 0 && N > 0, "colsum: bad argument");
    cudaStream_t st = as_stream(stream);
    int64_t parts = M < kColsumParts ? M : kColsumParts;
    int64_t rpb = ceil_div(M, parts);
    parts = ceil_div(M, rpb);
    dim3 grid((unsigned)ceil_div(N, 128), (unsigned)parts);
    colsum_partial_kernel<<<grid, 128, 0, st>>>(X, ldx, M, N, rpb, (float*)ws);
    SPGNN_LAUNCH_OK();
    const int64_t fblocks = ceil_div(N * 32, 128), fcap = (int64_t)sm_count() * 8;
    colsum_final_kernel<<<(unsigned)(fblocks < fcap ? fblocks : fcap), 128, 0, st>>>((const float*)ws, parts, N, out);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}
