// C-ABI entry points of the dense projection family; dispatches between the fp32 SIMT kernels
// (gemm_simt.cu) and the tcgen05 tensor-core kernels (gemm_tc.cu).
#include "common.cuh"

namespace spgnn {
int simt_linear_fwd(const float*, int64_t, int64_t, const float*, int64_t, int64_t, const float*, int64_t,
                    const float*, int, float, float*, int64_t, int64_t, int64_t, cudaStream_t);
int simt_linear_bwd_input(const float*, int64_t, const float*, int64_t, int64_t, float*, int64_t, int64_t, int64_t,
                          int64_t, cudaStream_t);
int64_t simt_linear_bwd_weight_ws(int64_t, int64_t, int64_t);
int simt_linear_bwd_weight(const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int64_t,
                           int64_t, void*, cudaStream_t);
int64_t tc_linear_ws_bytes(int64_t N, int64_t K1, int64_t K2);
int tc_linear_fwd(const float*, int64_t, int64_t, const float*, int64_t, int64_t, const float*, int64_t, const float*,
                  int, float, float*, int64_t, int64_t, int64_t, void*, cudaStream_t);
int64_t tc_linear_bwd_input_ws_bytes(int64_t N, int64_t K);
int tc_linear_bwd_input(const float*, int64_t, const float*, int64_t, int64_t, float*, int64_t, int64_t, int64_t,
                        int64_t, void*, cudaStream_t);
int64_t tc_linear_bwd_weight_ws_bytes(int64_t, int64_t, int64_t);
int tc_linear_bwd_weight(const float*, int64_t, const float*, int64_t, float*, int64_t, int64_t, int64_t, int64_t,
                         int64_t, void*, cudaStream_t);

int64_t tc_linear_bwd_weight2_ws_bytes(int64_t M, int64_t N, int64_t K1, int64_t K2);
int tc_linear_bwd_weight2(const float*, int64_t, const float*, int64_t, int64_t, const float*, int64_t, int64_t, float*,
                          int64_t, int64_t, int64_t, void*, cudaStream_t);

static inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }
}  // namespace spgnn
using namespace spgnn;

extern "C" int spgnn_linear_fwd(const float* A1, int64_t lda1, int64_t K1, const float* A2, int64_t lda2, int64_t K2,
                                const float* W, int64_t ldw, const float* bias, int act, float slope, float* C,
                                int64_t ldc, int64_t M, int64_t N, int mode, void* ws, int64_t ws_bytes, void* stream) {
    SPGNN_REQUIRE(A1 && W && C, "linear_fwd: null pointer");
    SPGNN_REQUIRE(M > 0 && N > 0 && K1 > 0 && K2 >= 0, "linear_fwd: bad shape M=%lld N=%lld K1=%lld K2=%lld",
                  (long long)M, (long long)N, (long long)K1, (long long)K2);
    SPGNN_REQUIRE(lda1 >= K1 && ldw >= K1 + (A2 ? K2 : 0) && ldc >= N && (!A2 || lda2 >= K2),
                  "linear_fwd: leading dimension smaller than row length");
    const bool tc_ok = al16(A1) && lda1 % 4 == 0 && (!A2 || (al16(A2) && lda2 % 4 == 0)) && al16(C) && ldc % 4 == 0 &&
                       K1 + K2 < (1 << 30) && N < (1 << 30);
    if (mode == 1 && tc_ok) {
        SPGNN_REQUIRE(ws && ws_bytes >= tc_linear_ws_bytes(N, K1, A2 ? K2 : 0), "linear_fwd: workspace too small");
        return tc_linear_fwd(A1, lda1, K1, A2, lda2, K2, W, ldw, bias, act, slope, C, ldc, M, N, ws, as_stream(stream));
    }
    return simt_linear_fwd(A1, lda1, K1, A2, lda2, K2, W, ldw, bias, act, slope, C, ldc, M, N, as_stream(stream));
}

extern "C" int spgnn_linear_bwd_input(const float* dC, int64_t lddc, const float* W, int64_t ldw, int64_t k_off,
                                      float* dA, int64_t ldda, int64_t M, int64_t N, int64_t K, int mode,
                                      void* ws, int64_t ws_bytes, void* stream) {
    SPGNN_REQUIRE(dC && W && dA && M > 0 && N > 0 && K > 0, "linear_bwd_input: bad argument");
    SPGNN_REQUIRE(lddc >= N && ldw >= k_off + K && ldda >= K, "linear_bwd_input: leading dimension too small");
    if (mode == 1 && al16(dC) && lddc % 4 == 0 && al16(dA) && ldda % 4 == 0) {
        SPGNN_REQUIRE(ws && ws_bytes >= tc_linear_bwd_input_ws_bytes(N, K), "linear_bwd_input: workspace too small");
        return tc_linear_bwd_input(dC, lddc, W, ldw, k_off, dA, ldda, M, N, K, ws, as_stream(stream));
    }
    return simt_linear_bwd_input(dC, lddc, W, ldw, k_off, dA, ldda, M, N, K, as_stream(stream));
}

extern "C" int64_t spgnn_linear_bwd_weight_ws(int64_t M, int64_t N, int64_t K) {
    const int64_t a = simt_linear_bwd_weight_ws(M, N, K), b = tc_linear_bwd_weight_ws_bytes(M, N, K);
    return a > b ? a : b;
}
extern "C" int64_t spgnn_linear_fwd_ws(int64_t N, int64_t K1, int64_t K2) { return tc_linear_ws_bytes(N, K1, K2); }
extern "C" int64_t spgnn_linear_bwd_input_ws(int64_t N, int64_t K) { return tc_linear_bwd_input_ws_bytes(N, K); }

extern "C" int spgnn_linear_bwd_weight(const float* dC, int64_t lddc, const float* A, int64_t lda, float* dW,
                                       int64_t lddw, int64_t k_off, int64_t M, int64_t N, int64_t K, void* ws,
                                       int mode, void* stream) {
    SPGNN_REQUIRE(dC && A && dW && ws && M > 0 && N > 0 && K > 0, "linear_bwd_weight: bad argument");
    SPGNN_REQUIRE(lddc >= N && lda >= K && lddw >= k_off + K, "linear_bwd_weight: leading dimension too small");
    if (mode == 1 && al16(dC) && lddc % 4 == 0 && al16(A) && lda % 4 == 0)
        return tc_linear_bwd_weight(dC, lddc, A, lda, dW, lddw, k_off, M, N, K, ws, as_stream(stream));
    return simt_linear_bwd_weight(dC, lddc, A, lda, dW, lddw, k_off, M, N, K, ws, as_stream(stream));
}

extern "C" int64_t spgnn_linear_bwd_weight2_ws(int64_t M, int64_t N, int64_t K1, int64_t K2) {
    int64_t m = simt_linear_bwd_weight_ws(M, N, K1);
    const int64_t cand[4] = {K2 > 0 ? simt_linear_bwd_weight_ws(M, N, K2) : 0, tc_linear_bwd_weight2_ws_bytes(M, N, K1, K2),
                             tc_linear_bwd_weight_ws_bytes(M, N, K1), K2 > 0 ? tc_linear_bwd_weight_ws_bytes(M, N, K2) : 0};
    for (int i = 0; i < 4; ++i) m = cand[i] > m ? cand[i] : m;
    return m;
}

extern "C" int spgnn_linear_bwd_weight2(const float* dC, int64_t lddc, const float* A1, int64_t lda1, int64_t K1,
                                        const float* A2, int64_t lda2, int64_t K2, float* dW, int64_t lddw, int64_t M,
                                        int64_t N, void* ws, int mode, void* stream) {
    SPGNN_REQUIRE(dC && A1 && dW && ws && M > 0 && N > 0 && K1 > 0 && K2 >= 0, "linear_bwd_weight2: bad argument");
    SPGNN_REQUIRE(lddc >= N && lda1 >= K1 && (!A2 || lda2 >= K2) && lddw >= K1 + (A2 ? K2 : 0),
                  "linear_bwd_weight2: leading dimension too small");
    const bool ok = al16(dC) && lddc % 4 == 0 && al16(A1) && lda1 % 4 == 0 && (!A2 || (al16(A2) && lda2 % 4 == 0));
    // mode 2: experimental 256 x BN two-accumulator kernel (one pass over dC for both sources).  Measured on B200 it
    // is latency-bound and slower than two passes of the 128x128 kernel (26 vs 19 ms on the 1063->1028 layer), so
    // mode 1 runs the 128x128 kernel once per source.
    if (mode == 2 && ok)
        return tc_linear_bwd_weight2(dC, lddc, A1, lda1, K1, A2, lda2, K2, dW, lddw, M, N, ws, as_stream(stream));
    if (mode == 1 && ok) {
        int rc1 = tc_linear_bwd_weight(dC, lddc, A1, lda1, dW, lddw, 0, M, N, K1, ws, as_stream(stream));
        if (rc1 || !A2 || K2 == 0) return rc1;
        return tc_linear_bwd_weight(dC, lddc, A2, lda2, dW, lddw, K1, M, N, K2, ws, as_stream(stream));
    }
    int rc = simt_linear_bwd_weight(dC, lddc, A1, lda1, dW, lddw, 0, M, N, K1, ws, as_stream(stream));
    if (rc || !A2 || K2 == 0) return rc;
    return simt_linear_bwd_weight(dC, lddc, A2, lda2, dW, lddw, K1, M, N, K2, ws, as_stream(stream));
}
