// GATConv weight packing (spgnn_gat_pack_weight, include/spgnn_b200.h) and its gradient.
//
// One projection of a GATConv yields z, the residual projection and both attention logits when its weight is
//   P = [ W_fc ; W_res ; W_l ; W_r ],   W_l[h, k] = sum_f W_fc[h*F + f, k] * attn_l[h*F + f]   (W_r likewise),
// because el = (x W_fc^T) . attn_l = x . (W_fc^T attn_l)  (DGL GATConv).  Building P with torch ops
// costs ~10 tiny kernels per layer forward and ~15 backward — most of the launches of a 64-tree step; here it is
// one kernel each way.  Parameters are a few MB: latency-bound, no roofline to speak of.
#include "common.cuh"

namespace spgnn {
namespace {

constexpr int kPackThreads = 256;

// blocks [0, n_red): one (head, 32-column tile) each — 8 warps stride over f, lanes over k, shared-memory reduce;
// blocks [n_red, grid): grid-stride copy of the W_fc / W_res rows and zero fill of the row padding.
__global__ void __launch_bounds__(kPackThreads) pack_weight_kernel(const float* __restrict__ Wfc, int64_t ldw,
                                                                   const float* __restrict__ Wres, int64_t ldr,
                                                                   const float* __restrict__ al,
                                                                   const float* __restrict__ ar, int H, int F, int K,
                                                                   float* __restrict__ out, int64_t ldo, int n_red) {
    __shared__ float sl[8][33], sr[8][33];
    const int HF = H * F;
    const int rows_w = HF * (Wres ? 2 : 1);
    if ((int)blockIdx.x < n_red) {
        const int ktiles = (K + 31) / 32;
        const int h = blockIdx.x / ktiles, k = (blockIdx.x - h * ktiles) * 32 + (threadIdx.x & 31);
        const int w = threadIdx.x >> 5;
        float a = 0.f, b = 0.f;
        if (k < K) {
#pragma unroll 4
            for (int f = w; f < F; f += 8) {
                const float x = __ldg(Wfc + (int64_t)(h * F + f) * ldw + k);
                a = fmaf(x, __ldg(al + h * F + f), a);
                b = fmaf(x, __ldg(ar + h * F + f), b);
            }
        }
        sl[w][threadIdx.x & 31] = a;
        sr[w][threadIdx.x & 31] = b;
        __syncthreads();
        if (w == 0 && k < K) {
            float ta = 0.f, tb = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) { ta += sl[q][threadIdx.x]; tb += sr[q][threadIdx.x]; }
            out[(int64_t)(rows_w + h) * ldo + k] = ta;
            out[(int64_t)(rows_w + H + h) * ldo + k] = tb;
        }
        return;
    }
    const int64_t total = (int64_t)(rows_w + 2 * H) * ldo;
    const int64_t nthreads = (int64_t)(gridDim.x - n_red) * blockDim.x;
    for (int64_t i = (int64_t)(blockIdx.x - n_red) * blockDim.x + threadIdx.x; i < total; i += nthreads) {
        const int64_t r = i / ldo;
        const int k = (int)(i - r * ldo);
        if (r >= rows_w) {
            if (k >= K) out[i] = 0.f;                       // padding of the el / er rows
            continue;
        }
        float v = 0.f;
        if (k < K) v = r < HF ? __ldg(Wfc + r * ldw + k) : __ldg(Wres + (r - HF) * ldr + k);
        out[i] = v;
    }
}

// one warp per row of W_fc (and of W_res): dW_fc[r, k] = dP[r, k] + al[r] dPl[h, k] + ar[r] dPr[h, k];
// d al[r] = sum_k W_fc[r, k] dPl[h, k], d ar[r] likewise; dW_res[r] = dP[HF + r].
__global__ void __launch_bounds__(kPackThreads) pack_weight_bwd_kernel(const float* __restrict__ dP, int64_t ldp,
                                                                       const float* __restrict__ Wfc, int64_t ldw,
                                                                       const float* __restrict__ al,
                                                                       const float* __restrict__ ar, int H, int F, int K,
                                                                       int has_res, float* __restrict__ dW, int64_t lddw,
                                                                       float* __restrict__ dR, int64_t lddr,
                                                                       float* __restrict__ dal, float* __restrict__ dar) {
    const int HF = H * F;
    const int rows_w = HF * (has_res ? 2 : 1);
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    if (row >= rows_w) return;
    if (row >= HF) {
        if (dR)
            for (int k = lane; k < K; k += 32) dR[(int64_t)(row - HF) * lddr + k] = dP[(int64_t)row * ldp + k];
        return;
    }
    const int h = row / F;
    const float* pl = dP + (int64_t)(rows_w + h) * ldp;
    const float* pr = dP + (int64_t)(rows_w + H + h) * ldp;
    const float a = __ldg(al + row), b = __ldg(ar + row);
    float sa = 0.f, sb = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float l = __ldg(pl + k), r = __ldg(pr + k);
        const float w = __ldg(Wfc + (int64_t)row * ldw + k);
        sa = fmaf(w, l, sa);
        sb = fmaf(w, r, sb);
        if (dW) dW[(int64_t)row * lddw + k] = dP[(int64_t)row * ldp + k] + a * l + b * r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xFFFFFFFFu, sa, o);
        sb += __shfl_xor_sync(0xFFFFFFFFu, sb, o);
    }
    if (lane == 0) {
        if (dal) dal[row] = sa;
        if (dar) dar[row] = sb;
    }
}

}  // namespace
}  // namespace spgnn

using namespace spgnn;

extern "C" int spgnn_gat_pack_weight(const float* W_fc, int64_t ldw, const float* W_res, int64_t ldr,
                                     const float* attn_l, const float* attn_r, int64_t H, int64_t F, int64_t K,
                                     float* out, int64_t ldo, void* stream) {
    SPGNN_REQUIRE(W_fc && attn_l && attn_r && out && H > 0 && F > 0 && K > 0, "gat_pack_weight: bad argument");
    SPGNN_REQUIRE(ldw >= K && ldo >= K && (!W_res || ldr >= K), "gat_pack_weight: leading dimension smaller than K");
    const int n_red = (int)(H * ceil_div(K, 32));
    const int64_t total = (H * F * (W_res ? 2 : 1) + 2 * H) * ldo;
    int64_t n_copy = ceil_div(total, (int64_t)kPackThreads * 4);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (n_copy > cap) n_copy = cap;
    if (n_copy < 1) n_copy = 1;
    pack_weight_kernel<<<(unsigned)(n_red + n_copy), kPackThreads, 0, as_stream(stream)>>>(
        W_fc, ldw, W_res, ldr, attn_l, attn_r, (int)H, (int)F, (int)K, out, ldo, n_red);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

extern "C" int spgnn_gat_pack_weight_bwd(const float* dP, int64_t ldp, const float* W_fc, int64_t ldw,
                                         const float* attn_l, const float* attn_r, int64_t H, int64_t F, int64_t K,
                                         int has_res, float* dW_fc, int64_t lddw, float* dW_res, int64_t lddr,
                                         float* d_attn_l, float* d_attn_r, void* stream) {
    SPGNN_REQUIRE(dP && W_fc && attn_l && attn_r && H > 0 && F > 0 && K > 0, "gat_pack_weight_bwd: bad argument");
    SPGNN_REQUIRE(ldp >= K && ldw >= K && (!dW_fc || lddw >= K) && (!dW_res || lddr >= K),
                  "gat_pack_weight_bwd: leading dimension smaller than K");
    const int64_t rows = H * F * (has_res ? 2 : 1);
    pack_weight_bwd_kernel<<<(unsigned)ceil_div(rows * 32, kPackThreads), kPackThreads, 0, as_stream(stream)>>>(
        dP, ldp, W_fc, ldw, attn_l, attn_r, (int)H, (int)F, (int)K, has_res, dW_fc, lddw, dW_res, lddr, d_attn_l,
        d_attn_r);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}
