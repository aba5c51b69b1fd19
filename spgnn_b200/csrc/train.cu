// Callers of the path kept on device: masked weighted cross-entropy (job_runner.py:1885-1900) and the
// SGD-momentum update over a flat parameter bucket (torch.optim.SGD semantics, job_runner.py:1919).
#include "common.cuh"

namespace spgnn {

constexpr int kT = 256;

__device__ __forceinline__ bool keep_node(const int64_t* y, const uint8_t* mask, float rate, uint64_t seed, int64_t i) {
    if (mask) return mask[i] != 0;
    return y[i] != 0 || u01(seed, (uint64_t)i) < rate;
}

__global__ void __launch_bounds__(kT) masked_ce_fwd_kernel(const float* __restrict__ logits, int64_t ld, int C,
                                                           const int64_t* __restrict__ y,
                                                           const uint8_t* __restrict__ mask, float rate, uint64_t seed,
                                                           const float* __restrict__ cw, int64_t N,
                                                           double* __restrict__ sums) {
    __shared__ double sh[2][kT / 32];
    double nll = 0.0, wsum = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        if (!keep_node(y, mask, rate, seed, i)) continue;
        const float* row = logits + i * ld;
        float m = -INFINITY;
        for (int c = 0; c < C; ++c) m = fmaxf(m, __ldg(row + c));
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += expf(__ldg(row + c) - m);
        const int64_t t = y[i];
        const float w = __ldg(cw + t);
        nll += (double)(w * (logf(s) + m - __ldg(row + t)));
        wsum += (double)w;
    }
    nll = warp_sum(nll);
    wsum = warp_sum(wsum);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sh[0][wid] = nll; sh[1][wid] = wsum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < kT / 32; ++k) { a += sh[0][k]; b += sh[1][k]; }
        atomicAdd(&sums[0], a);
        atomicAdd(&sums[1], b);
    }
}

__global__ void __launch_bounds__(kT) masked_ce_bwd_kernel(const float* __restrict__ logits, int64_t ld, int C,
                                                           const int64_t* __restrict__ y,
                                                           const uint8_t* __restrict__ mask, float rate, uint64_t seed,
                                                           const float* __restrict__ cw,
                                                           const double* __restrict__ sums, float scale, int64_t N,
                                                           float* __restrict__ dl, int64_t ldd) {
    const float k = (float)((double)scale / sums[1]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        float* drow = dl + i * ldd;
        if (!keep_node(y, mask, rate, seed, i)) {
            for (int c = 0; c < C; ++c) drow[c] = 0.f;
            continue;
        }
        const float* row = logits + i * ld;
        float m = -INFINITY;
        for (int c = 0; c < C; ++c) m = fmaxf(m, __ldg(row + c));
        float s = 0.f;
        for (int c = 0; c < C; ++c) s += expf(__ldg(row + c) - m);
        const int64_t t = y[i];
        const float w = __ldg(cw + t) * k;
        const float inv = 1.f / s;
        for (int c = 0; c < C; ++c) drow[c] = w * (expf(__ldg(row + c) - m) * inv - (c == t ? 1.f : 0.f));
    }
}

__global__ void sgd_momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                                    int64_t n, float lr, float mu, float gs, int first) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * gs;
        const float b = first ? gi : fmaf(mu, buf[i], gi);
        buf[i] = b;
        p[i] -= lr * b;
    }
}

// torch.optim.SGD._single_tensor_sgd (torch/optim/sgd.py): g += wd*p; buf = first ? g : mu*buf + (1-damp)*g;
// g = nesterov ? g + mu*buf : buf (only with momentum); p -= lr*g
__global__ void sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, int64_t n,
                                float lr, float mu, float damp, float wd, int nesterov, float gs, int first) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float pi = p[i];
        float gi = g[i] * gs;
        if (wd != 0.f) gi = fmaf(wd, pi, gi);
        if (mu != 0.f) {
            const float b = first ? gi : fmaf(mu, buf[i], (1.f - damp) * gi);
            buf[i] = b;
            gi = nesterov ? fmaf(mu, b, gi) : b;
        }
        p[i] = pi - lr * gi;
    }
}

static inline unsigned egrid(int64_t n) {
    int64_t want = ceil_div(n, kT), cap = (int64_t)sm_count() * 8;
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

}  // namespace spgnn

using namespace spgnn;

extern "C" int spgnn_masked_ce_fwd(const float* logits, int64_t ld, int64_t n_class, const int64_t* y,
                                   const uint8_t* mask, float rate, uint64_t seed, const float* class_w, int64_t N,
                                   double* sums, void* stream) {
    SPGNN_REQUIRE(logits && y && class_w && sums && N > 0 && n_class > 1 && ld >= n_class, "masked_ce_fwd: bad argument");
    cudaStream_t st = as_stream(stream);
    SPGNN_CUDA_OK(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st));
    masked_ce_fwd_kernel<<<egrid(N), kT, 0, st>>>(logits, ld, (int)n_class, y, mask, rate, seed, class_w, N, sums);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

extern "C" int spgnn_masked_ce_bwd(const float* logits, int64_t ld, int64_t n_class, const int64_t* y,
                                   const uint8_t* mask, float rate, uint64_t seed, const float* class_w,
                                   const double* sums, float scale, int64_t N, float* dlogits, int64_t ldd,
                                   void* stream) {
    SPGNN_REQUIRE(logits && y && class_w && sums && dlogits && N > 0 && n_class > 1 && ld >= n_class && ldd >= n_class,
                  "masked_ce_bwd: bad argument");
    masked_ce_bwd_kernel<<<egrid(N), kT, 0, as_stream(stream)>>>(logits, ld, (int)n_class, y, mask, rate, seed, class_w,
                                                                 sums, scale, N, dlogits, ldd);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

extern "C" int spgnn_sgd_step(float* p, const float* g, float* buf, int64_t n, float lr, float mu, float dampening,
                              float weight_decay, int nesterov, float grad_scale, int first_step, void* stream) {
    SPGNN_REQUIRE(p && g && n > 0 && (buf || mu == 0.f), "sgd_step: bad argument");
    SPGNN_REQUIRE(!nesterov || (mu > 0.f && dampening == 0.f), "sgd_step: nesterov needs momentum > 0 and dampening == 0");
    sgd_step_kernel<<<egrid(n), kT, 0, as_stream(stream)>>>(p, g, buf, n, lr, mu, dampening, weight_decay, nesterov,
                                                           grad_scale, first_step);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

extern "C" int spgnn_sgd_momentum(float* p, const float* g, float* buf, int64_t n, float lr, float mu,
                                  float grad_scale, int first_step, void* stream) {
    SPGNN_REQUIRE(p && g && buf && n > 0, "sgd_momentum: bad argument");
    sgd_momentum_kernel<<<egrid(n), kT, 0, as_stream(stream)>>>(p, g, buf, n, lr, mu, grad_scale, first_step);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

SPGNN_REGISTER_SALT(train)
