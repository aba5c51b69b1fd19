// GAT layer kernels of the planes pipeline (spgnn_gat_layer_fwd / _bwd): fused edge-softmax + neighbour aggregation +
// residual + bias + activation (+ head mean), reading the fp32 projection output Y and WRITING the next projections'
// operands directly as split-bf16 planes — with each consumer's feat_drop mask already applied — so neither
// torch.cat, nor dropout, nor any fp32 -> bf16 conversion ever makes a separate pass over HBM.
//
// Structure: a CTA owns a chunk of kNPC consecutive nodes (nodes of a tree are contiguous, so neighbour rows hit
// L1/L2).  Phase A is thread-parallel over (node, head): it walks the CSC/CSR segment, computes the edge softmax
// (or, in backward, re-reads it) and stages source ids and weights of the (<= 4, airway trees: <= 3 + self loop)
// edges in shared memory.  Phase B is warp-per-node over 128-bit column chunks with every address and weight coming
// from shared memory: no dependent global loads in the streaming loop.  Nodes with more than 4 edges take a generic
// loop in the same kernel.  The backward destination kernel has a third thread-parallel phase for the softmax /
// LeakyReLU backward, and accumulates the bias gradient in shared memory (per-CTA partial rows, fixed-order reduce).
//
// HBM roofline: forward 4*N*(HF_z + HF_res + 2H) read + 4*N*W_out per sink written;
// backward 4*N*(W_g + 2*HF + HF_res) read + 4*N*(2*HF (+HF_res)) written, + indices.
#include "layer_util.cuh"

namespace spgnn {
namespace layer {
#ifndef SPGNN_NPC
#define SPGNN_NPC 128
#endif
constexpr int kNPC = SPGNN_NPC;  // nodes per CTA chunk
constexpr int kThreads = 256;
constexpr int kMaxH = 8;

// shared-memory staging of one chunk
struct Stage {
    int* deg;        // [kNPC]
    int* beg;        // [kNPC]
    int* nb;         // [kNPC][4] neighbour ids (sources in fwd / bwd-dst, destinations in bwd-src)
    float* w;        // [kNPC][H][4] attention weight after dropout scaling (0 beyond the degree)
    float* at;       // [kNPC][H][4] softmax weight before dropout        (bwd-dst)
    float* lk;       // [kNPC][H][4] LeakyReLU slope factor of the logit   (bwd-dst)
    float* dd;       // [kNPC][H][4] <g, z_j>                               (bwd-dst)
    float* sb;       // [H*F] bias-gradient accumulator                      (bwd-dst)
};
__device__ __forceinline__ Stage carve(uint8_t* smem, int H, int HF) {
    Stage s;
    s.deg = reinterpret_cast<int*>(smem);
    s.beg = s.deg + kNPC;
    s.nb = s.beg + kNPC;
    s.w = reinterpret_cast<float*>(s.nb + 4 * kNPC);
    s.at = s.w + kNPC * H * 4;
    s.lk = s.at + kNPC * H * 4;
    s.dd = s.lk + kNPC * H * 4;
    s.sb = s.dd + kNPC * H * 4;
    return s;
}
static size_t stage_bytes(int H, int HF, bool bwd_dst) {
    size_t b = (size_t)kNPC * 6 * 4 + (size_t)kNPC * H * 4 * 4 * (bwd_dst ? 4 : 1);
    if (bwd_dst) b += (size_t)HF * 4;
    return b + 16;
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(kThreads) gat_layer_fwd_kernel(const Args a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const Stage st = carve(smem, a.H, a.H * a.F);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = a.H, F = a.F;
    const int64_t nchunks = (a.N + kNPC - 1) / kNPC;
    const float inv_h = 1.f / (float)H;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int64_t base = chunk * kNPC;
        // ---------------- phase A: edge softmax, one thread per (node, head)
        for (int it = threadIdx.x; it < kNPC * H; it += kThreads) {
            const int n = it / H, h = it - n * H;
            const int64_t v = base + n;
            if (v >= a.N) continue;
            const int beg = __ldg(a.in_ptr + v), deg = __ldg(a.in_ptr + v + 1) - beg;
            const float er = __ldg(a.Y + v * a.ldy + a.er_off + h);
            if (h == 0) { st.deg[n] = deg; st.beg[n] = beg; }
            if (deg <= 4) {
                int u[4];
                float e[4], m = -INFINITY;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    u[j] = deg > 0 ? __ldg(a.in_src + beg + min(j, deg - 1)) : (int)v;
                    e[j] = leaky(__ldg(a.Y + (int64_t)u[j] * a.ldy + a.el_off + h) + er, a.neg_slope);
                    if (j < deg) m = fmaxf(m, e[j]);
                }
                float p[4], sum = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) { p[j] = j < deg ? __expf(e[j] - m) : 0.f; sum += p[j]; }
                const float inv = deg > 0 ? 1.f / sum : 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float at = p[j] * inv;
                    if (j < deg) a.att[(int64_t)(beg + j) * H + h] = at;
                    st.w[(n * H + h) * 4 + j] = j < deg ? at * keep_scale(a, beg + j, h) : 0.f;
                    if (h == 0) st.nb[n * 4 + j] = u[j];
                }
            } else {
                float m = -INFINITY;
                for (int s = beg; s < beg + deg; ++s)
                    m = fmaxf(m, leaky(__ldg(a.Y + (int64_t)__ldg(a.in_src + s) * a.ldy + a.el_off + h) + er, a.neg_slope));
                float sum = 0.f;
                for (int s = beg; s < beg + deg; ++s)
                    sum += __expf(leaky(__ldg(a.Y + (int64_t)__ldg(a.in_src + s) * a.ldy + a.el_off + h) + er, a.neg_slope) - m);
                const float inv = 1.f / sum;
                for (int s = beg; s < beg + deg; ++s)
                    a.att[(int64_t)s * H + h] =
                        __expf(leaky(__ldg(a.Y + (int64_t)__ldg(a.in_src + s) * a.ldy + a.el_off + h) + er, a.neg_slope) - m) * inv;
            }
        }
        __syncthreads();
        // ---------------- phase B: aggregation, one warp per node
        for (int n = warp; n < kNPC; n += kThreads / 32) {
            const int64_t v = base + n;
            if (v >= a.N) break;
            const int deg = st.deg[n];
            const float* yv = a.Y + v * a.ldy;
            if (deg <= 4) {
                const int4 nb = *reinterpret_cast<const int4*>(st.nb + n * 4);
                const float* r0 = a.Y + (int64_t)nb.x * a.ldy;
                const float* r1 = a.Y + (int64_t)nb.y * a.ldy;
                const float* r2 = a.Y + (int64_t)nb.z * a.ldy;
                const float* r3 = a.Y + (int64_t)nb.w * a.ldy;
                if (a.mean_heads) {
                    for (int col = lane * 4; col < F; col += 128) {
                        float4 mean = zero4();
                        for (int h = 0; h < H; ++h) {
                            const float4 w = *reinterpret_cast<const float4*>(st.w + (n * H + h) * 4);
                            const int hc = h * F + col;
                            const float4 z0 = ldg4(r0 + hc), z1 = ldg4(r1 + hc), z2 = ldg4(r2 + hc), z3 = ldg4(r3 + hc);
                            float4 acc = a.res_mode == 1 ? ldg4(yv + a.res_off + hc) : zero4();
                            if (a.bias) acc = add4(acc, ldg4(a.bias + hc));
                            acc = fma4(w.x, z0, acc); acc = fma4(w.y, z1, acc);
                            acc = fma4(w.z, z2, acc); acc = fma4(w.w, z3, acc);
                            mean = add4(mean, act4(acc, a.act));
                        }
                        emit(a, v, col, scale4(inv_h, mean));
                    }
                } else {
                    for (int h = 0; h < H; ++h) {
                        const float4 w = *reinterpret_cast<const float4*>(st.w + (n * H + h) * 4);
                        for (int col = lane * 4; col < F; col += 128) {
                            const int hc = h * F + col;
                            const float4 z0 = ldg4(r0 + hc), z1 = ldg4(r1 + hc), z2 = ldg4(r2 + hc), z3 = ldg4(r3 + hc);
                            float4 acc = a.res_mode == 1 ? ldg4(yv + a.res_off + hc) : zero4();
                            if (a.bias) acc = add4(acc, ldg4(a.bias + hc));
                            acc = fma4(w.x, z0, acc); acc = fma4(w.y, z1, acc);
                            acc = fma4(w.z, z2, acc); acc = fma4(w.w, z3, acc);
                            emit(a, v, hc, act4(acc, a.act));
                        }
                    }
                }
            } else {
                // generic degree: weights from att[] (written by this CTA in phase A: plain loads)
                const int beg = st.beg[n];
                for (int col = lane * 4; col < F; col += 128) {
                    float4 mean = zero4();
                    for (int h = 0; h < H; ++h) {
                        const int hc = h * F + col;
                        float4 acc = a.res_mode == 1 ? ldg4(yv + a.res_off + hc) : zero4();
                        if (a.bias) acc = add4(acc, ldg4(a.bias + hc));
                        for (int s = beg; s < beg + deg; ++s) {
                            const float w = a.att[(int64_t)s * H + h] * keep_scale(a, s, h);
                            acc = fma4(w, ldg4(a.Y + (int64_t)__ldg(a.in_src + s) * a.ldy + hc), acc);
                        }
                        const float4 y = act4(acc, a.act);
                        if (a.mean_heads) mean = add4(mean, y);
                        else emit(a, v, hc, y);
                    }
                    if (a.mean_heads) emit(a, v, col, scale4(inv_h, mean));
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ backward, dst side
// gradient of the layer output chunk (v, columns c..c+3): sum over consumers of mask * d(consumer input)
__device__ __forceinline__ void store_G(const Args& a, int64_t v, int hc, float4 g) {
    if (a.res_mode == 1) store_planes4(a.dY + v * a.dld + a.res_off + hc, a.dps, g);
    else st4(a.g_ws + v * (int64_t)(a.H * a.F) + hc, g);
}
__device__ __forceinline__ float4 load_G(const Args& a, int64_t v, int hc) {
    if (a.res_mode == 1) return load_planes4(a.dY + v * a.dld + a.res_off + hc, a.dps);
    return *reinterpret_cast<const float4*>(a.g_ws + v * (int64_t)(a.H * a.F) + hc);
}

__global__ void __launch_bounds__(kThreads) gat_layer_bwd_dst_kernel(const Args a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int H = a.H, F = a.F, HF = H * F;
    const Stage st = carve(smem, H, HF);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nchunks = (a.N + kNPC - 1) / kNPC;
    const float inv_h = 1.f / (float)H;
    for (int i = threadIdx.x; i < HF; i += kThreads) st.sb[i] = 0.f;
    __syncthreads();
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int64_t base = chunk * kNPC;
        // ---------------- phase A: stage edges, weights and LeakyReLU slopes
        for (int it = threadIdx.x; it < kNPC * H; it += kThreads) {
            const int n = it / H, h = it - n * H;
            const int64_t v = base + n;
            if (v >= a.N) continue;
            const int beg = __ldg(a.in_ptr + v), deg = __ldg(a.in_ptr + v + 1) - beg;
            if (h == 0) { st.deg[n] = deg; st.beg[n] = beg; }
            if (deg <= 4) {
                const float er = __ldg(a.Y + v * a.ldy + a.er_off + h);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int u = deg > 0 ? __ldg(a.in_src + beg + min(j, deg - 1)) : (int)v;
                    const float raw = __ldg(a.Y + (int64_t)u * a.ldy + a.el_off + h) + er;
                    const float at = j < deg ? __ldg(a.att + (int64_t)(beg + j) * H + h) : 0.f;
                    const int o = (n * H + h) * 4 + j;
                    st.at[o] = at;
                    st.w[o] = j < deg ? at * keep_scale(a, beg + j, h) : 0.f;
                    st.lk[o] = raw > 0.f ? 1.f : a.neg_slope;
                    if (h == 0) st.nb[n * 4 + j] = u;
                }
            }
        }
        __syncthreads();
        // ---------------- phase B: g = g_out * act'(y) (y recomputed), G store, <g, z_j>, bias gradient
        for (int n = warp; n < kNPC; n += kThreads / 32) {
            const int64_t v = base + n;
            if (v >= a.N) break;
            const int deg = st.deg[n];
            const float* yv = a.Y + v * a.ldy;
            if (deg <= 4) {
                const int4 nb = *reinterpret_cast<const int4*>(st.nb + n * 4);
                const float* r0 = a.Y + (int64_t)nb.x * a.ldy;
                const float* r1 = a.Y + (int64_t)nb.y * a.ldy;
                const float* r2 = a.Y + (int64_t)nb.z * a.ldy;
                const float* r3 = a.Y + (int64_t)nb.w * a.ldy;
                for (int h = 0; h < H; ++h) {
                    const float4 w = *reinterpret_cast<const float4*>(st.w + (n * H + h) * 4);
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
                    for (int col = lane * 4; col < F; col += 128) {
                        const int hc = h * F + col;
                        const float4 z0 = ldg4(r0 + hc), z1 = ldg4(r1 + hc), z2 = ldg4(r2 + hc), z3 = ldg4(r3 + hc);
                        float4 acc = a.res_mode == 1 ? ldg4(yv + a.res_off + hc) : zero4();
                        if (a.bias) acc = add4(acc, ldg4(a.bias + hc));
                        float4 go = load_g(a, v, a.mean_heads ? col : hc);
                        if (a.mean_heads) go = scale4(inv_h, go);
                        acc = fma4(w.x, z0, acc); acc = fma4(w.y, z1, acc);
                        acc = fma4(w.z, z2, acc); acc = fma4(w.w, z3, acc);
                        const float4 gq = mul4(go, actgrad4(act4(acc, a.act), a.act));
                        store_G(a, v, hc, gq);
                        if (a.dbias_ws) {
                            atomicAdd(st.sb + hc, gq.x); atomicAdd(st.sb + hc + 1, gq.y);
                            atomicAdd(st.sb + hc + 2, gq.z); atomicAdd(st.sb + hc + 3, gq.w);
                        }
                        d0 += dot4(gq, z0); d1 += dot4(gq, z1); d2 += dot4(gq, z2); d3 += dot4(gq, z3);
                    }
                    d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2); d3 = warp_sum(d3);
                    if (lane == 0) *reinterpret_cast<float4*>(st.dd + (n * H + h) * 4) = make_float4(d0, d1, d2, d3);
                }
            } else {
                const int beg = st.beg[n];
                for (int h = 0; h < H; ++h) {
                    // pass 1: g chunks (kept in global G), pass 2: per-edge dot products staged in ds[]
                    for (int col = lane * 4; col < F; col += 128) {
                        const int hc = h * F + col;
                        float4 acc = a.res_mode == 1 ? ldg4(yv + a.res_off + hc) : zero4();
                        if (a.bias) acc = add4(acc, ldg4(a.bias + hc));
                        for (int s = beg; s < beg + deg; ++s) {
                            const float w = __ldg(a.att + (int64_t)s * H + h) * keep_scale(a, s, h);
                            acc = fma4(w, ldg4(a.Y + (int64_t)__ldg(a.in_src + s) * a.ldy + hc), acc);
                        }
                        float4 go = load_g(a, v, a.mean_heads ? col : hc);
                        if (a.mean_heads) go = scale4(inv_h, go);
                        const float4 gq = mul4(go, actgrad4(act4(acc, a.act), a.act));
                        store_G(a, v, hc, gq);
                        if (a.dbias_ws) {
                            atomicAdd(st.sb + hc, gq.x); atomicAdd(st.sb + hc + 1, gq.y);
                            atomicAdd(st.sb + hc + 2, gq.z); atomicAdd(st.sb + hc + 3, gq.w);
                        }
                    }
                    __syncwarp();
                    for (int s = beg; s < beg + deg; ++s) {
                        const int u = __ldg(a.in_src + s);
                        float d = 0.f;
                        for (int col = lane * 4; col < F; col += 128) {
                            const int hc = h * F + col;
                            d += dot4(load_G(a, v, hc), ldg4(a.Y + (int64_t)u * a.ldy + hc));
                        }
                        d = warp_sum(d);
                        if (lane == 0) a.ds[(int64_t)s * H + h] = d * keep_scale(a, s, h);
                    }
                }
            }
        }
        __syncthreads();
        // ---------------- phase C: softmax + LeakyReLU backward, one thread per (node, head)
        for (int it = threadIdx.x; it < kNPC * H; it += kThreads) {
            const int n = it / H, h = it - n * H;
            const int64_t v = base + n;
            if (v >= a.N) continue;
            const int deg = st.deg[n], beg = st.beg[n];
            float der = 0.f;
            if (deg <= 4) {
                const int o = (n * H + h) * 4;
                float da[4], wsum = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float at = st.at[o + j];
                    // d(a_drop)/d(a) = keep/(1-p) = w/att (0 when dropped or beyond the degree)
                    da[j] = at > 0.f ? st.dd[o + j] * (st.w[o + j] / at) : 0.f;
                    wsum += at * da[j];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j < deg) {
                        const float dsv = st.at[o + j] * (da[j] - wsum) * st.lk[o + j];
                        a.ds[(int64_t)(beg + j) * H + h] = dsv;
                        der += dsv;
                    }
                }
            } else {
                const float er = __ldg(a.Y + v * a.ldy + a.er_off + h);
                float wsum = 0.f;
                for (int s = beg; s < beg + deg; ++s) wsum += __ldg(a.att + (int64_t)s * H + h) * a.ds[(int64_t)s * H + h];
                for (int s = beg; s < beg + deg; ++s) {
                    const float raw = __ldg(a.Y + (int64_t)__ldg(a.in_src + s) * a.ldy + a.el_off + h) + er;
                    const float dsv = __ldg(a.att + (int64_t)s * H + h) * (a.ds[(int64_t)s * H + h] - wsum) *
                                      (raw > 0.f ? 1.f : a.neg_slope);
                    a.ds[(int64_t)s * H + h] = dsv;
                    der += dsv;
                }
            }
            store_planes1(a.dY + v * a.dld + a.er_off + h, a.dps, der);
        }
        __syncthreads();
    }
    if (a.dbias_ws)
        for (int i = threadIdx.x; i < HF; i += kThreads) a.dbias_ws[(int64_t)blockIdx.x * HF + i] = st.sb[i];
}

// ------------------------------------------------------------------------------------------------ backward, src side
//   del[u,h] = sum over out-edges of ds;   dz[u,h,:] = sum over out-edges (u->v) of a_drop * g[v,h,:]
__global__ void __launch_bounds__(kThreads) gat_layer_bwd_src_kernel(const Args a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int H = a.H, F = a.F;
    const Stage st = carve(smem, H, H * F);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nchunks = (a.N + kNPC - 1) / kNPC;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int64_t base = chunk * kNPC;
        for (int it = threadIdx.x; it < kNPC * H; it += kThreads) {
            const int n = it / H, h = it - n * H;
            const int64_t u = base + n;
            if (u >= a.N) continue;
            const int beg = __ldg(a.out_ptr + u), deg = __ldg(a.out_ptr + u + 1) - beg;
            if (h == 0) { st.deg[n] = deg; st.beg[n] = beg; }
            float del = 0.f;
            if (deg <= 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int q = beg + min(j, max(deg - 1, 0));
                    const int s = deg > 0 ? __ldg(a.out_slot + q) : 0;
                    const int v = deg > 0 ? __ldg(a.out_dst + q) : (int)u;
                    const bool on = j < deg;
                    st.w[(n * H + h) * 4 + j] = on ? __ldg(a.att + (int64_t)s * H + h) * keep_scale(a, s, h) : 0.f;
                    if (on) del += a.ds[(int64_t)s * H + h];
                    if (h == 0) st.nb[n * 4 + j] = v;
                }
            } else {
                for (int q = beg; q < beg + deg; ++q) del += a.ds[(int64_t)__ldg(a.out_slot + q) * H + h];
            }
            store_planes1(a.dY + u * a.dld + a.el_off + h, a.dps, del);
        }
        __syncthreads();
        for (int n = warp; n < kNPC; n += kThreads / 32) {
            const int64_t u = base + n;
            if (u >= a.N) break;
            const int deg = st.deg[n];
            __nv_bfloat16* drow = a.dY + u * a.dld;
            if (deg <= 4) {
                const int4 nb = *reinterpret_cast<const int4*>(st.nb + n * 4);
                for (int h = 0; h < H; ++h) {
                    const float4 w = *reinterpret_cast<const float4*>(st.w + (n * H + h) * 4);
                    for (int col = lane * 4; col < F; col += 128) {
                        const int hc = h * F + col;
                        const float4 g0 = load_G(a, nb.x, hc), g1 = load_G(a, nb.y, hc), g2 = load_G(a, nb.z, hc),
                                     g3 = load_G(a, nb.w, hc);
                        float4 acc = scale4(w.x, g0);
                        acc = fma4(w.y, g1, acc); acc = fma4(w.z, g2, acc); acc = fma4(w.w, g3, acc);
                        store_planes4(drow + hc, a.dps, acc);
                    }
                }
            } else {
                const int beg = st.beg[n];
                for (int h = 0; h < H; ++h) {
                    for (int col = lane * 4; col < F; col += 128) {
                        const int hc = h * F + col;
                        float4 acc = zero4();
                        for (int q = beg; q < beg + deg; ++q) {
                            const int s = __ldg(a.out_slot + q);
                            const float w = __ldg(a.att + (int64_t)s * H + h) * keep_scale(a, s, h);
                            acc = fma4(w, load_G(a, __ldg(a.out_dst + q), hc), acc);
                        }
                        store_planes4(drow + hc, a.dps, acc);
                    }
                }
            }
        }
        __syncthreads();
    }
}

// fixed-order sum of the per-CTA bias-gradient rows
__global__ void dbias_reduce_kernel(const float* __restrict__ part, int64_t nparts, int64_t HF, float* __restrict__ out) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < HF; c += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int64_t p = 0; p < nparts; ++p) s += part[p * HF + c];
        out[c] = s;
    }
}

static unsigned layer_grid(int64_t N) {
    const int64_t chunks = ceil_div(N, kNPC), cap = (int64_t)sm_count() * 8;
    return (unsigned)(chunks < cap ? chunks : cap);
}

}  // namespace layer
}  // namespace spgnn

namespace spgnn {
namespace tree {
int launch_fwd(const layer::Args& a, const spgnn_gat_layer* L, cudaStream_t st, bool* handled);
int launch_bwd(const layer::Args& a, const spgnn_gat_layer* L, cudaStream_t st, bool* handled);
}
}  // namespace spgnn

using namespace spgnn;
using namespace spgnn::layer;

extern "C" int64_t spgnn_gat_layer_sizeof(void) { return (int64_t)sizeof(spgnn_gat_layer); }

extern "C" int64_t spgnn_gat_layer_dbias_ws(int64_t N, int64_t H, int64_t F) {
    // one partial row per CTA: the chunk kernels run layer_grid(N) CTAs, the per-tree kernels min(B, #SMs) — with many
    // tiny graphs (B > N / kNPC) the latter is the larger one (a batch of 1-, 2- and 3-node trees overran a workspace
    // sized by the former: tests/test_gpu_parity.py::test_gat3_on_edge_case_graphs_vs_oracle[tiny_trees])
    const int64_t rows = (int64_t)layer_grid(N) > (int64_t)sm_count() ? (int64_t)layer_grid(N) : (int64_t)sm_count();
    return rows * H * F * (int64_t)sizeof(float);
}

extern "C" int spgnn_gat_layer_fwd(const spgnn_gat_layer* L, void* stream) {
    Args a{};
    int rc = fill_common(a, L);
    if (rc) return rc;
    SPGNN_REQUIRE(L->n_sinks >= 0 && L->n_sinks <= 2 && (L->out || L->n_sinks > 0), "gat_layer_fwd: no output");
    const int W = L->mean_heads ? L->F : L->H * L->F;
    SPGNN_REQUIRE(!L->out || (L->ldo % 4 == 0 && L->ldo >= W && ((uintptr_t)L->out & 15) == 0), "gat_layer_fwd: out alignment");
    a.out = L->out; a.ldo = L->ldo; a.n_sinks = L->n_sinks;
    for (int s = 0; s < L->n_sinks; ++s) {
        const spgnn_sink& k = L->sinks[s];
        SPGNN_REQUIRE(k.hi && k.ld % 4 == 0 && k.ld >= W && k.plane_stride % 4 == 0 && ((uintptr_t)k.hi & 7) == 0 &&
                          k.drop_p >= 0.f && k.drop_p < 1.f,
                      "gat_layer_fwd: sink %d: ld (%lld) / plane stride must be multiples of 4", s, (long long)k.ld);
        a.sinks[s] = Sink{reinterpret_cast<__nv_bfloat16*>(k.hi), k.ld, k.plane_stride, k.concat_chunks, k.chunk_off,
                          thr_of(k.drop_p), k.drop_p > 0.f ? 1.f / (1.f - k.drop_p) : 1.f, k.seed};
    }
    bool handled = false;
    rc = tree::launch_fwd(a, L, as_stream(stream), &handled);
    if (rc || handled) return rc;
    const size_t smem = stage_bytes(a.H, a.H * a.F, false);
    gat_layer_fwd_kernel<<<layer_grid(a.N), kThreads, smem, as_stream(stream)>>>(a);
    SPGNN_LAUNCH_OK();
    return SPGNN_OK;
}

extern "C" int spgnn_gat_layer_bwd(const spgnn_gat_layer* L, void* stream) {
    Args a{};
    int rc = fill_common(a, L);
    if (rc) return rc;
    SPGNN_REQUIRE(L->out_ptr && L->out_dst && L->out_slot && L->dY_hi && L->ds_ws, "gat_layer_bwd: null pointer");
    SPGNN_REQUIRE(L->n_gsrc >= 1 && L->n_gsrc <= 3, "gat_layer_bwd: 1..3 gradient sources");
    SPGNN_REQUIRE(L->res_mode == 1 || L->g_ws, "gat_layer_bwd: g_ws [N, H*F] required without a linear residual");
    SPGNN_REQUIRE(L->dY_ld % 4 == 0 && L->dY_ps % 4 == 0 && ((uintptr_t)L->dY_hi & 7) == 0 && L->dY_ld >= L->er_off + L->H,
                  "gat_layer_bwd: dY planes ld (%lld) too small or misaligned", (long long)L->dY_ld);
    const int HF = L->H * L->F;
    SPGNN_REQUIRE(HF * 4 + stage_bytes(L->H, 0, true) <= 200 * 1024, "gat_layer_bwd: H*F = %d too large", HF);
    a.n_g = L->n_gsrc;
    for (int s = 0; s < L->n_gsrc; ++s) {
        const spgnn_gsrc& k = L->gsrc[s];
        SPGNN_REQUIRE(k.g && k.ld % 4 == 0 && ((uintptr_t)k.g & 15) == 0 && k.drop_p >= 0.f && k.drop_p < 1.f,
                      "gat_layer_bwd: gradient source %d must be 16-byte aligned with ld %% 4 == 0", s);
        a.gs[s] = GSrc{k.g, k.ld, k.concat_chunks, k.chunk_off, thr_of(k.drop_p),
                       k.drop_p > 0.f ? 1.f / (1.f - k.drop_p) : 1.f, k.seed};
    }
    a.dY = reinterpret_cast<__nv_bfloat16*>(L->dY_hi); a.dld = L->dY_ld; a.dps = L->dY_ps;
    a.g_ws = L->g_ws; a.ds = L->ds_ws; a.dbias_ws = L->dbias ? L->dbias_ws : nullptr;
    SPGNN_REQUIRE(!L->dbias || L->dbias_ws, "gat_layer_bwd: dbias needs dbias_ws (spgnn_gat_layer_dbias_ws bytes)");
    cudaStream_t st = as_stream(stream);
    bool handled = false;
    rc = tree::launch_bwd(a, L, st, &handled);
    if (rc || handled) return rc;
    const unsigned grid = layer_grid(a.N);
    const size_t smem_dst = stage_bytes(a.H, HF, true);
    static DeviceOnce attr;
    if (attr.pending()) {
        SPGNN_CUDA_OK(cudaFuncSetAttribute(gat_layer_bwd_dst_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr.done();
    }
    gat_layer_bwd_dst_kernel<<<grid, kThreads, smem_dst, st>>>(a);
    SPGNN_LAUNCH_OK();
    gat_layer_bwd_src_kernel<<<grid, kThreads, stage_bytes(a.H, HF, false), st>>>(a);
    SPGNN_LAUNCH_OK();
    if (L->dbias) {
        dbias_reduce_kernel<<<(unsigned)ceil_div(HF, 128), 128, 0, st>>>(L->dbias_ws, grid, HF, L->dbias);
        SPGNN_LAUNCH_OK();
    }
    return SPGNN_OK;
}

SPGNN_REGISTER_SALT(gat_layer)
