// Device helpers shared by the GAT layer kernels (gat_layer.cu, gat_wide.cu): float4 arithmetic, the feat_drop mask
// convention of the planes pipeline, and split-bf16 plane loads / stores.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace spgnn {
namespace layer {
using ptx::join2;
using ptx::split2;

__device__ __forceinline__ float4 fma4(float s, float4 x, float4 acc) {
    acc.x = fmaf(s, x.x, acc.x); acc.y = fmaf(s, x.y, acc.y);
    acc.z = fmaf(s, x.z, acc.z); acc.w = fmaf(s, x.w, acc.w);
    return acc;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 scale4(float s, float4 a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 act4(float4 p, int act) {
    return make_float4(act_fast(p.x, act), act_fast(p.y, act), act_fast(p.z, act), act_fast(p.w, act));
}
__device__ __forceinline__ float4 actgrad4(float4 y, int act) {
    return make_float4(act_grad_from_out(y.x, act, 0.f), act_grad_from_out(y.y, act, 0.f),
                       act_grad_from_out(y.z, act, 0.f), act_grad_from_out(y.w, act, 0.f));
}
__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : slope * x; }
__device__ __forceinline__ uint64_t chunk_hash(uint64_t seed, uint64_t idx) { return mix64(seed ^ (idx * 0xD1B54A32D192ED03ull)); }

// feat_drop of a consumer on a 4-column chunk (16 hash bits per element; same convention as spgnn_split_planes)
__device__ __forceinline__ float4 drop4(float4 v, uint32_t thr, float scale, uint64_t seed, uint64_t chunk_idx) {
    if (!thr) return v;
    const uint64_t h = chunk_hash(seed, chunk_idx);
    v.x = ((uint32_t)(h) & 0xFFFFu) >= thr ? v.x * scale : 0.f;
    v.y = ((uint32_t)(h >> 16) & 0xFFFFu) >= thr ? v.y * scale : 0.f;
    v.z = ((uint32_t)(h >> 32) & 0xFFFFu) >= thr ? v.z * scale : 0.f;
    v.w = ((uint32_t)(h >> 48) & 0xFFFFu) >= thr ? v.w * scale : 0.f;
    return v;
}
__device__ __forceinline__ void store_planes4(__nv_bfloat16* hi, int64_t ps, float4 v) {
    uint32_t h0, l0, h1, l1;
    split2(v.x, v.y, h0, l0);
    split2(v.z, v.w, h1, l1);
    *reinterpret_cast<uint2*>(hi) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(hi + ps) = make_uint2(l0, l1);
}
__device__ __forceinline__ float4 load_planes4(const __nv_bfloat16* hi, int64_t ps) {
    const uint2 h = *reinterpret_cast<const uint2*>(hi);
    const uint2 l = *reinterpret_cast<const uint2*>(hi + ps);
    float4 v;
    join2(h.x, l.x, v.x, v.y);
    join2(h.y, l.y, v.z, v.w);
    return v;
}
__device__ __forceinline__ void store_planes1(__nv_bfloat16* hi, int64_t ps, float x) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[0] = h;
    hi[ps] = __float2bfloat16_rn(x - __bfloat162float(h));
}

}  // namespace layer
}  // namespace spgnn
