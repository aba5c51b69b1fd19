// Microbenchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16, operands in shared memory) as a function of N and
// of how many independent TMEM accumulators consecutive UMMAs rotate over.  Answers: is a chain of UMMAs into ONE
// accumulator latency-bound (the wide kernel issues 72 N = 128 UMMAs in a row into one accumulator)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spgnn_b200/csrc scripts/micro/umma_chain.cu -o /tmp/umma_chain
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace spgnn::ptx;

// fill != 0: warp 1 streams `fill` KB chunks global -> shared with cp.async.bulk (an L2-resident source) for as long as the
// UMMA chain runs: the operand-fetch path of the tensor core against the TMA write path into the same shared memory.
__global__ void __launch_bounds__(128, 1) chain_kernel(int n, int nacc, int iters, int same_operands, long long* out,
                                                       const uint8_t* src, int fill_kb, long long* filled) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint64_t fbar[2];
    __shared__ volatile int done;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sm = reinterpret_cast<uint8_t*>(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&fbar[0]), 1); mbar_init(smem_u32(&fbar[1]), 1); done = 0; fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_base_s), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base_s;
    const uint32_t base = smem_u32(sm);
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc(n, false);
        long long t0 = clock64();
        if (elect_one()) {
            for (int it = 0; it < iters; ++it) {
                const int acc = it % nacc;
                // 4 k-steps x 3 passes per "k-block": operands at different smem addresses as in the GEMMs
                const uint32_t ko = same_operands ? 0u : (uint32_t)((it & 3) * 32);
                const uint64_t da = make_desc(base + ko, 16, 1024);
                const uint64_t db = make_desc(base + 32768 + ko, 16, 1024);
                umma_bf16(tb + (uint32_t)(acc * n), da, db, idesc, it >= nacc);
            }
            umma_commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (threadIdx.x == 0) { out[blockIdx.x] = t1 - t0; done = 1; }
    } else if (threadIdx.x == 32 && fill_kb > 0) {
        // two chunks in flight, destination: the upper 64 KB of the buffer (not the operands)
        const uint32_t bytes = (uint32_t)fill_kb * 1024u;
        long long n_chunks = 0;
        uint32_t ph[2] = {0, 0};
        for (int i = 0; i < 2; ++i) {
            mbar_expect_tx(smem_u32(&fbar[i]), bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + 65536 + i * 16384),
                         "l"(src + ((size_t)(blockIdx.x * 2 + i) % 64) * 16384), "r"(bytes), "r"(smem_u32(&fbar[i])) : "memory");
        }
        while (!done) {
            for (int i = 0; i < 2; ++i) {
                mbar_wait(smem_u32(&fbar[i]), ph[i]); ph[i] ^= 1; ++n_chunks;
                mbar_expect_tx(smem_u32(&fbar[i]), bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + 65536 + i * 16384),
                             "l"(src + ((size_t)(blockIdx.x * 2 + i + n_chunks) % 64) * 16384), "r"(bytes), "r"(smem_u32(&fbar[i])) : "memory");
            }
        }
        for (int i = 0; i < 2; ++i) mbar_wait(smem_u32(&fbar[i]), ph[i]);
        filled[blockIdx.x] = n_chunks * (long long)bytes;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 148 * sizeof(long long));
    long long* f; cudaMalloc(&f, 148 * sizeof(long long));
    uint8_t* src; cudaMalloc(&src, 64 * 16384); cudaMemset(src, 0, 64 * 16384);
    cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int iters = 16384;
    printf("cycles per UMMA (M=128, K=16, bf16; %d UMMAs; floor = N/2 cycles)\n", iters);
    for (int grid : {1, 148}) {
        for (int n : {64, 128, 256}) {
            for (int nacc : {1, 2}) {
                if (n * nacc > 512) continue;
                for (int fill : {0, 4, 16}) {
                    const int same = 0;
                    cudaMemset(f, 0, 148 * sizeof(long long));
                    chain_kernel<<<grid, 128, 100 * 1024>>>(n, nacc, iters, same, d, src, fill, f);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    long long h[148]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                    double s = 0; for (int i = 0; i < grid; ++i) s += (double)h[i];
                    long long hf[148]; cudaMemcpy(hf, f, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                    double fb = 0; for (int i = 0; i < grid; ++i) fb += (double)hf[i];
                    printf("grid %3d  N=%3d  accumulators=%d  fill chunks %2d KB : %.1f cycles/UMMA (floor %d), fill %.1f B/clk/SM\n", grid, n, nacc,
                           fill, s / grid / iters, n / 2, fb / s);
                }
            }
        }
    }
    return 0;
}
