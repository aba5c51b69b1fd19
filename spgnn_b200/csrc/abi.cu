// Error reporting, version and device queries of the C ABI.
#include "common.cuh"
#include <string.h>

namespace spgnn {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        cached = p.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}
static SaltSetter g_salt_setters[32];
static int g_n_salt_setters = 0;
SaltRegistrar::SaltRegistrar(SaltSetter fn) {
    if (g_n_salt_setters < 32) g_salt_setters[g_n_salt_setters++] = fn;
}
static unsigned long long device_bit() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return 0ull;      // unknown device: never cached
    return 1ull << dev;
}
bool DeviceOnce::pending() const { return (__atomic_load_n(&mask, __ATOMIC_ACQUIRE) & device_bit()) == 0ull; }
void DeviceOnce::done() { __atomic_fetch_or(&mask, device_bit(), __ATOMIC_RELEASE); }
}  // namespace spgnn

extern "C" const char* spgnn_last_error(void) { return spgnn::g_err; }
extern "C" int spgnn_abi_version(void) { return 1; }
extern "C" int64_t spgnn_launch_count(void) { return (int64_t)spgnn::launches(); }
__global__ void salt_zero_kernel(unsigned long long* p) { *p = 0ull; }

extern "C" int spgnn_seed_salt_set(const uint64_t* dev_value, void* stream) {
    SPGNN_REQUIRE(dev_value, "seed_salt_set: null pointer");
    for (int i = 0; i < spgnn::g_n_salt_setters; ++i) {
        const int e = spgnn::g_salt_setters[i](dev_value, spgnn::as_stream(stream));
        if (e != 0) {
            spgnn::set_error("seed_salt_set: cudaMemcpyToSymbolAsync failed: %s", cudaGetErrorString((cudaError_t)e));
            return SPGNN_E_CUDA;
        }
    }
    return SPGNN_OK;
}
extern "C" int spgnn_seed_salt_units(void) { return spgnn::g_n_salt_setters; }

extern "C" int spgnn_device_info(int* sm, int* major, int* minor) {
    int dev = 0;
    SPGNN_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    SPGNN_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    if (sm) *sm = p.multiProcessorCount;
    if (major) *major = p.major;
    if (minor) *minor = p.minor;
    return SPGNN_OK;
}
