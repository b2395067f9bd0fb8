// Library-level entry points of libb200surv: version, architecture gate, error string.
#include <stdarg.h>

#include <atomic>

#include <cstdlib>

#include "common.cuh"

namespace b200surv {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
bool pdl_enabled() {
    static const bool on = [] { const char *e = getenv("B200SURV_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
void count_launches(int k) { g_launches.fetch_add((unsigned long long)k, std::memory_order_relaxed); }
unsigned long long launches_so_far() { return g_launches.load(std::memory_order_relaxed); }

int current_device() {
    int dev = -1;
    return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int num_sms() {
    static volatile int cached[MAX_DEVICES];  // 0 = not queried yet; writes of the same value race harmlessly
    const int dev = current_device();
    if (dev < 0) return 148;
    if (dev < MAX_DEVICES && cached[dev] > 0) return cached[dev];
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return 148;
    if (dev < MAX_DEVICES) cached[dev] = sms;
    return sms;
}

}  // namespace b200surv

extern "C" {

int32_t b200surv_version(void) { return 1000 * 0 + 1; }

const char *b200surv_last_error(void) { return b200surv::g_err; }

uint64_t b200surv_debug_launch_count(void) { return b200surv::launches_so_far(); }

int32_t b200surv_arch_check(int32_t device) {
    int major = 0, minor = 0;
    B200_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    B200_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    if (major != 10) {
        b200surv::set_error("device %d is sm_%d%d; libb200surv contains sm_100a code only", device, major,
                            minor);
        return B200SURV_UNSUPPORTED_ARCH;
    }
    return B200SURV_OK;
}

// ---- peer buffers: device memory of one rank that the other ranks of the box map through CUDA IPC
int32_t b200surv_peer_alloc(size_t bytes, void **dev_ptr, unsigned char *handle) {
    B200_REQUIRE(bytes > 0 && dev_ptr && handle, "bytes, dev_ptr, handle");
    static_assert(sizeof(cudaIpcMemHandle_t) == B200SURV_PEER_HANDLE_BYTES, "IPC handle size");
    void *p = nullptr;
    B200_CHECK_CUDA(cudaMalloc(&p, bytes));
    B200_CHECK_CUDA(cudaMemset(p, 0, bytes));
    B200_CHECK_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        b200surv::set_error("cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
        cudaFree(p);
        return B200SURV_CUDA_ERROR;
    }
    memcpy(handle, &h, sizeof(h));
    *dev_ptr = p;
    return B200SURV_OK;
}

int32_t b200surv_peer_open(const unsigned char *handle, void **dev_ptr) {
    B200_REQUIRE(handle && dev_ptr, "handle, dev_ptr");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    B200_CHECK_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return B200SURV_OK;
}

int32_t b200surv_peer_close(void *dev_ptr) {
    if (dev_ptr) B200_CHECK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return B200SURV_OK;
}

int32_t b200surv_peer_free(void *dev_ptr) {
    if (dev_ptr) B200_CHECK_CUDA(cudaFree(dev_ptr));
    return B200SURV_OK;
}

}  // extern "C"
