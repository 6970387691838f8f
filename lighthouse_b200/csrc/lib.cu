// lib.cu — lifecycle and error plumbing of liblhb200.so.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include "ctx.h"

namespace lhb200 {

static thread_local char t_err[512] = "";

Ctx& ctx() {
    static Ctx c;
    return c;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}

int32_t cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    if (e == cudaErrorMemoryAllocation) return LHB200_ENOMEM;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) return LHB200_ENODEV;
    return LHB200_ECUDA;
}

void* dev_scratch(size_t nbytes) {
    Ctx& c = ctx();
    if (nbytes <= c.d_scratch_bytes) return c.d_scratch;
    if (c.d_scratch) {
        cudaStreamSynchronize(c.stream);
        cudaFree(c.d_scratch);
        c.d_scratch = nullptr;
        c.d_scratch_bytes = 0;
    }
    size_t want = nbytes + (nbytes >> 2) + (1 << 20);
    cudaError_t e = cudaMalloc(&c.d_scratch, want);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaMalloc(scratch)");
        return nullptr;
    }
    c.d_scratch_bytes = want;
    return c.d_scratch;
}

void* pinned_scratch(size_t nbytes) {
    Ctx& c = ctx();
    if (nbytes <= c.h_pinned_bytes) return c.h_pinned;
    if (c.h_pinned) {
        cudaStreamSynchronize(c.stream);
        cudaFreeHost(c.h_pinned);
        c.h_pinned = nullptr;
        c.h_pinned_bytes = 0;
    }
    size_t want = nbytes + (nbytes >> 2) + (1 << 16);
    cudaError_t e = cudaHostAlloc(&c.h_pinned, want, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaHostAlloc(staging)");
        return nullptr;
    }
    c.h_pinned_bytes = want;
    return c.h_pinned;
}

int32_t merkle_init();  // merkle_host.cu
void merkle_shutdown();
int32_t bls_init();     // bls_host.cu
void bls_shutdown();

}  // namespace lhb200

using namespace lhb200;

extern "C" {

int32_t lhb200_init(int32_t device) {
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    if (c.ready) {
        if (c.device == device) return LHB200_OK;
        set_error("already initialised on device %d (one process per GPU)", c.device);
        return LHB200_EINVAL;
    }
    // Every verify call drives three streams of its own handle; with the default 8 hardware work queues the streams of
    // concurrent callers alias and serialise (8 gossip workers: 830 batches/s; with 32 queues: 1 350, and 2 140 from 16
    // workers — profiles/r2_gossip_concurrency.jsonl).  Read by the driver when the context is created, so it only takes
    // effect if this is the process's first CUDA call (a host that creates the context earlier sets it itself).
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device visible (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return LHB200_ENODEV;
    }
    if (device < 0 || device >= n) {
        set_error("device %d out of range (0..%d)", device, n - 1);
        return LHB200_EINVAL;
    }
    LHB_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    LHB_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        set_error("device %d is sm_%d%d; this build carries sm_100a code only", device, p.major, p.minor);
        return LHB200_ENODEV;
    }
    LHB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.device = device;
    int32_t rc = merkle_init();
    if (rc != LHB200_OK) return rc;
    rc = bls_init();
    if (rc != LHB200_OK) return rc;
    c.ready = true;
    return LHB200_OK;
}

void lhb200_shutdown(void) {
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    if (!c.ready) return;
    cudaStreamSynchronize(c.stream);
    bls_shutdown();     // idle batch handles (device buffers, streams, events of the old context)
    merkle_shutdown();  // the recycled state arena
    if (c.d_scratch) cudaFree(c.d_scratch);
    if (c.h_pinned) cudaFreeHost(c.h_pinned);
    c.d_scratch = nullptr; c.d_scratch_bytes = 0;
    c.h_pinned = nullptr; c.h_pinned_bytes = 0;
    cudaStreamDestroy(c.stream);
    c.stream = nullptr;
    c.ready = false;
    c.device = -1;
}

const char* lhb200_last_error(void) { return t_err; }

int32_t lhb200_pinned_alloc(void** out, uint64_t nbytes) {
    LHB_REQUIRE_READY();
    if (!out) return LHB200_EINVAL;
    LHB_CUDA(cudaHostAlloc(out, nbytes ? nbytes : 1, cudaHostAllocDefault));
    return LHB200_OK;
}
int32_t lhb200_pinned_free(void* p) {
    LHB_REQUIRE_READY();
    LHB_CUDA(cudaFreeHost(p));
    return LHB200_OK;
}
uint64_t lhb200_launch_count(void) { return ctx().launches.load(); }

}  // extern "C"
