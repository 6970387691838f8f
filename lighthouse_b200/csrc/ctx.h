// ctx.h — process-wide library context: one process drives one GPU (one rank per GPU).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/lhb200.h"

namespace lhb200 {

struct Ctx {
    bool ready = false;
    int device = -1;
    cudaStream_t stream = nullptr;
    std::recursive_mutex mu;  // host entry points are serialised (callers are spawn_blocking workers)
    std::atomic<uint64_t> launches{0};
    uint8_t zero_hashes[65][32];
    // scratch for host-buffer entry points
    void* d_scratch = nullptr;
    size_t d_scratch_bytes = 0;
    void* h_pinned = nullptr;
    size_t h_pinned_bytes = 0;
};

Ctx& ctx();
void set_error(const char* fmt, ...);
int32_t cuda_fail(cudaError_t e, const char* what);
// grow-only scratch (device / pinned host); returns nullptr on failure (error set)
void* dev_scratch(size_t nbytes);
void* pinned_scratch(size_t nbytes);

#define LHB_CUDA(expr)                                                 \
    do {                                                               \
        cudaError_t _e = (expr);                                       \
        if (_e != cudaSuccess) return ::lhb200::cuda_fail(_e, #expr);  \
    } while (0)

#define LHB_REQUIRE_READY()                                                          \
    do {                                                                             \
        if (!::lhb200::ctx().ready) {                                                \
            ::lhb200::set_error("lhb200_init() has not succeeded (no CPU fallback)"); \
            return LHB200_ENODEV;                                                    \
        }                                                                            \
    } while (0)

// the library's NCCL communicator (comm.cu); no-ops / local copies when no communicator is initialised
bool comm_active();
int comm_world();
int comm_rank();
int32_t comm_allreduce_min_u8(void* d_buf, size_t n, cudaStream_t s);
int32_t comm_allgather_bytes(const void* d_send, void* d_recv, size_t bytes_per_rank, cudaStream_t s);

inline void count_launch(uint64_t n = 1) { ctx().launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace lhb200
