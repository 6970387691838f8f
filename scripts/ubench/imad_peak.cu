// imad_peak.cu — resident integer-multiply micro-benchmark: the measured denominators of the BLS ALU roofline
// ("peak measured by a resident IMAD micro-benchmark on the same box").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_peak imad_peak.cu ; run on the B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
// 8 independent carry chains of mad.lo.cc/madc.hi.cc pairs (fused by ptxas to IMAD.WIDE.U32.X), like fp_mul's rows
__global__ void k_wide_x(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t r[16];
    for (int i = 0; i < 16; i++) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                         "madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
                         "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
                         "madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
                         : "+r"(r[8 * c + 0]), "+r"(r[8 * c + 1]), "+r"(r[8 * c + 2]), "+r"(r[8 * c + 3]), "+r"(r[8 * c + 4]),
                           "+r"(r[8 * c + 5]), "+r"(r[8 * c + 6]), "+r"(r[8 * c + 7])
                         : "r"(a), "r"(b));
        }
    }
    uint32_t s = 0;
    for (int i = 0; i < 16; i++) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// independent 32x32+64 multiply-adds without carry flags (IMAD.WIDE.U32)
__global__ void k_wide(uint32_t* out, uint32_t a, uint32_t b) {
    unsigned long long r[8];
    for (int i = 0; i < 8; i++) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < 8; c++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r[c]) : "r"(a + c), "r"(b));
    }
    unsigned long long s = 0;
    for (int i = 0; i < 8; i++) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}
// plain 32-bit IMAD (low half)
__global__ void k_lo(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t r[8];
    for (int i = 0; i < 8; i++) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < 8; c++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[c]) : "r"(a + c), "r"(b));
    }
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_hi(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t r[8];
    for (int i = 0; i < 8; i++) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < 8; c++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(r[c]) : "r"(a + c), "r"(b));
    }
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
static void run(const char* name, K kern, int ops_per_iter, uint32_t* d) {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) kern<<<blocks, threads>>>(d, 12345u, 67890u);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int w = 0; w < reps; w++) kern<<<blocks, threads>>>(d, 12345u, 67890u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)blocks * threads * ITERS * ops_per_iter * reps;   // thread-level multiply instructions
    const double per_s = inst / (ms * 1e-3);
    printf("{\"kernel\": \"%s\", \"thread_inst_per_s\": %.4e, \"warp_inst_per_cycle_per_sm\": %.4f, \"sm_clock_mhz_nominal\": %d}\n",
           name, per_s, per_s / 32.0 / sms / (clk * 1e3), clk / 1000);
}

int main() {
    uint32_t* d;
    cudaMalloc(&d, 148 * 8 * 256 * 4 * 4);
    run("IMAD.WIDE.U32.X carry chains (16 per iter)", k_wide_x, 16 / 2 * 1, d);   // 8 fused wide ops per 16 mad halves
    run("IMAD.WIDE.U32 independent", k_wide, 8, d);
    run("IMAD (32-bit lo)", k_lo, 8, d);
    run("IMAD.HI", k_hi, 8, d);
    return 0;
}
