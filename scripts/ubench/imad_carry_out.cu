// imad_carry_out.cu — can the Montgomery rows drop their carry CHAINS?
// Round 1 measured IMAD.WIDE.U32.X (carry-in + carry-out, what a mad.lo.cc/madc.hi.cc chain compiles to) at
// 0.99 warp-instructions/cycle/SM and plain IMAD.WIDE.U32 at 1.85.  This probe measures the form in between:
//   IMAD.WIDE.U32 R, P, a, b, R   (carry-OUT only)  +  IADD3.X cnt, cnt, RZ, RZ, P   (carry counted on the ALU pipe)
// i.e. every partial product is an independent instruction and the carries are tallied instead of chained.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_carry_out imad_carry_out.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048

// N independent (lo, hi, cnt) accumulators: mad + carry tally
template <int N>
__global__ void k_cout(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t lo[N], hi[N], cnt[N];
    for (int i = 0; i < N; i++) { lo[i] = threadIdx.x + i; hi[i] = i; cnt[i] = 0; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < N; i++)
            asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
                         : "+r"(lo[i]), "+r"(hi[i]), "+r"(cnt[i]) : "r"(a + i), "r"(b));
    }
    uint32_t s = 0;
    for (int i = 0; i < N; i++) s ^= lo[i] ^ hi[i] ^ cnt[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// chained reference: N/2-long carry chains (2 chains), as in the shipped rows
template <int N>
__global__ void k_chain(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t r[2][2 * N + 1];
    for (int c = 0; c < 2; c++) for (int i = 0; i <= 2 * N; i++) r[c][i] = threadIdx.x + i + c;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(r[c][0]), "+r"(r[c][1]) : "r"(a), "r"(b));
#pragma unroll
            for (int i = 1; i < N; i++)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(r[c][2 * i]), "+r"(r[c][2 * i + 1]) : "r"(a + i), "r"(b));
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(r[c][2 * N]));
        }
    }
    uint32_t s = 0;
    for (int c = 0; c < 2; c++) for (int i = 0; i <= 2 * N; i++) s ^= r[c][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
static void run(const char* name, K kern, double prod_per_iter, uint32_t* d, int threads, int blocks_per_sm) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) kern<<<blocks, threads>>>(d, 12345u, 67890u);
    cudaEventRecord(e0);
    const int reps = 4;
    for (int w = 0; w < reps; w++) kern<<<blocks, threads>>>(d, 12345u, 67890u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double prods = (double)blocks * threads * ITERS * prod_per_iter * reps;
    const double per_s = prods / (ms * 1e-3);
    printf("{\"kernel\": \"%s\", \"threads_per_sm\": %d, \"products_per_s\": %.4e, \"warp_products_per_cycle_per_sm\": %.4f, \"err\": \"%s\"}\n",
           name, threads * blocks_per_sm, per_s, per_s / 32 / sms / 1.965e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    uint32_t* d;
    cudaMalloc(&d, 148 * 16 * 1024 * 4);
    for (int occ : {1, 2, 4}) {   // 256, 512, 1024 threads per SM
        run("carry-out + tally, 12 independent accumulators", k_cout<12>, 12, d, 256, occ);
        run("carry-out + tally, 24 independent accumulators", k_cout<24>, 24, d, 256, occ);
        run("2 carry chains of 6 (shipped row shape)", k_chain<6>, 12, d, 256, occ);
        run("2 carry chains of 12", k_chain<12>, 24, d, 256, occ);
    }
    return 0;
}
