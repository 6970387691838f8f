// icache_probe.cu — where is the instruction-cache cliff of an SM?  K distinct ~12.3 KB code bodies (straight-line
// dependent multiply / rotate-xor chains, no memory traffic; nvcc inlines them into one 300 KB kernel, each guarded by
// `ID < k`) are executed round-robin by every warp; the time per executed instruction is flat
// while K x 8 KB fits the instruction cache hierarchy and rises once the hot footprint falls out of it.  Written after
// round 1 found every code-growing BLS variant (lazy reduction, fused light leaves) losing in the full kernels although
// it won in isolation: the next round sizes the cooperative Miller loop's hot set against this curve.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 icache_probe.cu -o icache_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int BODY = 512;   // instructions per function (~8 KB of SASS): 256 x (IMAD, SHF+LOP3 fused or separate)
constexpr int KMAX = 24;

template <int ID>
__device__ __noinline__ uint32_t body(uint32_t x) {
#pragma unroll
    for (int i = 0; i < BODY / 2; i++)   // multiply + rotate-xor: not an affine map, so the chain cannot be folded
        x = (x * (2654435761u + 2u * (uint32_t)(ID * BODY + i))) ^ __funnelshift_r(x, x, 7 + (i & 15));
    return x;
}
template <int ID>
__device__ __forceinline__ uint32_t call_upto(uint32_t x, int k) {
    if constexpr (ID < KMAX) {
        if (ID < k) x = body<ID>(x);
        return call_upto<ID + 1>(x, k);
    } else {
        return x;
    }
}
__global__ void __launch_bounds__(64) k_probe(uint32_t* out, int k, int iters) {
    uint32_t x = threadIdx.x + 1;
    for (int it = 0; it < iters; it++) x = call_upto<0>(x, k);
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

int main() {
    uint32_t* d_out;
    cudaMalloc(&d_out, 148 * 8 * 64 * 4);
    for (int warps_per_sm : {8, 16}) {
        const int blocks = 148 * warps_per_sm / 2;
        for (int k = 1; k <= KMAX; k += (k < 8 ? 1 : 2)) {
            const int iters = 4096 / k;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            k_probe<<<blocks, 64>>>(d_out, k, iters);
            cudaEventRecord(e0);
            k_probe<<<blocks, 64>>>(d_out, k, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double insts = (double)iters * k * 788.0;                    // per warp: 788 SASS instructions per body (cuobjdump)
            const double cyc = ms * 1e-3 * 1.965e9;
            printf("{\"warps_per_sm\": %d, \"functions\": %d, \"footprint_kb\": %d, \"ms\": %.3f, \"cycles_per_inst_per_warp\": %.3f, \"err\": \"%s\"}\n",
                   warps_per_sm, k, (int)(k * 788 * 16 / 1024), ms, cyc / insts, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
