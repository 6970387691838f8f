// fp2_leaf.cu — how close do the out-of-line Fp2 leaves and the tower operations built on them get to the
// IMAD.WIDE roofline at the occupancy the BLS kernels actually run at (4 CTAs x 64 threads per SM)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -rdc=true fp2_leaf.cu ../../lighthouse_b200/csrc/bls/fp_core.cu -o fp2_leaf
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define LHB_FP_DECL_ONLY 1
#include "../../lighthouse_b200/csrc/bls/fp2.cuh"
using namespace lhb200::bls;

__device__ void init(Fp2* v, int n, const uint32_t* in) {
    for (int k = 0; k < n; k++)
        for (int i = 0; i < 12; i++) {
            v[k].c0.v[i] = in[(i + k) % 24] + threadIdx.x;
            v[k].c1.v[i] = in[(i + 2 * k + 5) % 24] ^ threadIdx.x;
        }
    for (int k = 0; k < n; k++) { v[k].c0.v[11] &= 0x0fffffff; v[k].c1.v[11] &= 0x0fffffff; }
}
__device__ uint32_t fold(const Fp2* v, int n) {
    uint32_t s = 0;
    for (int k = 0; k < n; k++)
        for (int i = 0; i < 12; i++) s ^= v[k].c0.v[i] ^ v[k].c1.v[i];
    return s;
}

template <int MODE>
__global__ void __launch_bounds__(64) k_bench(uint32_t* out, const uint32_t* in, int iters) {
    Fp12 f, g;
    Fp2* fv = reinterpret_cast<Fp2*>(&f);
    Fp2* gv = reinterpret_cast<Fp2*>(&g);
    init(fv, 6, in);
    init(gv, 6, in + 3);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { fp2_mul(fv[0], fv[1], fv[2]); fp2_mul(fv[1], fv[2], fv[0]); fp2_mul(fv[2], fv[0], fv[1]); }
        if (MODE == 1) { fp2_sqr(fv[0], fv[1]); fp2_sqr(fv[1], fv[2]); fp2_sqr(fv[2], fv[0]); }
        if (MODE == 2) { fp6_mul(f.c0, f.c1, g.c0); fp6_mul(f.c1, f.c0, g.c1); }
        if (MODE == 3) fp12_sqr(f, f);
        if (MODE == 4) fp12_mul_by_014(f, f, gv[0], gv[1], gv[2]);
        if (MODE == 5) { fp2_add(fv[0], fv[1], fv[2]); fp2_sub(fv[1], fv[2], fv[0]); fp2_add(fv[2], fv[0], fv[1]); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = fold(fv, 6);
}

template <class K>
static void run(const char* name, K kern, double fpmul_per_iter, uint32_t* d_out, uint32_t* d_in, int ctas_per_sm, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * ctas_per_sm;
    kern<<<blocks, 64>>>(d_out, d_in, iters);
    cudaEventRecord(e0);
    kern<<<blocks, 64>>>(d_out, d_in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double muls = (double)blocks * 64 * iters * fpmul_per_iter;
    printf("{\"op\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"g_fpmul_per_s\": %.2f, \"frac_of_30.5\": %.3f, \"err\": \"%s\"}\n", name,
           ctas_per_sm, ms, fpmul_per_iter > 0 ? muls / (ms * 1e-3) / 1e9 : 0.0, fpmul_per_iter > 0 ? muls / (ms * 1e-3) / 30.5e9 : 0.0,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    uint32_t h_in[32];
    for (int i = 0; i < 32; i++) h_in[i] = 0x9e3779b9u * (i + 1);
    uint32_t *d_in, *d_out;
    cudaMalloc(&d_in, sizeof h_in);
    cudaMalloc(&d_out, 148 * 16 * 64 * 4);
    cudaMemcpy(d_in, h_in, sizeof h_in, cudaMemcpyHostToDevice);
    for (int c : {4, 8}) {
        run("fp2_mul x3", k_bench<0>, 9, d_out, d_in, c, 2000);
        run("fp2_sqr x3", k_bench<1>, 6, d_out, d_in, c, 2000);
        run("fp6_mul x2", k_bench<2>, 36, d_out, d_in, c, 600);
        run("fp12_sqr", k_bench<3>, 36, d_out, d_in, c, 600);
        run("fp12_mul_by_014", k_bench<4>, 39, d_out, d_in, c, 600);
        run("fp2_add/sub x3 (ms only)", k_bench<5>, 0, d_out, d_in, c, 2000);
    }
    return 0;
}
