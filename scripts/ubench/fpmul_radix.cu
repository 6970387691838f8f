// fpmul_radix.cu — experiment: Montgomery multiplication throughput, radix 2^32 carry chains (current fp_mul)
// versus radix 2^28 product scanning with carry-free 64-bit IMAD.WIDE accumulation.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../lighthouse_b200/csrc/bls/fp.cuh"
using namespace lhb200::bls;

// ---- radix 2^28, 14 limbs, R = 2^392 ----
#define NL28 14
#define MASK28 0x0fffffffu
__device__ __constant__ uint32_t P28[NL28];
__device__ __constant__ uint32_t PINV28;   // -p^-1 mod 2^28

__device__ __forceinline__ void mul28(uint32_t r[NL28], const uint32_t a[NL28], const uint32_t b[NL28]) {
    uint32_t m[NL28];
    unsigned long long acc = 0, acc2 = 0;
#pragma unroll
    for (int k = 0; k < NL28; k++) {
#pragma unroll
        for (int i = 0; i < k; i++) {
            acc += (unsigned long long)a[i] * b[k - i];
            acc2 += (unsigned long long)m[i] * P28[k - i];
        }
        acc += (unsigned long long)a[k] * b[0];
        acc += acc2; acc2 = 0;
        m[k] = ((uint32_t)acc * PINV28) & MASK28;
        acc += (unsigned long long)m[k] * P28[0];
        acc >>= 28;
    }
#pragma unroll
    for (int k = NL28; k < 2 * NL28 - 1; k++) {
#pragma unroll
        for (int i = k - NL28 + 1; i < NL28; i++) {
            acc += (unsigned long long)a[i] * b[k - i];
            acc2 += (unsigned long long)m[i] * P28[k - i];
        }
        acc += acc2; acc2 = 0;
        r[k - NL28] = (uint32_t)acc & MASK28;
        acc >>= 28;
    }
    r[NL28 - 1] = (uint32_t)acc;
}

#define CHAIN 256
__global__ void k_mul32(uint32_t* out, const uint32_t* in) {
    Fp x, y;
    for (int i = 0; i < 12; i++) { x.v[i] = in[i] + threadIdx.x; y.v[i] = in[12 + i]; }
    x.v[11] &= 0x0fffffff; y.v[11] &= 0x0fffffff;
    for (int it = 0; it < CHAIN; it++) { Fp t; fp_mul_inl(t, x, y); y = x; x = t; }
    uint32_t s = 0;
    for (int i = 0; i < 12; i++) s ^= x.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mul28(uint32_t* out, const uint32_t* in) {
    uint32_t x[NL28], y[NL28], t[NL28];
    for (int i = 0; i < NL28; i++) { x[i] = (in[i] + threadIdx.x) & MASK28; y[i] = in[14 + i] & MASK28; }
    for (int it = 0; it < CHAIN; it++) {
        mul28(t, x, y);
        for (int i = 0; i < NL28; i++) { y[i] = x[i]; x[i] = t[i]; }
    }
    uint32_t s = 0;
    for (int i = 0; i < NL28; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
static void run(const char* name, K kern, uint32_t* d_out, uint32_t* d_in, int blocks, int threads) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) kern<<<blocks, threads>>>(d_out, d_in);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int w = 0; w < reps; w++) kern<<<blocks, threads>>>(d_out, d_in);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double muls = (double)blocks * threads * CHAIN * reps;
    printf("{\"kernel\": \"%s\", \"blocks\": %d, \"threads\": %d, \"gmul_per_s\": %.3f, \"err\": \"%s\"}\n", name, blocks, threads,
           muls / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    // p in radix 2^28
    const char* hexp = "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab";
    unsigned char bytes[48];
    for (int i = 0; i < 48; i++) { unsigned v; sscanf(hexp + 2 * i, "%2x", &v); bytes[i] = (unsigned char)v; }
    uint32_t p28[NL28] = {0};
    for (int bit = 0; bit < 384; bit++) {
        int byte = 47 - bit / 8;
        if ((bytes[byte] >> (bit % 8)) & 1) p28[bit / 28] |= 1u << (bit % 28);
    }
    // -p^-1 mod 2^28 by Newton iteration
    uint32_t p0 = p28[0], inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - p0 * inv;
    uint32_t pinv = (0u - inv) & MASK28;
    cudaMemcpyToSymbol(P28, p28, sizeof p28);
    cudaMemcpyToSymbol(PINV28, &pinv, 4);
    uint32_t h_in[32];
    for (int i = 0; i < 32; i++) h_in[i] = 0x9e3779b9u * (i + 1);
    uint32_t *d_in, *d_out;
    cudaMalloc(&d_in, sizeof h_in);
    cudaMalloc(&d_out, 148 * 16 * 256 * 4);
    cudaMemcpy(d_in, h_in, sizeof h_in, cudaMemcpyHostToDevice);
    for (int bps : {2, 4, 8}) {
        run("radix32 carry-chain fp_mul", k_mul32, d_out, d_in, 148 * bps, 128);
        run("radix28 product-scanning", k_mul28, d_out, d_in, 148 * bps, 128);
    }
    return 0;
}
