// coop.cuh — block-cooperative Fp12 arithmetic for the serial tail of a batch (the final exponentiation).
//
// One final exponentiation is ~8 k dependent Fp multiplications; on a single thread that is ~11 ms of pure
// latency (profiles/r1_launches_bls_100k.csv).  Here one CTA of 64 threads works on ONE Fp12 value held in shared
// memory as six Fp2 coefficients of w^k (Fp12 = Fp2[w]/(w^6 - xi)):
//     k:      0      1      2      3      4      5
//     tower:  c0.c0  c1.c0  c0.c1  c1.c1  c0.c2  c1.c2
//   * multiplication: 36 coefficient products on 36 threads, then 6 threads fold them (schoolbook over Fp2);
//   * cyclotomic squaring (Granger–Scott): 9 Fp2 squarings on 9 threads, 6 threads combine;
//   * Frobenius maps: one coefficient per thread.
// Every function below must be called by ALL threads of the block (they contain __syncthreads()).
#pragma once
#include "pairing.cuh"

namespace lhb200 {
namespace bls {

struct Fp12W {  // w-basis coefficient array
    Fp2 k[6];
};
struct CoopScratch {
    Fp2 prod[36];
    Fp2 aux[12];
};

__device__ __forceinline__ void coop_from_tower(Fp12W& r, const Fp12& a) {  // thread 0 only
    r.k[0] = a.c0.c0; r.k[1] = a.c1.c0; r.k[2] = a.c0.c1; r.k[3] = a.c1.c1; r.k[4] = a.c0.c2; r.k[5] = a.c1.c2;
}
__device__ __forceinline__ void coop_to_tower(Fp12& r, const Fp12W& a) {  // thread 0 only
    r.c0.c0 = a.k[0]; r.c1.c0 = a.k[1]; r.c0.c1 = a.k[2]; r.c1.c1 = a.k[3]; r.c0.c2 = a.k[4]; r.c1.c2 = a.k[5];
}

// r = a * b   (r may alias a or b)
__device__ __noinline__ void coop_mul(Fp12W& r, const Fp12W& a, const Fp12W& b, CoopScratch& s) {
    const int t = threadIdx.x;
    if (t < 36) fp2_mul(s.prod[t], a.k[t / 6], b.k[t % 6]);
    __syncthreads();
    if (t < 6) {
        // c_t = sum_{i+j=t} p_ij + xi * sum_{i+j=t+6} p_ij
        Fp2 lo, hi;
        bool have_lo = false, have_hi = false;
        for (int i = 0; i < 6; i++) {
            const int j = t - i;
            if (j >= 0 && j < 6) {
                if (have_lo) fp2_add(lo, lo, s.prod[6 * i + j]);
                else { lo = s.prod[6 * i + j]; have_lo = true; }
            }
            const int j2 = t + 6 - i;
            if (j2 >= 0 && j2 < 6) {
                if (have_hi) fp2_add(hi, hi, s.prod[6 * i + j2]);
                else { hi = s.prod[6 * i + j2]; have_hi = true; }
            }
        }
        if (have_hi) {
            fp2_mul_xi(hi, hi);
            fp2_add(lo, lo, hi);
        }
        s.aux[t] = lo;
    }
    __syncthreads();
    if (t < 6) r.k[t] = s.aux[t];
    __syncthreads();
}

// r = a^2 for a in the cyclotomic subgroup (same formulas as fp12_cyclotomic_sqr; pairs (a0,a3) (a1,a4) (a2,a5))
__device__ __noinline__ void coop_cyc_sqr(Fp12W& r, const Fp12W& a, CoopScratch& s) {
    const int t = threadIdx.x;
    if (t < 9) {
        const int p = t / 3, kind = t % 3;  // pair p: x = a_p, y = a_{p+3}; kind 0: x^2, 1: y^2, 2: (x+y)^2
        Fp2 v;
        if (kind == 0) v = a.k[p];
        else if (kind == 1) v = a.k[p + 3];
        else fp2_add(v, a.k[p], a.k[p + 3]);
        fp2_sqr(s.prod[t], v);
    }
    __syncthreads();
    if (t < 6) {
        // sq4(p) = (x^2 + xi y^2, (x+y)^2 - x^2 - y^2) =: (A_p, B_p)
        // a0' = 3 A_0 - 2 a0   a3' = 3 B_0 + 2 a3
        // a1' = 3 xi B_2 + 2 a1   a4' = 3 A_2 - 2 a4      (pair 2 = (a2, a5))
        // a2' = 3 A_1 - 2 a2   a5' = 3 B_1 + 2 a5         (pair 1 = (a1, a4))
        const int pair_of[6] = {0, 2, 1, 0, 2, 1};
        const bool wantB[6] = {false, true, false, true, false, true};
        const int p = pair_of[t];
        Fp2 v, u;
        if (!wantB[t]) {  // A_p
            fp2_mul_xi(u, s.prod[3 * p + 1]);
            fp2_add(v, s.prod[3 * p], u);
        } else {          // B_p
            fp2_sub(v, s.prod[3 * p + 2], s.prod[3 * p]);
            fp2_sub(v, v, s.prod[3 * p + 1]);
            if (t == 1) fp2_mul_xi(v, v);
        }
        // out = 3 v -+ 2 a_t   (minus for A-type, plus for B-type)
        if (!wantB[t]) fp2_sub(u, v, a.k[t]);
        else fp2_add(u, v, a.k[t]);
        fp2_dbl(u, u);
        fp2_add(u, u, v);
        s.aux[t] = u;
    }
    __syncthreads();
    if (t < 6) r.k[t] = s.aux[t];
    __syncthreads();
}

__device__ __noinline__ void coop_conj(Fp12W& r, const Fp12W& a) {  // a^(p^6): negate odd powers of w
    const int t = threadIdx.x;
    if (t < 6) {
        Fp2 v = a.k[t];
        if (t & 1) fp2_neg(v, v);
        r.k[t] = v;
    }
    __syncthreads();
}
__device__ __noinline__ void coop_frob(Fp12W& r, const Fp12W& a) {  // a^p
    const int t = threadIdx.x;
    if (t < 6) {
        Fp2 v;
        fp2_conj(v, a.k[t]);
        if (t) fp2_mul(v, v, FROB_G1[t]);
        r.k[t] = v;
    }
    __syncthreads();
}
__device__ __noinline__ void coop_frob2(Fp12W& r, const Fp12W& a) {  // a^(p^2)
    const int t = threadIdx.x;
    if (t < 6) {
        Fp2 v = a.k[t];
        if (t) fp2_mul_fp(v, v, FROB_G2[t]);
        r.k[t] = v;
    }
    __syncthreads();
}
__device__ __noinline__ void coop_copy(Fp12W& r, const Fp12W& a) {
    const int t = threadIdx.x;
    if (t < 6) r.k[t] = a.k[t];
    __syncthreads();
}

// r = a^x (x < 0): conj(a^|x|), a in the cyclotomic subgroup.  r must not alias a.
__device__ __noinline__ void coop_pow_x(Fp12W& r, const Fp12W& a, CoopScratch& s) {
    coop_copy(r, a);
    for (int i = 62; i >= 0; i--) {
        coop_cyc_sqr(r, r, s);
        if ((BLS_X_ABS >> i) & 1) coop_mul(r, r, a, s);
    }
    coop_conj(r, r);
}

// Shared-memory working set of the cooperative final exponentiation.
struct CoopFinalSmem {
    Fp12W f, t0, t1, t2;
    CoopScratch s;
    Fp12 tower;  // staging for the single-thread inversion
};

// f <- f^(3 (p^12-1)/r)  (same exponent as final_exp()).  All threads of the block.
__device__ __noinline__ void coop_final_exp(CoopFinalSmem& m) {
    // easy part: f^((p^6-1)(p^2+1)); the one Fp12 inversion stays on thread 0
    if (threadIdx.x == 0) {
        Fp12 a, inv;
        coop_to_tower(a, m.f);
        fp12_inv(inv, a);
        coop_from_tower(m.t1, inv);
    }
    __syncthreads();
    coop_conj(m.t0, m.f);
    coop_mul(m.f, m.t0, m.t1, m.s);
    coop_frob2(m.t0, m.f);
    coop_mul(m.f, m.t0, m.f, m.s);
    // hard part (see final_exp in pairing.cuh)
    coop_pow_x(m.t0, m.f, m.s);
    coop_conj(m.t1, m.f);
    coop_mul(m.t0, m.t0, m.t1, m.s);      // a = f^(x-1)
    coop_pow_x(m.t1, m.t0, m.s);
    coop_conj(m.t2, m.t0);
    coop_mul(m.t0, m.t1, m.t2, m.s);      // a = f^((x-1)^2)
    coop_pow_x(m.t1, m.t0, m.s);          // a^x
    coop_frob(m.t2, m.t0);                // a^p
    coop_mul(m.t0, m.t1, m.t2, m.s);      // b = a^(x+p)
    coop_pow_x(m.t1, m.t0, m.s);
    coop_pow_x(m.t2, m.t1, m.s);          // b^(x^2)
    coop_frob2(m.t1, m.t0);               // b^(p^2)
    coop_mul(m.t2, m.t2, m.t1, m.s);
    coop_conj(m.t1, m.t0);                // b^-1
    coop_mul(m.t2, m.t2, m.t1, m.s);      // c = b^(x^2+p^2-1)
    coop_cyc_sqr(m.t1, m.f, m.s);
    coop_mul(m.t1, m.t1, m.f, m.s);       // f^3
    coop_mul(m.f, m.t2, m.t1, m.s);
}

constexpr int COOP_THREADS = 64;

// the w-basis coefficient k of a tower element (see the table at the top of this file)
__device__ __forceinline__ const Fp2& coop_tower_coeff(const Fp12& a, int k) {
    return k == 0 ? a.c0.c0 : k == 1 ? a.c1.c0 : k == 2 ? a.c0.c1 : k == 3 ? a.c1.c1 : k == 4 ? a.c0.c2 : a.c1.c2;
}
// Largest tail of the Miller product tree that k_final_coop folds itself (a cooperative Fp12 product is ~6 us, one level
// of the single-thread tree ~350 us of latency).
constexpr uint32_t COOP_TAIL = 64;

// verdict = !fail && final_exp(prod[0] * ... * prod[n_prod-1] * f_last) == 1
__global__ void __launch_bounds__(COOP_THREADS) k_final_coop(const Fp12* __restrict__ prod, uint32_t n_prod,
                                                              const Fp12* __restrict__ f_last,
                                                              const uint32_t* __restrict__ fail, uint8_t* __restrict__ ok,
                                                              Fp12* __restrict__ gt_out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    CoopFinalSmem& m = *reinterpret_cast<CoopFinalSmem*>(smem_raw);
    if (*fail) {  // uniform across the block
        if (threadIdx.x == 0) *ok = 0;
        return;
    }
    const int t = threadIdx.x;
    if (t < 6) {   // f_last == nullptr: the (-g1, sum r sig) pair is already inside prod (cooperative Miller kernel)
        if (f_last) m.f.k[t] = coop_tower_coeff(*f_last, t);
        else { fp2_set_zero(m.f.k[t]); if (t == 0) m.f.k[0].c0 = FP_ONE; }
    }
    __syncthreads();
    for (uint32_t i = 0; i < n_prod; i++) {
        if (t < 6) m.t1.k[t] = coop_tower_coeff(prod[i], t);
        __syncthreads();
        coop_mul(m.f, m.f, m.t1, m.s);
    }
    coop_final_exp(m);
    if (threadIdx.x == 0) {
        Fp12 g;
        coop_to_tower(g, m.f);
        if (gt_out) *gt_out = g;
        *ok = fp12_is_one(g) ? 1 : 0;
    }
}

}  // namespace bls
}  // namespace lhb200
