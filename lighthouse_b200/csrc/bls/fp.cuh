// fp.cuh — BLS12-381 base field Fp, 12 x 32-bit limbs, Montgomery form (R = 2^384), for sm_100a.
//
// This is the arithmetic Lighthouse gets from blst (crate not vendored; call sites
// /root/reference/crypto/bls/src/impls/blst.rs:11-12,75,103,114,139,193).  It is NOT a port of blst: blst is
// 6x64-bit mulx/adx assembly; here a product row is a carry chain of 32x32+64 multiply-adds
// (mad.lo.cc/madc.hi.cc pairs, which ptxas fuses into one IMAD.WIDE.U32.X each), with the even- and
// odd-indexed partial products kept in two accumulators so a row needs no carry fix-ups.  A full Montgomery
// multiplication is ~300 IMAD.WIDE + ~45 IADD3 (see profiles/ for the SASS census).
//
// The same source compiles for the host (LHB_HOSTSIM, carries emulated in C) so the tests can exercise the
// exact limb algorithms without a GPU; the product library never contains or calls that build.
#pragma once
#include <stdint.h>

#ifdef LHB_HOSTSIM
#define LHB_HD
#define LHB_NOINLINE __attribute__((noinline))
#define LHB_INLINE inline
#define LHB_CONST static const
#else
#define LHB_HD __device__
#ifndef LHB_NOINLINE
#define LHB_NOINLINE __noinline__
#endif
#define LHB_INLINE __forceinline__
#define LHB_CONST static __device__ __constant__ const
#endif

namespace lhb200 {
namespace bls {

constexpr int NL = 12;  // limbs

struct alignas(16) Fp {
    uint32_t v[NL];
};
struct Fp2 {
    Fp c0, c1;
};

#include "consts.inc"

// ------------------------------------------------------------------------------------------------ limb primitives
#if defined(__CUDA_ARCH__)
#define LHB_DEV 1
__device__ LHB_INLINE void mul_pair(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
__device__ LHB_INLINE void mad_pair_first(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
__device__ LHB_INLINE void mad_pair(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
__device__ LHB_INLINE void mad_pair_sh(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b, uint32_t slo, uint32_t shi) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                 : "=r"(dlo), "=r"(dhi) : "r"(a), "r"(b), "r"(slo), "r"(shi));
}
__device__ LHB_INLINE void mad_pair_last(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;" : "=r"(dlo), "=r"(dhi) : "r"(a), "r"(b));
}
__device__ LHB_INLINE void add_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
__device__ LHB_INLINE void addc_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
__device__ LHB_INLINE void addc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
__device__ LHB_INLINE void sub_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
__device__ LHB_INLINE void subc_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
__device__ LHB_INLINE void subc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
#else
// Host emulation of the PTX carry flag (tests only).
static thread_local uint32_t g_cf = 0;
static inline void emu_add(uint32_t& r, uint32_t a, uint32_t b, uint32_t cin, bool setc) {
    uint64_t s = (uint64_t)a + b + cin;
    r = (uint32_t)s;
    if (setc) g_cf = (uint32_t)(s >> 32);
}
static inline void mul_pair(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    uint64_t p = (uint64_t)a * b; lo = (uint32_t)p; hi = (uint32_t)(p >> 32);
}
static inline void mad_pair_first(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    uint64_t p = (uint64_t)a * b; emu_add(lo, lo, (uint32_t)p, 0, true); emu_add(hi, hi, (uint32_t)(p >> 32), g_cf, true);
}
static inline void mad_pair(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    uint64_t p = (uint64_t)a * b; emu_add(lo, lo, (uint32_t)p, g_cf, true); emu_add(hi, hi, (uint32_t)(p >> 32), g_cf, true);
}
static inline void mad_pair_sh(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b, uint32_t slo, uint32_t shi) {
    uint64_t p = (uint64_t)a * b; emu_add(dlo, slo, (uint32_t)p, g_cf, true); emu_add(dhi, shi, (uint32_t)(p >> 32), g_cf, true);
}
static inline void mad_pair_last(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b) {
    uint64_t p = (uint64_t)a * b; emu_add(dlo, 0, (uint32_t)p, g_cf, true); emu_add(dhi, 0, (uint32_t)(p >> 32), g_cf, false);
}
static inline void add_cc(uint32_t& r, uint32_t a, uint32_t b) { emu_add(r, a, b, 0, true); }
static inline void addc_cc(uint32_t& r, uint32_t a, uint32_t b) { emu_add(r, a, b, g_cf, true); }
static inline void addc(uint32_t& r, uint32_t a, uint32_t b) { emu_add(r, a, b, g_cf, false); }
// PTX sub.cc: CF = borrow
static inline void emu_sub(uint32_t& r, uint32_t a, uint32_t b, uint32_t bin, bool setc) {
    uint64_t s = (uint64_t)a - b - bin;
    r = (uint32_t)s;
    if (setc) g_cf = (uint32_t)((s >> 32) & 1);
}
static inline void sub_cc(uint32_t& r, uint32_t a, uint32_t b) { emu_sub(r, a, b, 0, true); }
static inline void subc_cc(uint32_t& r, uint32_t a, uint32_t b) { emu_sub(r, a, b, g_cf, true); }
static inline void subc(uint32_t& r, uint32_t a, uint32_t b) { emu_sub(r, a, b, g_cf, false); }
#endif

// ------------------------------------------------------------------------------------------------ basic ops
LHB_HD LHB_INLINE bool fp_is_zero(const Fp& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < NL; i++) o |= a.v[i];
    return o == 0;
}
LHB_HD LHB_INLINE bool fp_eq(const Fp& a, const Fp& b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < NL; i++) o |= a.v[i] ^ b.v[i];
    return o == 0;
}
LHB_HD LHB_INLINE void fp_set_zero(Fp& a) {
#pragma unroll
    for (int i = 0; i < NL; i++) a.v[i] = 0;
}
LHB_HD LHB_INLINE void fp_cmov(Fp& r, const Fp& a, bool c) {  // r = c ? a : r
#pragma unroll
    for (int i = 0; i < NL; i++) r.v[i] = c ? a.v[i] : r.v[i];
}

// r = (t >= p) ? t - p : t      (t < 2p, may carry a 13th bit in `top`)
LHB_HD LHB_INLINE void fp_final_sub(Fp& r, const uint32_t t[NL], uint32_t top) {
    uint32_t s[NL], brw;
    sub_cc(s[0], t[0], FP_P.v[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) subc_cc(s[i], t[i], FP_P.v[i]);
    subc(brw, top, 0);  // brw = top - borrow: 0xffffffff iff t < p
    const bool keep = (brw >> 31) != 0;
#pragma unroll
    for (int i = 0; i < NL; i++) r.v[i] = keep ? t[i] : s[i];
}

LHB_HD LHB_INLINE void fp_add_inl(Fp& r, const Fp& a, const Fp& b) {
    uint32_t t[NL], top;
    add_cc(t[0], a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) addc_cc(t[i], a.v[i], b.v[i]);
    addc(top, 0, 0);
    fp_final_sub(r, t, top);
}

LHB_HD LHB_INLINE void fp_sub_inl(Fp& r, const Fp& a, const Fp& b) {
    uint32_t t[NL], m;
    sub_cc(t[0], a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) subc_cc(t[i], a.v[i], b.v[i]);
    subc(m, 0, 0);  // 0xffffffff on borrow
    uint32_t q[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) q[i] = FP_P.v[i] & m;
    add_cc(r.v[0], t[0], q[0]);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) addc_cc(r.v[i], t[i], q[i]);
    addc(r.v[NL - 1], t[NL - 1], q[NL - 1]);
}
#ifdef LHB_FP_DECL_ONLY
LHB_HD void fp_add(Fp& r, const Fp& a, const Fp& b);
LHB_HD void fp_sub(Fp& r, const Fp& a, const Fp& b);
#else
LHB_HD LHB_NOINLINE void fp_add(Fp& r, const Fp& a, const Fp& b) { Fp x = a, y = b, o; fp_add_inl(o, x, y); r = o; }
LHB_HD LHB_NOINLINE void fp_sub(Fp& r, const Fp& a, const Fp& b) { Fp x = a, y = b, o; fp_sub_inl(o, x, y); r = o; }
#endif

LHB_HD LHB_INLINE void fp_neg(Fp& r, const Fp& a) {
    const bool z = fp_is_zero(a);
    uint32_t t[NL];
    sub_cc(t[0], FP_P.v[0], a.v[0]);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) subc_cc(t[i], FP_P.v[i], a.v[i]);
    subc(t[NL - 1], FP_P.v[NL - 1], a.v[NL - 1]);
#pragma unroll
    for (int i = 0; i < NL; i++) r.v[i] = z ? 0u : t[i];
}
LHB_HD LHB_INLINE void fp_dbl(Fp& r, const Fp& a) { fp_add(r, a, a); }

// ------------------------------------------------------------------------------------------------ Montgomery mul
// One row: T += a*bi; m = T[0]*M0; T += m*p; T >>= 32, with T = even + odd*2^32 (roles swap every row).
LHB_HD LHB_INLINE void mont_row(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi, bool first) {
    if (first) {
#pragma unroll
        for (int j = 0; j < NL; j += 2) {
            mul_pair(odd[j], odd[j + 1], a[j + 1], bi);
            mul_pair(even[j], even[j + 1], a[j], bi);
        }
    } else {
        // shift-in of the previous row (its low limb is zero) fused with this row's odd-index products
        add_cc(even[0], even[0], odd[1]);
#pragma unroll
        for (int j = 0; j < NL - 2; j += 2) mad_pair_sh(odd[j], odd[j + 1], a[j + 1], bi, odd[j + 2], odd[j + 3]);
        mad_pair_last(odd[NL - 2], odd[NL - 1], a[NL - 1], bi);
        mad_pair_first(even[0], even[1], a[0], bi);
#pragma unroll
        for (int j = 2; j < NL; j += 2) mad_pair(even[j], even[j + 1], a[j], bi);
        addc(odd[NL - 1], odd[NL - 1], 0);
    }
    const uint32_t mi = even[0] * LHB_FP_M0;
    mad_pair_first(odd[0], odd[1], FP_P.v[1], mi);
#pragma unroll
    for (int j = 2; j < NL; j += 2) mad_pair(odd[j], odd[j + 1], FP_P.v[j + 1], mi);
    mad_pair_first(even[0], even[1], FP_P.v[0], mi);
#pragma unroll
    for (int j = 2; j < NL; j += 2) mad_pair(even[j], even[j + 1], FP_P.v[j], mi);
    addc(odd[NL - 1], odd[NL - 1], 0);
}

LHB_HD LHB_INLINE void fp_mul_inl(Fp& r, const Fp& a, const Fp& b) {
    uint32_t even[NL], odd[NL];
#pragma unroll
    for (int i = 0; i < NL; i += 2) {
        mont_row(even, odd, a.v, b.v[i], i == 0);
        mont_row(odd, even, a.v, b.v[i + 1], false);
    }
    // T/2^32 of the last row: even = even + (odd >> 32)
    add_cc(even[0], even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) addc_cc(even[i], even[i], odd[i + 1]);
    addc(even[NL - 1], even[NL - 1], 0);
    fp_final_sub(r, even, 0);
}

// ---- split form: full 768-bit product and separate Montgomery reduction (lazy reduction in Fp2, fp2.cuh) ----
// Same even/odd carry-chain rows as mont_row, but the product rows carry no reduction (the low limb of every row is
// an output limb) and the reduction rows carry no product.  fp_redc_inl(fp_mulw_inl(a, b)) == fp_mul_inl(a, b).
LHB_HD LHB_INLINE void mulw_row(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi, bool first, uint32_t& out) {
    if (first) {
#pragma unroll
        for (int j = 0; j < NL; j += 2) {
            mul_pair(odd[j], odd[j + 1], a[j + 1], bi);
            mul_pair(even[j], even[j + 1], a[j], bi);
        }
    } else {
        add_cc(even[0], even[0], odd[1]);   // the previous row's low limb has been emitted: shift by one limb
#pragma unroll
        for (int j = 0; j < NL - 2; j += 2) mad_pair_sh(odd[j], odd[j + 1], a[j + 1], bi, odd[j + 2], odd[j + 3]);
        mad_pair_last(odd[NL - 2], odd[NL - 1], a[NL - 1], bi);
        mad_pair_first(even[0], even[1], a[0], bi);
#pragma unroll
        for (int j = 2; j < NL; j += 2) mad_pair(even[j], even[j + 1], a[j], bi);
        addc(odd[NL - 1], odd[NL - 1], 0);
    }
    out = even[0];
}
// w[0..23] = a * b   (a, b < 2^384; no reduction)
LHB_HD LHB_INLINE void fp_mulw_inl(uint32_t* w, const Fp& a, const Fp& b) {
    uint32_t even[NL], odd[NL];
#pragma unroll
    for (int i = 0; i < NL; i += 2) {
        mulw_row(even, odd, a.v, b.v[i], i == 0, w[i]);
        mulw_row(odd, even, a.v, b.v[i + 1], false, w[i + 1]);
    }
    add_cc(w[NL], even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) addc_cc(w[NL + i], even[i], odd[i + 1]);
    addc(w[2 * NL - 1], even[NL - 1], 0);
}
// one reduction row: m = T[0] * M0; T += m * p; (shift of the previous row fused, T[0] becomes 0)
LHB_HD LHB_INLINE void redc_row(uint32_t* even, uint32_t* odd, bool first) {
    if (first) {
        const uint32_t mi = even[0] * LHB_FP_M0;
        mad_pair_first(odd[0], odd[1], FP_P.v[1], mi);
#pragma unroll
        for (int j = 2; j < NL; j += 2) mad_pair(odd[j], odd[j + 1], FP_P.v[j + 1], mi);
        mad_pair_first(even[0], even[1], FP_P.v[0], mi);
#pragma unroll
        for (int j = 2; j < NL; j += 2) mad_pair(even[j], even[j + 1], FP_P.v[j], mi);
        addc(odd[NL - 1], odd[NL - 1], 0);
    } else {
        add_cc(even[0], even[0], odd[1]);
        const uint32_t mi = even[0] * LHB_FP_M0;   // mul.lo leaves the carry flag alone
#pragma unroll
        for (int j = 0; j < NL - 2; j += 2) mad_pair_sh(odd[j], odd[j + 1], FP_P.v[j + 1], mi, odd[j + 2], odd[j + 3]);
        mad_pair_last(odd[NL - 2], odd[NL - 1], FP_P.v[NL - 1], mi);
        mad_pair_first(even[0], even[1], FP_P.v[0], mi);
#pragma unroll
        for (int j = 2; j < NL; j += 2) mad_pair(even[j], even[j + 1], FP_P.v[j], mi);
        addc(odd[NL - 1], odd[NL - 1], 0);
    }
}
// r = w / 2^384 mod p   for w < p * 2^384
LHB_HD LHB_INLINE void fp_redc_inl(Fp& r, const uint32_t* w) {
    uint32_t even[NL], odd[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) { even[i] = w[i]; odd[i] = 0; }
#pragma unroll
    for (int i = 0; i < NL; i += 2) {
        redc_row(even, odd, i == 0);
        redc_row(odd, even, false);
    }
    // U = (w_low + M p) / 2^384 <= p, then + w_high (< p): below 2p, 13th bit in `top`
    uint32_t top;
    add_cc(even[0], even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) addc_cc(even[i], even[i], odd[i + 1]);
    addc(even[NL - 1], even[NL - 1], 0);
    add_cc(even[0], even[0], w[NL]);
#pragma unroll
    for (int i = 1; i < NL; i++) addc_cc(even[i], even[i], w[NL + i]);
    addc(top, 0, 0);
    fp_final_sub(r, even, top);
}

// w[0..23] = a^2: the 66 cross products a_i a_j (i < j) once, doubled, plus the 12 squares — 78 multiply
// instructions instead of 144.  Cross products whose pair starts at an even limb go to `e`, at an odd limb to `o`,
// so every row is two carry chains over ascending, disjoint limb pairs; the carry out of a chain lands in a limb no
// earlier row has touched (rows ascend), so there is never a ripple.
LHB_HD LHB_INLINE void fp_sqrw_inl(uint32_t* w, const Fp& a) {
    uint32_t e[2 * NL + 2], o[2 * NL + 2];
#pragma unroll
    for (int i = 0; i < 2 * NL + 2; i++) { e[i] = 0; o[i] = 0; }
#pragma unroll
    for (int i = 0; i < NL - 1; i++) {
        // products a_i a_j at limb i + j: chain A takes j = i+1, i+3, ... ; chain B takes j = i+2, i+4, ...
#pragma unroll
        for (int c = 1; c <= 2; c++) {
            uint32_t* acc = ((2 * i + c) & 1) ? o : e;   // parity of the first pair's limb index
            bool first = true;
            int last = -1;
#pragma unroll
            for (int j = i + c; j < NL; j += 2) {
                if (first) mad_pair_first(acc[i + j], acc[i + j + 1], a.v[i], a.v[j]);
                else mad_pair(acc[i + j], acc[i + j + 1], a.v[i], a.v[j]);
                first = false;
                last = i + j;
            }
            if (last >= 0) addc(acc[last + 2], acc[last + 2], 0);
        }
    }
    // cross = e + o, doubled; then the squares on the even-aligned pairs in one chain
    add_cc(w[0], e[0], o[0]);
#pragma unroll
    for (int i = 1; i < 2 * NL - 1; i++) addc_cc(w[i], e[i], o[i]);
    addc(w[2 * NL - 1], e[2 * NL - 1], o[2 * NL - 1]);
#pragma unroll
    for (int i = 2 * NL - 1; i > 0; i--) w[i] = (w[i] << 1) | (w[i - 1] >> 31);
    w[0] <<= 1;
    mad_pair_first(w[0], w[1], a.v[0], a.v[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) mad_pair(w[2 * i], w[2 * i + 1], a.v[i], a.v[i]);
}
LHB_HD LHB_INLINE void fp_sqr_inl(Fp& r, const Fp& a) {
    uint32_t w[2 * NL];
    fp_sqrw_inl(w, a);
    fp_redc_inl(r, w);
}

#ifdef LHB_FP_DECL_ONLY
LHB_HD void fp_mul(Fp& r, const Fp& a, const Fp& b);
#else
LHB_HD LHB_NOINLINE void fp_mul(Fp& r, const Fp& a, const Fp& b) { Fp x = a, y = b, o; fp_mul_inl(o, x, y); r = o; }
#endif
// EXPERIMENT (-DLHB_FP_DEDICATED_SQR, not shipped — DESIGN.md §9): a dedicated squaring leaf (fp_sqrw_inl + the split
// reduction) has 208 multiply instructions instead of 276, but its short, strictly ordered carry chains (rows of
// 11..1 cross products, the doubling, one 24-limb chain of squares) give ptxas much less to interleave than the fused
// product rows: a 100 k-set verify went from 91.0 to 94.4 ms with it.
#ifdef LHB_FP_DEDICATED_SQR
#ifdef LHB_FP_DECL_ONLY
LHB_HD void fp_sqr(Fp& r, const Fp& a);
#else
LHB_HD LHB_NOINLINE void fp_sqr(Fp& r, const Fp& a) { Fp x = a, o; fp_sqr_inl(o, x); r = o; }
#endif
#else
LHB_HD LHB_INLINE void fp_sqr(Fp& r, const Fp& a) { fp_mul(r, a, a); }
#endif

LHB_HD LHB_INLINE void fp_to_mont(Fp& r, const Fp& a) { fp_mul(r, a, FP_R2); }
LHB_HD LHB_INLINE void fp_from_mont(Fp& r, const Fp& a) {
    Fp one;
    fp_set_zero(one);
    one.v[0] = 1;
    fp_mul(r, a, one);
}

// r = a^((p-3)/4): 4-bit fixed window (15-entry table, 380 squarings + ~95 multiplications).
// Serves sqrt (a * r), inverse (r^4 * a) and the quadratic-residue test (a * r^2 = a^((p-1)/2)).
#ifdef LHB_FP_DECL_ONLY
LHB_HD void fp_pow_pm3d4(Fp& r, const Fp& a);
#else
LHB_HD LHB_NOINLINE void fp_pow_pm3d4(Fp& r, const Fp& a) {
    Fp tab[16];
    tab[1] = a;
    fp_sqr(tab[2], a);
    for (int i = 3; i < 16; i++) fp_mul(tab[i], tab[i - 1], a);
    constexpr int nbits = LHB_EXP_PM3D4_BITS;           // 379
    constexpr int nwin = (nbits + 3) / 4;                // 95
    bool started = false;
    Fp acc;
    for (int w = nwin - 1; w >= 0; w--) {
        const int bit = 4 * w;
        const uint32_t d = (FP_EXP_PM3D4[bit >> 5] >> (bit & 31)) & 15u;
        if (started) {
            fp_sqr(acc, acc); fp_sqr(acc, acc); fp_sqr(acc, acc); fp_sqr(acc, acc);
            if (d) fp_mul(acc, acc, tab[d]);
        } else if (d) {
            acc = tab[d];
            started = true;
        }
    }
    r = acc;
}
#endif

// r = 1/a (a != 0); 0 -> 0
LHB_HD LHB_INLINE void fp_inv(Fp& r, const Fp& a) {
    Fp t;
    fp_pow_pm3d4(t, a);
    fp_sqr(t, t);
    fp_sqr(t, t);
    fp_mul(r, t, a);
}

// r = 1/a by the binary extended Euclidean algorithm (HAC 14.61) — VARIABLE TIME, for public values only (the final
// exponentiation's one inversion: ~50 us on a lone thread against ~500 us for the 475-multiplication Fermat chain above).
// a in Montgomery form (a R); the loop inverts the integer a R, one Montgomery product by R^3 restores the form.
LHB_HD LHB_INLINE bool fp_limbs_ge(const uint32_t* a, const uint32_t* b) {   // a >= b
    for (int i = NL - 1; i >= 0; i--) {
        if (a[i] != b[i]) return a[i] > b[i];
    }
    return true;
}
LHB_HD LHB_INLINE void fp_limbs_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {   // r = a - b (a >= b)
    sub_cc(r[0], a[0], b[0]);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) subc_cc(r[i], a[i], b[i]);
    subc(r[NL - 1], a[NL - 1], b[NL - 1]);
}
LHB_HD LHB_INLINE void fp_limbs_half_mod(uint32_t* x) {   // x = x / 2 mod p (x < p)
    uint32_t top = 0;
    if (x[0] & 1u) {
        add_cc(x[0], x[0], FP_P.v[0]);
#pragma unroll
        for (int i = 1; i < NL; i++) addc_cc(x[i], x[i], FP_P.v[i]);
        addc(top, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < NL - 1; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
    x[NL - 1] = (x[NL - 1] >> 1) | (top << 31);
}
#ifndef LHB_FP_CORE_ONLY   // (compiled in the callers' translation unit only)
LHB_HD LHB_NOINLINE void fp_inv_vartime(Fp& r, const Fp& a) {
    if (fp_is_zero(a)) { fp_set_zero(r); return; }
    uint32_t u[NL], v[NL], x1[NL], x2[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) { u[i] = a.v[i]; v[i] = FP_P.v[i]; x1[i] = 0; x2[i] = 0; }
    x1[0] = 1;
    auto is_one = [](const uint32_t* t) {
        uint32_t o = t[0] ^ 1u;
        for (int i = 1; i < NL; i++) o |= t[i];
        return o == 0;
    };
    for (int guard = 0; guard < 2 * 384 + 8; guard++) {
        if (is_one(u) || is_one(v)) break;
        while (!(u[0] & 1u)) {
            for (int i = 0; i < NL - 1; i++) u[i] = (u[i] >> 1) | (u[i + 1] << 31);
            u[NL - 1] >>= 1;
            fp_limbs_half_mod(x1);
        }
        while (!(v[0] & 1u)) {
            for (int i = 0; i < NL - 1; i++) v[i] = (v[i] >> 1) | (v[i + 1] << 31);
            v[NL - 1] >>= 1;
            fp_limbs_half_mod(x2);
        }
        if (fp_limbs_ge(u, v)) {
            fp_limbs_sub(u, u, v);
            if (fp_limbs_ge(x1, x2)) fp_limbs_sub(x1, x1, x2);
            else { uint32_t t[NL]; fp_limbs_sub(t, x2, x1); fp_limbs_sub(x1, FP_P.v, t); }
        } else {
            fp_limbs_sub(v, v, u);
            if (fp_limbs_ge(x2, x1)) fp_limbs_sub(x2, x2, x1);
            else { uint32_t t[NL]; fp_limbs_sub(t, x1, x2); fp_limbs_sub(x2, FP_P.v, t); }
        }
    }
    Fp inv;
    const uint32_t* src = is_one(u) ? x1 : x2;
#pragma unroll
    for (int i = 0; i < NL; i++) inv.v[i] = src[i];
    fp_mul(r, inv, FP_R3);   // (a R)^-1 * R^3 / R = a^-1 R
}
#endif

// square root candidate: r = a^((p+1)/4); returns true iff r^2 == a
LHB_HD LHB_INLINE bool fp_sqrt(Fp& r, const Fp& a) {
    Fp t, c;
    fp_pow_pm3d4(t, a);
    fp_mul(t, t, a);
    fp_sqr(c, t);
    r = t;
    return fp_eq(c, a);
}

// canonical (non-Montgomery) comparison a > (p-1)/2, used for the serialisation sign bit
LHB_HD LHB_INLINE bool fp_canon_gt_half(const Fp& canon) {
    uint32_t t, brw;
    sub_cc(t, FP_HALF_P.v[0], canon.v[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) subc_cc(t, FP_HALF_P.v[i], canon.v[i]);
    subc(brw, 0, 0);
    (void)t;
    return brw != 0;  // borrow <=> canon > half
}
// canonical value < p ?
LHB_HD LHB_INLINE bool fp_canon_lt_p(const Fp& canon) {
    uint32_t t, brw;
    sub_cc(t, canon.v[0], FP_P.v[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) subc_cc(t, canon.v[i], FP_P.v[i]);
    subc(brw, 0, 0);
    (void)t;
    return brw != 0;
}

// 48 big-endian bytes <-> canonical limbs
LHB_HD LHB_INLINE void fp_from_be48(Fp& canon, const uint8_t* b) {
#pragma unroll
    for (int i = 0; i < NL; i++) {
        const uint8_t* q = b + 4 * (NL - 1 - i);
        canon.v[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
}
LHB_HD LHB_INLINE void fp_to_be48(uint8_t* b, const Fp& canon) {
#pragma unroll
    for (int i = 0; i < NL; i++) {
        uint8_t* q = b + 4 * (NL - 1 - i);
        q[0] = canon.v[i] >> 24; q[1] = canon.v[i] >> 16; q[2] = canon.v[i] >> 8; q[3] = canon.v[i];
    }
}

}  // namespace bls
}  // namespace lhb200
