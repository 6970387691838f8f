// fp2.cuh — Fp2 = Fp[i]/(i^2+1); Fp6 = Fp2[v]/(v^3 - xi), xi = 1+i; Fp12 = Fp6[w]/(w^2 - v).
#pragma once
#include "fp.cuh"

namespace lhb200 {
namespace bls {

struct Fp6 {
    Fp2 c0, c1, c2;
};
struct Fp12 {
    Fp6 c0, c1;
};

// ------------------------------------------------------------------------------------------------ Fp2
// The Fp2 leaf operations are out-of-line and REGISTER-FUSED: operands are loaded once (LD.128), the whole Fp2
// operation (three Montgomery products for a multiplication) runs in registers with its independent carry
// chains interleaved by ptxas, and only the result is stored.  Versus composing fp_mul/fp_add calls this cuts
// local-memory traffic ~3x and call overhead ~5x (profiles/r1_ncu_k_miller_100k_baseline.txt: the baseline was
// stalled on local-memory latency).  Like the Fp leafs they live in the opaque TU (bls/fp_core.cu, §7 DESIGN.md).
LHB_HD LHB_INLINE bool fp2_is_zero(const Fp2& a) { return fp_is_zero(a.c0) & fp_is_zero(a.c1); }
LHB_HD LHB_INLINE bool fp2_eq(const Fp2& a, const Fp2& b) { return fp_eq(a.c0, b.c0) & fp_eq(a.c1, b.c1); }
LHB_HD LHB_INLINE void fp2_set_zero(Fp2& a) { fp_set_zero(a.c0); fp_set_zero(a.c1); }
LHB_HD LHB_INLINE void fp2_set_one(Fp2& a) { a.c0 = FP_ONE; fp_set_zero(a.c1); }
LHB_HD LHB_INLINE void fp2_cmov(Fp2& r, const Fp2& a, bool c) { fp_cmov(r.c0, a.c0, c); fp_cmov(r.c1, a.c1, c); }

#ifdef LHB_FP_DECL_ONLY
LHB_HD void fp2_add(Fp2& r, const Fp2& a, const Fp2& b);
LHB_HD void fp2_sub(Fp2& r, const Fp2& a, const Fp2& b);
#ifdef LHB_FP2_ALIAS_LIGHT
LHB_CONST Fp2 FP2_ZERO_C = {{{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}}, {{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}}};
LHB_HD LHB_INLINE void fp2_dbl(Fp2& r, const Fp2& a) { fp2_add(r, a, a); }
LHB_HD LHB_INLINE void fp2_neg(Fp2& r, const Fp2& a) { fp2_sub(r, FP2_ZERO_C, a); }
#else
LHB_HD void fp2_dbl(Fp2& r, const Fp2& a);
LHB_HD void fp2_neg(Fp2& r, const Fp2& a);
#endif
LHB_HD void fp2_conj(Fp2& r, const Fp2& a);
LHB_HD void fp2_mul(Fp2& r, const Fp2& a, const Fp2& b);
LHB_HD void fp2_sqr(Fp2& r, const Fp2& a);
LHB_HD void fp2_mul_fp(Fp2& r, const Fp2& a, const Fp& s);
LHB_HD void fp2_mul_xi(Fp2& r, const Fp2& a);
#else
#ifndef LHB_FP2_INLINE_MUL
// The Montgomery product is ONE shared out-of-line body instead of three inlined copies per Fp2 multiplication:
// with the copies inlined the Miller loop's hot code was ~75 KB and k_miller lost 2.4 warp-cycles per issue to
// instruction fetch (`no_instruction`); shared, it is ~45 KB, the stall drops to 0.15 and k_miller from 55.0 to
// 45.6 ms at 100 k sets (profiles/r1_ncu_bls_100k_smallcode.txt).  -DLHB_FP2_INLINE_MUL restores the old form.
LHB_HD LHB_NOINLINE Fp fp_mul_rr(Fp a, Fp b) { Fp o; fp_mul_inl(o, a, b); return o; }
#define LHB_MUL(o, a, b) o = fp_mul_rr(a, b)
#else
#define LHB_MUL(o, a, b) fp_mul_inl(o, a, b)
#endif
LHB_HD LHB_NOINLINE void fp2_add(Fp2& r, const Fp2& a, const Fp2& b) {
    Fp2 x = a, y = b, o;
    fp_add_inl(o.c0, x.c0, y.c0);
    fp_add_inl(o.c1, x.c1, y.c1);
    r = o;
}
LHB_HD LHB_NOINLINE void fp2_sub(Fp2& r, const Fp2& a, const Fp2& b) {
    Fp2 x = a, y = b, o;
    fp_sub_inl(o.c0, x.c0, y.c0);
    fp_sub_inl(o.c1, x.c1, y.c1);
    r = o;
}
#ifdef LHB_FP2_ALIAS_LIGHT
// fewer distinct leaf bodies in the instruction cache: 2a = a + a, -a = 0 - a
LHB_CONST Fp2 FP2_ZERO_C = {{{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}}, {{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}}};
LHB_HD LHB_INLINE void fp2_dbl(Fp2& r, const Fp2& a) { fp2_add(r, a, a); }
LHB_HD LHB_INLINE void fp2_neg(Fp2& r, const Fp2& a) { fp2_sub(r, FP2_ZERO_C, a); }
#else
LHB_HD LHB_NOINLINE void fp2_dbl(Fp2& r, const Fp2& a) {
    Fp2 x = a, o;
    fp_add_inl(o.c0, x.c0, x.c0);
    fp_add_inl(o.c1, x.c1, x.c1);
    r = o;
}
LHB_HD LHB_NOINLINE void fp2_neg(Fp2& r, const Fp2& a) {
    Fp2 x = a, o;
    fp_neg(o.c0, x.c0);
    fp_neg(o.c1, x.c1);
    r = o;
}
#endif
LHB_HD LHB_NOINLINE void fp2_conj(Fp2& r, const Fp2& a) {
    Fp2 x = a, o;
    o.c0 = x.c0;
    fp_neg(o.c1, x.c1);
    r = o;
}
// The three product leaves keep almost nothing live across their fp_mul_rr calls: operands are (re)loaded from memory
// (L1) right before each product instead of being held in registers for the whole function.  fp_mul_rr itself needs
// ~100 registers, so a caller that keeps 4 operands + 2 products alive pushed every Fp2 kernel to 229 registers
// (8 warps/SM); in this form (bls/fp_core.cu is compiled with -maxrregcount=128) the Fp2 kernels need 150-168
// registers (12 warps/SM) and a 100 k-set verify went from 95.0 to 92.6 ms.
#ifdef __CUDA_ARCH__
#define LHB_BARRIER() asm volatile("" ::: "memory")
#else
#define LHB_BARRIER() do { } while (0)
#endif
#ifdef LHB_FP2_LAZY
// Karatsuba with LAZY REDUCTION: three full 768-bit products, the recombination in double width, two Montgomery
// reductions instead of three (744 multiply instructions instead of 900).  Operand sums stay unreduced (< 2p < 2^382),
// a0 b0 - a1 b1 is lifted by p 2^384 when negative so both reductions see a value in [0, p 2^384).
LHB_HD LHB_INLINE void fp_add_nored(Fp& r, const Fp& a, const Fp& b) {
    add_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) addc_cc(r.v[i], a.v[i], b.v[i]);
    addc(r.v[NL - 1], a.v[NL - 1], b.v[NL - 1]);
}
// EXPERIMENT (-DLHB_FP2_LAZY, not the shipped default — DESIGN.md §9): in isolation this leaf is 22 % faster than the
// eager one (scripts/ubench/fp2_leaf.cu), but with three products and two reductions inlined it is 20 KB of code and
// the Fp2 kernels fall out of the instruction cache (no_instruction stall 0.19 -> 1.62 in k_miller_multi; 92.5 ->
// 95.8-100.6 ms per 100 k sets); with the product and the reduction as shared out-of-line bodies the 24-limb values
// travel through local memory and the step is 107-113 ms.
struct FpW {
    uint32_t v[2 * NL];
};
LHB_HD LHB_INLINE FpW fp_mulw_rr(Fp a, Fp b) { FpW o; fp_mulw_inl(o.v, a, b); return o; }
#ifdef LHB_FP2_LAZY_SHARED_REDC
// the reduction as ONE shared body; the 768-bit value travels as two 12-limb register arguments (like fp_mul_rr's)
LHB_HD LHB_NOINLINE Fp fp_redc_rr2(Fp lo, Fp hi) {
    uint32_t w[2 * NL];
#pragma unroll
    for (int i = 0; i < NL; i++) { w[i] = lo.v[i]; w[NL + i] = hi.v[i]; }
    Fp o;
    fp_redc_inl(o, w);
    return o;
}
LHB_HD LHB_INLINE Fp fp_redc_rr(const FpW& w) {
    Fp lo, hi;
#pragma unroll
    for (int i = 0; i < NL; i++) { lo.v[i] = w.v[i]; hi.v[i] = w.v[NL + i]; }
    return fp_redc_rr2(lo, hi);
}
#else
LHB_HD LHB_INLINE Fp fp_redc_rr(const FpW& w) { Fp o; fp_redc_inl(o, w.v); return o; }
#endif
LHB_HD LHB_NOINLINE void fp2_mul(Fp2& r, const Fp2& a, const Fp2& b) {
    FpW w0, w1, w2;
    { Fp p = a.c0, q = b.c0; w0 = fp_mulw_rr(p, q); }
    { Fp p = a.c1, q = b.c1; w1 = fp_mulw_rr(p, q); }
    {
        Fp s0, s1;
        { Fp p = a.c0, q = a.c1; fp_add_nored(s0, p, q); }
        { Fp p = b.c0, q = b.c1; fp_add_nored(s1, p, q); }
        w2 = fp_mulw_rr(s0, s1);
    }
    uint32_t *t0 = w0.v, *t1 = w1.v, *t2 = w2.v;
    // c1 = (a0+a1)(b0+b1) - a0 b0 - a1 b1  (= a0 b1 + a1 b0 >= 0, < 2 p^2)
    sub_cc(t2[0], t2[0], t0[0]);
#pragma unroll
    for (int i = 1; i < 2 * NL - 1; i++) subc_cc(t2[i], t2[i], t0[i]);
    subc(t2[2 * NL - 1], t2[2 * NL - 1], t0[2 * NL - 1]);
    sub_cc(t2[0], t2[0], t1[0]);
#pragma unroll
    for (int i = 1; i < 2 * NL - 1; i++) subc_cc(t2[i], t2[i], t1[i]);
    subc(t2[2 * NL - 1], t2[2 * NL - 1], t1[2 * NL - 1]);
    // c0 = a0 b0 - a1 b1, + p 2^384 when negative
    uint32_t m;
    sub_cc(t0[0], t0[0], t1[0]);
#pragma unroll
    for (int i = 1; i < 2 * NL; i++) subc_cc(t0[i], t0[i], t1[i]);
    subc(m, 0, 0);  // 0xffffffff on borrow
    add_cc(t0[NL], t0[NL], FP_P.v[0] & m);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) addc_cc(t0[NL + i], t0[NL + i], FP_P.v[i] & m);
    addc(t0[2 * NL - 1], t0[2 * NL - 1], FP_P.v[NL - 1] & m);
    const Fp o0 = fp_redc_rr(w0);
    const Fp o1 = fp_redc_rr(w2);
    r.c0 = o0;
    r.c1 = o1;
}
#else
// Karatsuba: 3 Montgomery products
LHB_HD LHB_NOINLINE void fp2_mul(Fp2& r, const Fp2& a, const Fp2& b) {
    Fp t0, t1, s0, s1;
    { Fp p = a.c0, q = b.c0; LHB_MUL(t0, p, q); }
    LHB_BARRIER();
    { Fp p = a.c1, q = b.c1; LHB_MUL(t1, p, q); }
    LHB_BARRIER();
    { Fp p = a.c0, q = a.c1; fp_add_inl(s0, p, q); }
    { Fp p = b.c0, q = b.c1; fp_add_inl(s1, p, q); }
    LHB_MUL(s0, s0, s1);
    Fp o0;
    fp_sub_inl(o0, t0, t1);
    fp_sub_inl(s0, s0, t0);
    fp_sub_inl(s0, s0, t1);
    r.c0 = o0;
    r.c1 = s0;
}
#endif  // LHB_FP2_LAZY
// (a0+a1)(a0-a1), 2 a0 a1 : 2 Montgomery products
LHB_HD LHB_NOINLINE void fp2_sqr(Fp2& r, const Fp2& a) {
    Fp s, d, m;
    { Fp p = a.c0, q = a.c1; LHB_MUL(m, p, q); }
    LHB_BARRIER();
    { Fp p = a.c0, q = a.c1; fp_add_inl(s, p, q); fp_sub_inl(d, p, q); }
    LHB_MUL(s, s, d);
    fp_add_inl(m, m, m);
    r.c0 = s;
    r.c1 = m;
}
LHB_HD LHB_NOINLINE void fp2_mul_fp(Fp2& r, const Fp2& a, const Fp& s) {
    Fp o0, o1;
    { Fp p = a.c0, k = s; LHB_MUL(o0, p, k); }
    LHB_BARRIER();
    { Fp p = a.c1, k = s; LHB_MUL(o1, p, k); }
    r.c0 = o0;
    r.c1 = o1;
}
// multiply by xi = 1 + i
LHB_HD LHB_NOINLINE void fp2_mul_xi(Fp2& r, const Fp2& a) {
    Fp2 x = a, o;
    fp_sub_inl(o.c0, x.c0, x.c1);
    fp_add_inl(o.c1, x.c0, x.c1);
    r = o;
}
#endif
#ifndef LHB_FP_CORE_ONLY  // everything below is compiled only in the callers' TU
LHB_HD LHB_INLINE void fp2_inv(Fp2& r, const Fp2& a) {
    Fp n, t;
    fp_sqr(n, a.c0);
    fp_sqr(t, a.c1);
    fp_add(n, n, t);
    fp_inv(n, n);
    fp_mul(r.c0, a.c0, n);
    fp_mul(t, a.c1, n);
    fp_neg(r.c1, t);
}
// the same with the variable-time Fp inversion (public values only: the final exponentiation)
LHB_HD LHB_INLINE void fp2_inv_vartime(Fp2& r, const Fp2& a) {
    Fp n, t;
    fp_sqr(n, a.c0);
    fp_sqr(t, a.c1);
    fp_add(n, n, t);
    fp_inv_vartime(n, n);
    fp_mul(r.c0, a.c0, n);
    fp_mul(t, a.c1, n);
    fp_neg(r.c1, t);
}
// RFC 9380 sgn0 (m = 2) of a Montgomery-form element
LHB_HD LHB_INLINE uint32_t fp2_sgn0(const Fp2& a) {
    Fp c0, c1;
    fp_from_mont(c0, a.c0);
    fp_from_mont(c1, a.c1);
    const uint32_t s0 = c0.v[0] & 1, z0 = fp_is_zero(c0) ? 1u : 0u, s1 = c1.v[0] & 1;
    return s0 | (z0 & s1);
}

// Square root in Fp2 by the norm method, two Fp exponentiations:
//   n = sqrt(a0^2 + a1^2);  d = (a0 + n)/2;  if d is a square: (sqrt d, a1 / (2 sqrt d)) else (a1 / (2 s), s), s = sqrt(-d).
// Returns false when a is not a square.  `norm_root`/`is_qr` let SSWU reuse the first exponentiation.
LHB_HD LHB_NOINLINE bool fp2_sqrt_with_norm_root(Fp2& r, const Fp2& a, const Fp& n) {
    // caller guarantees n^2 == a0^2 + a1^2
    Fp d, t, x0, chk, inv;
    fp_add(d, a.c0, n);
    fp_mul(d, d, FP_INV2);
    if (fp_is_zero(d)) {
        // a0 + n == 0: take the other root of the norm (d' = a0 when a1 == 0 and a0 = -n)
        fp_sub(d, a.c0, n);
        fp_mul(d, d, FP_INV2);
    }
    fp_pow_pm3d4(t, d);      // t = d^((p-3)/4)
    fp_mul(x0, t, d);        // candidate sqrt(d) (or sqrt(-d))
    fp_sqr(chk, x0);
    const bool d_is_qr = fp_eq(chk, d);
    // 1/x0 : x0 * t = d^((p-1)/2) = +-1  =>  1/x0 = +-t
    fp_mul(inv, t, FP_INV2);  // t/2
    if (!d_is_qr) fp_neg(inv, inv);
    Fp other;
    fp_mul(other, a.c1, inv);  // a1 / (2 x0)
    if (d_is_qr) { r.c0 = x0; r.c1 = other; }
    else { r.c0 = other; r.c1 = x0; }
    Fp2 sq;
    fp2_sqr(sq, r);
    return fp2_eq(sq, a);
}
LHB_HD LHB_INLINE bool fp2_sqrt(Fp2& r, const Fp2& a) {
    Fp n, t;
    fp_sqr(n, a.c0);
    fp_sqr(t, a.c1);
    fp_add(n, n, t);
    Fp root;
    if (!fp_sqrt(root, n)) return false;
    return fp2_sqrt_with_norm_root(r, a, root);
}

// ------------------------------------------------------------------------------------------------ Fp6
LHB_HD LHB_INLINE void fp6_add(Fp6& r, const Fp6& a, const Fp6& b) { fp2_add(r.c0, a.c0, b.c0); fp2_add(r.c1, a.c1, b.c1); fp2_add(r.c2, a.c2, b.c2); }
LHB_HD LHB_INLINE void fp6_sub(Fp6& r, const Fp6& a, const Fp6& b) { fp2_sub(r.c0, a.c0, b.c0); fp2_sub(r.c1, a.c1, b.c1); fp2_sub(r.c2, a.c2, b.c2); }
LHB_HD LHB_INLINE void fp6_neg(Fp6& r, const Fp6& a) { fp2_neg(r.c0, a.c0); fp2_neg(r.c1, a.c1); fp2_neg(r.c2, a.c2); }
LHB_HD LHB_INLINE void fp6_set_zero(Fp6& a) { fp2_set_zero(a.c0); fp2_set_zero(a.c1); fp2_set_zero(a.c2); }
LHB_HD LHB_INLINE bool fp6_eq(const Fp6& a, const Fp6& b) { return fp2_eq(a.c0, b.c0) & fp2_eq(a.c1, b.c1) & fp2_eq(a.c2, b.c2); }
// multiply by v: (c0, c1, c2) -> (xi c2, c0, c1)
LHB_HD LHB_INLINE void fp6_mul_v(Fp6& r, const Fp6& a) {
    Fp2 t;
    fp2_mul_xi(t, a.c2);
    r.c2 = a.c1;
    r.c1 = a.c0;
    r.c0 = t;
}
// Karatsuba-style (Devegili et al.): 6 Fp2 multiplications
LHB_HD LHB_NOINLINE void fp6_mul(Fp6& r, const Fp6& a, const Fp6& b) {
    Fp2 v0, v1, v2, t0, t1, t2, c0, c1, c2;
    fp2_mul(v0, a.c0, b.c0);
    fp2_mul(v1, a.c1, b.c1);
    fp2_mul(v2, a.c2, b.c2);
    // c0 = v0 + xi((a1+a2)(b1+b2) - v1 - v2)
    fp2_add(t0, a.c1, a.c2);
    fp2_add(t1, b.c1, b.c2);
    fp2_mul(t2, t0, t1);
    fp2_sub(t2, t2, v1);
    fp2_sub(t2, t2, v2);
    fp2_mul_xi(t2, t2);
    fp2_add(c0, t2, v0);
    // c1 = (a0+a1)(b0+b1) - v0 - v1 + xi v2
    fp2_add(t0, a.c0, a.c1);
    fp2_add(t1, b.c0, b.c1);
    fp2_mul(t2, t0, t1);
    fp2_sub(t2, t2, v0);
    fp2_sub(t2, t2, v1);
    fp2_mul_xi(t0, v2);
    fp2_add(c1, t2, t0);
    // c2 = (a0+a2)(b0+b2) - v0 - v2 + v1
    fp2_add(t0, a.c0, a.c2);
    fp2_add(t1, b.c0, b.c2);
    fp2_mul(t2, t0, t1);
    fp2_sub(t2, t2, v0);
    fp2_sub(t2, t2, v2);
    fp2_add(c2, t2, v1);
    r.c0 = c0; r.c1 = c1; r.c2 = c2;
}
// a * (b0 + b1 v): 5 Fp2 multiplications
LHB_HD LHB_NOINLINE void fp6_mul_by_01(Fp6& r, const Fp6& a, const Fp2& b0, const Fp2& b1) {
    Fp2 v0, v1, t0, t1, t2, c0, c1, c2;
    fp2_mul(v0, a.c0, b0);
    fp2_mul(v1, a.c1, b1);
    // c0 = v0 + xi * a2 b1
    fp2_mul(t0, a.c2, b1);
    fp2_mul_xi(t0, t0);
    fp2_add(c0, t0, v0);
    // c1 = (a0+a1)(b0+b1) - v0 - v1
    fp2_add(t0, a.c0, a.c1);
    fp2_add(t1, b0, b1);
    fp2_mul(t2, t0, t1);
    fp2_sub(t2, t2, v0);
    fp2_sub(c1, t2, v1);
    // c2 = a2 b0 + v1
    fp2_mul(t0, a.c2, b0);
    fp2_add(c2, t0, v1);
    r.c0 = c0; r.c1 = c1; r.c2 = c2;
}
// a * (b1 v): 3 Fp2 multiplications
LHB_HD LHB_INLINE void fp6_mul_by_1(Fp6& r, const Fp6& a, const Fp2& b1) {
    Fp2 c0, c1, c2;
    fp2_mul(c0, a.c2, b1);
    fp2_mul_xi(c0, c0);
    fp2_mul(c1, a.c0, b1);
    fp2_mul(c2, a.c1, b1);
    r.c0 = c0; r.c1 = c1; r.c2 = c2;
}
LHB_HD LHB_NOINLINE void fp6_inv(Fp6& r, const Fp6& a) {
    Fp2 t0, t1, t2, d, s;
    // t0 = a0^2 - xi a1 a2 ; t1 = xi a2^2 - a0 a1 ; t2 = a1^2 - a0 a2
    fp2_sqr(t0, a.c0);
    fp2_mul(s, a.c1, a.c2);
    fp2_mul_xi(s, s);
    fp2_sub(t0, t0, s);
    fp2_sqr(t1, a.c2);
    fp2_mul_xi(t1, t1);
    fp2_mul(s, a.c0, a.c1);
    fp2_sub(t1, t1, s);
    fp2_sqr(t2, a.c1);
    fp2_mul(s, a.c0, a.c2);
    fp2_sub(t2, t2, s);
    // d = a0 t0 + xi (a2 t1 + a1 t2)
    fp2_mul(d, a.c2, t1);
    fp2_mul(s, a.c1, t2);
    fp2_add(d, d, s);
    fp2_mul_xi(d, d);
    fp2_mul(s, a.c0, t0);
    fp2_add(d, d, s);
    fp2_inv_vartime(d, d);   // fp6_inv / fp12_inv only serve the final exponentiation (public values)
    fp2_mul(r.c0, t0, d);
    fp2_mul(r.c1, t1, d);
    fp2_mul(r.c2, t2, d);
}

// ------------------------------------------------------------------------------------------------ Fp12
LHB_HD LHB_INLINE void fp12_set_one(Fp12& a) {
    fp6_set_zero(a.c0);
    fp6_set_zero(a.c1);
    a.c0.c0.c0 = FP_ONE;
}
LHB_HD LHB_INLINE bool fp12_eq(const Fp12& a, const Fp12& b) { return fp6_eq(a.c0, b.c0) & fp6_eq(a.c1, b.c1); }
LHB_HD LHB_INLINE bool fp12_is_one(const Fp12& a) {
    Fp12 o;
    fp12_set_one(o);
    return fp12_eq(a, o);
}
LHB_HD LHB_INLINE void fp12_conj(Fp12& r, const Fp12& a) { r.c0 = a.c0; fp6_neg(r.c1, a.c1); }
// 3 Fp6 multiplications
LHB_HD LHB_NOINLINE void fp12_mul(Fp12& r, const Fp12& a, const Fp12& b) {
    Fp6 t0, t1, s0, s1, m;
    fp6_mul(t0, a.c0, b.c0);
    fp6_mul(t1, a.c1, b.c1);
    fp6_add(s0, a.c0, a.c1);
    fp6_add(s1, b.c0, b.c1);
    fp6_mul(m, s0, s1);
    fp6_sub(m, m, t0);
    fp6_sub(r.c1, m, t1);
    fp6_mul_v(t1, t1);
    fp6_add(r.c0, t0, t1);
}
// complex squaring: 2 Fp6 multiplications, three Fp6 temporaries (stack footprint matters: DESIGN.md §2.3)
LHB_HD LHB_NOINLINE void fp12_sqr(Fp12& r, const Fp12& a) {
    Fp6 m, s, t;
    fp6_mul(m, a.c0, a.c1);          // a0 a1
    fp6_mul_v(t, a.c1);
    fp6_add(t, t, a.c0);             // a0 + v a1
    fp6_add(s, a.c0, a.c1);          // a0 + a1
    fp6_mul(s, s, t);                // a0^2 + v a1^2 + (1+v) a0 a1
    fp6_sub(s, s, m);
    fp6_mul_v(t, m);
    fp6_sub(r.c0, s, t);
    fp6_add(r.c1, m, m);
}
// Sparse multiplication by a line  l = c0 + c1 v + c4 (v w)   (Fp12 slots (0,0),(0,1),(1,1)): 13 Fp2 muls
LHB_HD LHB_NOINLINE void fp12_mul_by_014(Fp12& r, const Fp12& a, const Fp2& c0, const Fp2& c1, const Fp2& c4) {
    Fp6 t0, t1, s;
    Fp2 c14;
    fp6_mul_by_01(t0, a.c0, c0, c1);   // a0 * (c0 + c1 v)
    fp6_mul_by_1(t1, a.c1, c4);        // a1 * (c4 v)
    fp2_add(c14, c1, c4);
    fp6_add(s, a.c0, a.c1);
    fp6_mul_by_01(s, s, c0, c14);      // (a0+a1)(c0 + (c1+c4) v)
    fp6_sub(s, s, t0);
    fp6_sub(r.c1, s, t1);
    fp6_mul_v(t1, t1);
    fp6_add(r.c0, t0, t1);
}
LHB_HD LHB_NOINLINE void fp12_inv(Fp12& r, const Fp12& a) {
    Fp6 t0, t1;
    fp6_mul(t0, a.c0, a.c0);
    fp6_mul(t1, a.c1, a.c1);
    fp6_mul_v(t1, t1);
    fp6_sub(t0, t0, t1);
    fp6_inv(t0, t0);
    fp6_mul(r.c0, a.c0, t0);
    fp6_mul(t1, a.c1, t0);
    fp6_neg(r.c1, t1);
}
// Frobenius a -> a^p.  Coefficient of w^k (k = 0..5 over Fp2: c0.c0,c1.c0,c0.c1,c1.c1,c0.c2,c1.c2) is
// conjugated and multiplied by FROB_G1[k] = xi^(k(p-1)/6).
LHB_HD LHB_NOINLINE void fp12_frob(Fp12& r, const Fp12& a) {
    Fp2 t;
    fp2_conj(r.c0.c0, a.c0.c0);
    fp2_conj(t, a.c1.c0); fp2_mul(r.c1.c0, t, FROB_G1[1]);
    fp2_conj(t, a.c0.c1); fp2_mul(r.c0.c1, t, FROB_G1[2]);
    fp2_conj(t, a.c1.c1); fp2_mul(r.c1.c1, t, FROB_G1[3]);
    fp2_conj(t, a.c0.c2); fp2_mul(r.c0.c2, t, FROB_G1[4]);
    fp2_conj(t, a.c1.c2); fp2_mul(r.c1.c2, t, FROB_G1[5]);
}
// a -> a^(p^2): coefficient of w^k times FROB_G2[k] (in Fp), no conjugation
LHB_HD LHB_NOINLINE void fp12_frob2(Fp12& r, const Fp12& a) {
    r.c0.c0 = a.c0.c0;
    fp2_mul_fp(r.c1.c0, a.c1.c0, FROB_G2[1]);
    fp2_mul_fp(r.c0.c1, a.c0.c1, FROB_G2[2]);
    fp2_mul_fp(r.c1.c1, a.c1.c1, FROB_G2[3]);
    fp2_mul_fp(r.c0.c2, a.c0.c2, FROB_G2[4]);
    fp2_mul_fp(r.c1.c2, a.c1.c2, FROB_G2[5]);
}

// Granger–Scott squaring for elements of the cyclotomic subgroup (after the easy part of the final
// exponentiation): 9 Fp2 squarings instead of 12 Fp2 multiplications.
LHB_HD LHB_INLINE void fp4_sqr(Fp2& r0, Fp2& r1, const Fp2& a, const Fp2& b) {
    // (a + b s)^2 with s^2 = xi : r0 = a^2 + xi b^2, r1 = 2ab = (a+b)^2 - a^2 - b^2
    Fp2 t0, t1, t2;
    fp2_sqr(t0, a);
    fp2_sqr(t1, b);
    fp2_add(t2, a, b);
    fp2_sqr(t2, t2);
    fp2_sub(t2, t2, t0);
    fp2_sub(r1, t2, t1);
    fp2_mul_xi(t1, t1);
    fp2_add(r0, t0, t1);
}
LHB_HD LHB_NOINLINE void fp12_cyclotomic_sqr(Fp12& r, const Fp12& a) {
    // a = g0 + g1 w, g0 = (z0, z4, z3), g1 = (z2, z1, z5) in the notation of Granger-Scott (ePrint 2009/565 §3.2)
    const Fp2 &z0 = a.c0.c0, &z4 = a.c0.c1, &z3 = a.c0.c2, &z2 = a.c1.c0, &z1 = a.c1.c1, &z5 = a.c1.c2;
    Fp2 t0, t1, t2, t3, t4, t5, u;
    fp4_sqr(t0, t1, z0, z1);
    fp4_sqr(t2, t3, z2, z3);
    fp4_sqr(t4, t5, z4, z5);
    Fp12 o;
    // z0' = 3 t0 - 2 z0
    fp2_sub(u, t0, z0); fp2_dbl(u, u); fp2_add(o.c0.c0, u, t0);
    // z1' = 3 t1 + 2 z1
    fp2_add(u, t1, z1); fp2_dbl(u, u); fp2_add(o.c1.c1, u, t1);
    // z2' = 3 xi t5 + 2 z2
    Fp2 x5;
    fp2_mul_xi(x5, t5);
    fp2_add(u, x5, z2); fp2_dbl(u, u); fp2_add(o.c1.c0, u, x5);
    // z3' = 3 t4 - 2 z3
    fp2_sub(u, t4, z3); fp2_dbl(u, u); fp2_add(o.c0.c2, u, t4);
    // z4' = 3 t2 - 2 z4
    fp2_sub(u, t2, z4); fp2_dbl(u, u); fp2_add(o.c0.c1, u, t2);
    // z5' = 3 t3 + 2 z5
    fp2_add(u, t3, z5); fp2_dbl(u, u); fp2_add(o.c1.c2, u, t3);
    r = o;
}

#endif  // !LHB_FP_CORE_ONLY

}  // namespace bls
}  // namespace lhb200
