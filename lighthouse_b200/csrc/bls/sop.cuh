// sop.cuh — fused SUM-OF-PRODUCTS Montgomery arithmetic and Fp2 operations on shared-memory "columns".
//
//   fp_sop2<K>:  rA = (sum_{q<K} xA_q * yA_q) / R mod p   and   rB = (sum_{q<K} xB_q * yB_q) / R mod p
//
// computed row by row in ONE sliding 13-limb window per result: every row adds the K partial-product rows and then
// one Montgomery reduction row, so a sum of K products costs 144 K + 144 multiply instructions instead of the
// 288 K of K separate Montgomery products (lazy reduction without any double-width intermediate), and an Fp2
// product  (u0 v0 - u1 v1,  u0 v1 + u1 v0)  needs no Karatsuba glue at all.  The x operands live in registers; the y
// operands are STREAMED from shared memory one limb per row, which lets the row loop stay rolled (the hot code of a
// K = 6 instance is ~7 KB instead of ~40 KB: the round-1 kernels were instruction-cache bound).
// The two results are independent carry chains that ptxas interleaves (the second half of the ILP the kernels need at
// the low occupancy that shared-memory-resident state implies).
//
// Bounds (p ~ 0.1016 * 2^384): with x_q < X p and y_q < Y p the window never exceeds 13 limbs when K X <= 8, and the
// result is < (0.1016 K X Y + 1) p, so one conditional subtraction returns it to [0, p) when K X Y <= 9.  Every call
// site states its (K, X, Y).
//
// This is the arithmetic Lighthouse gets from blst behind crypto/bls/src/impls/blst.rs:114-118; it is not a port of
// blst's 6 x 64-bit mulx/adx code.  Compiles for the host too (LHB_HOSTSIM) so tests check it limb-exactly.
#pragma once
#include "fp.cuh"

namespace lhb200 {
namespace bls {

// ------------------------------------------------------------------------------------------------ window rows
// window value = sum even[i] 2^(32 i) + sum odd[i] 2^(32 (i+1)); rows alternate the roles of the two arrays.
// (T >> 32) + a * bi, where on entry `odd` holds the previous row's even array (its limb 0 is zero after the
// reduction row) and `even` the previous row's odd array.  Also correct on an all-zero window (first row).
LHB_HD LHB_INLINE void sop_shift_acc(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi) {
    add_cc(even[0], even[0], odd[1]);
#pragma unroll
    for (int j = 0; j < NL - 2; j += 2) mad_pair_sh(odd[j], odd[j + 1], a[j + 1], bi, odd[j + 2], odd[j + 3]);
    mad_pair_last(odd[NL - 2], odd[NL - 1], a[NL - 1], bi);
    mad_pair_first(even[0], even[1], a[0], bi);
#pragma unroll
    for (int j = 2; j < NL; j += 2) mad_pair(even[j], even[j + 1], a[j], bi);
    addc(odd[NL - 1], odd[NL - 1], 0);
}
// T += a * bi (same alignment)
LHB_HD LHB_INLINE void sop_acc(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi) {
    mad_pair_first(odd[0], odd[1], a[1], bi);
#pragma unroll
    for (int j = 2; j < NL; j += 2) mad_pair(odd[j], odd[j + 1], a[j + 1], bi);
    mad_pair_first(even[0], even[1], a[0], bi);
#pragma unroll
    for (int j = 2; j < NL; j += 2) mad_pair(even[j], even[j + 1], a[j], bi);
    addc(odd[NL - 1], odd[NL - 1], 0);
}
// m = T[0] * M0 ; T += m * p   (T[0] becomes 0)
LHB_HD LHB_INLINE void sop_redc(uint32_t* even, uint32_t* odd) {
    const uint32_t mi = even[0] * LHB_FP_M0;
    mad_pair_first(odd[0], odd[1], FP_P.v[1], mi);
#pragma unroll
    for (int j = 2; j < NL; j += 2) mad_pair(odd[j], odd[j + 1], FP_P.v[j + 1], mi);
    mad_pair_first(even[0], even[1], FP_P.v[0], mi);
#pragma unroll
    for (int j = 2; j < NL; j += 2) mad_pair(even[j], even[j + 1], FP_P.v[j], mi);
    addc(odd[NL - 1], odd[NL - 1], 0);
}
// result = (T >> 32) after the last row, then one conditional subtraction (T >> 32 < 2p)
LHB_HD LHB_INLINE void sop_finish(Fp& r, uint32_t* even, uint32_t* odd) {
    add_cc(even[0], even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) addc_cc(even[i], even[i], odd[i + 1]);
    addc(even[NL - 1], even[NL - 1], 0);
    fp_final_sub(r, even, 0);
}

#ifndef LHB_SOP_UNROLL_SMALL
#define LHB_SOP_UNROLL_SMALL 0   // measured: -11 % instructions but +1.3 ms in k_miller_coop (instruction cache), DESIGN.md §9
#endif
#ifndef LHB_SOP_UNROLL_BIG
#define LHB_SOP_UNROLL_BIG 1
#endif
// K register-resident x operands
template <int K>
struct SopX {
    Fp x[K];
};
// K y operands streamed from a strided word array (shared memory column, or a plain array on the host):
// limb j of operand q is base[q][j * stride]
template <int K>
struct SopY {
    const uint32_t* base[K];
    int stride;
    LHB_HD LHB_INLINE uint32_t limb(int q, int j) const { return base[q][j * stride]; }
};

template <int K>
LHB_HD LHB_INLINE void sop_row(uint32_t* even, uint32_t* odd, const SopX<K>& x, const uint32_t* y) {
    sop_shift_acc(even, odd, x.x[0].v, y[0]);
#pragma unroll
    for (int q = 1; q < K; q++) sop_acc(even, odd, x.x[q].v, y[q]);
    sop_redc(even, odd);
}

// Two fused sums of K products each (see the header comment).  The row loop is rolled (6 iterations of two rows per
// window).  The y limbs of a row pair are fetched at the top of its iteration: ~30 cycles of shared-memory latency per
// ~2000 cycles of multiply work; an explicit double buffer cost 4 K register copies per iteration and 4 K registers.
template <int K>
LHB_HD LHB_INLINE void fp_sop2(Fp& ra, Fp& rb, const SopX<K>& xa, const SopY<K>& ya, const SopX<K>& xb,
                               const SopY<K>& yb) {
    uint32_t ea[NL], oa[NL], eb[NL], ob[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) { ea[i] = 0; oa[i] = 0; eb[i] = 0; ob[i] = 0; }
    // the rolled loop rotates its window registers with MOVs (a sixth of the instructions); LHB_SOP_UNROLL_SMALL=1 unrolls
    // the K <= 2 bodies fully to remove them — fewer instructions, but the 4 x 19 KB bodies miss the instruction cache
    constexpr int ROW_UNROLL = (LHB_SOP_UNROLL_SMALL && K <= 2) ? 6 : LHB_SOP_UNROLL_BIG;
#pragma unroll ROW_UNROLL
    for (int jp = 0; jp < NL; jp += 2) {
        uint32_t sa0[K], sa1[K], sb0[K], sb1[K];
#pragma unroll
        for (int q = 0; q < K; q++) {
            sa0[q] = ya.limb(q, jp); sa1[q] = ya.limb(q, jp + 1);
            sb0[q] = yb.limb(q, jp); sb1[q] = yb.limb(q, jp + 1);
        }
        sop_row<K>(ea, oa, xa, sa0);
        sop_row<K>(eb, ob, xb, sb0);
        sop_row<K>(oa, ea, xa, sa1);
        sop_row<K>(ob, eb, xb, sb1);
    }
    sop_finish(ra, ea, oa);
    sop_finish(rb, eb, ob);
}
// single-result form
template <int K>
LHB_HD LHB_INLINE void fp_sop1(Fp& r, const SopX<K>& x, const SopY<K>& y) {
    uint32_t e[NL], o[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) { e[i] = 0; o[i] = 0; }
#pragma unroll 1
    for (int jp = 0; jp < NL; jp += 2) {
        uint32_t s0[K], s1[K];
#pragma unroll
        for (int q = 0; q < K; q++) { s0[q] = y.limb(q, jp); s1[q] = y.limb(q, jp + 1); }
        sop_row<K>(e, o, x, s0);
        sop_row<K>(o, e, x, s1);
    }
    sop_finish(r, e, o);
}

// ------------------------------------------------------------------------------------------------ light Fp helpers
// a + b without reduction (callers guarantee the sum fits 12 limbs: a, b < 2^383)
LHB_HD LHB_INLINE void fp_add_nr(Fp& r, const Fp& a, const Fp& b) {
    add_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) addc_cc(r.v[i], a.v[i], b.v[i]);
    addc(r.v[NL - 1], a.v[NL - 1], b.v[NL - 1]);
}
// p - a  for a in [0, p]  (maps 0 to p: a valid operand < 2p wherever the result feeds a product)
LHB_HD LHB_INLINE void fp_neg_nr(Fp& r, const Fp& a) {
    sub_cc(r.v[0], FP_P.v[0], a.v[0]);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) subc_cc(r.v[i], FP_P.v[i], a.v[i]);
    subc(r.v[NL - 1], FP_P.v[NL - 1], a.v[NL - 1]);
}
// a / 2 mod p
LHB_HD LHB_INLINE void fp_half(Fp& r, const Fp& a) {
    const uint32_t m = 0u - (a.v[0] & 1u);   // odd: add p first (a + p < 2^383 is even)
    uint32_t t[NL];
    add_cc(t[0], a.v[0], FP_P.v[0] & m);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) addc_cc(t[i], a.v[i], FP_P.v[i] & m);
    addc(t[NL - 1], a.v[NL - 1], FP_P.v[NL - 1] & m);
#pragma unroll
    for (int i = 0; i < NL - 1; i++) r.v[i] = (t[i] >> 1) | (t[i + 1] << 31);
    r.v[NL - 1] = t[NL - 1] >> 1;
}

// ------------------------------------------------------------------------------------------------ columns
// A "column" is one lane's private strip of a word-interleaved shared-memory array of NT lanes (one warp's region on the
// device: NT = 32): Fp slot s, limb w of lane t is word (s * 12 + w) * NT + t, so a warp's access to the same limb of
// the same slot is one conflict-free row, and so is any permutation of lanes (a lane reading a neighbour's column).
// On the device a column is a 32-bit word offset into the kernel's dynamic shared memory, so that the out-of-line
// operations below address it with LDS/STS (a generic pointer argument would turn every access into LD/ST).
#if !defined(LHB_HOSTSIM)
extern __shared__ __align__(16) uint32_t lhb_dyn_smem[];
#endif
template <int NT>
struct Col {
#ifdef LHB_HOSTSIM
    uint32_t* p;  // &smem[tid]
    LHB_HD LHB_INLINE uint32_t* base() const { return p; }
    static LHB_HD LHB_INLINE Col make(uint32_t* smem, int tid) { return Col{smem + tid}; }
    LHB_HD LHB_INLINE Col lane(int delta) const { return Col{p + delta}; }
#else
    uint32_t off;
    LHB_HD LHB_INLINE uint32_t* base() const { return lhb_dyn_smem + off; }
    static LHB_HD LHB_INLINE Col make(uint32_t* region, int lane) { return Col{(uint32_t)(region - lhb_dyn_smem) + (uint32_t)lane}; }
    LHB_HD LHB_INLINE Col lane(int delta) const { return Col{off + (uint32_t)delta}; }
#endif
    LHB_HD LHB_INLINE const uint32_t* at(int slot) const { return base() + slot * NL * NT; }
    LHB_HD LHB_INLINE void ld(Fp& r, int slot) const {
        const uint32_t* q = base() + slot * NL * NT;
#pragma unroll
        for (int i = 0; i < NL; i++) r.v[i] = q[i * NT];
    }
    LHB_HD LHB_INLINE void st(int slot, const Fp& a) const {
        uint32_t* q = base() + slot * NL * NT;
#pragma unroll
        for (int i = 0; i < NL; i++) q[i * NT] = a.v[i];
    }
    LHB_HD LHB_INLINE void ld2(Fp2& r, int slot) const { ld(r.c0, slot); ld(r.c1, slot + 1); }
    LHB_HD LHB_INLINE void st2(int slot, const Fp2& a) const { st(slot, a.c0); st(slot + 1, a.c1); }
};

// ---- Fp2 operations on column slots (an Fp2 occupies slots s, s+1).  dst may alias any source: all reads of the
// sources complete before the first store.  `scr` is one scratch Fp slot of the same column.
// They are OUT OF LINE on purpose: their arguments are slot numbers (the operands never travel through registers or the
// stack), so a call costs a few instructions and every operation has exactly one body in the instruction cache.
// d = a * b            K = 2, X = 1 (p - u1 may equal p: X <= 1.0001), Y = 1
template <int NT>
LHB_HD LHB_INLINE void c2_mul_u(const Col<NT>& c, int d, const Fp2& u, int b) {
    SopX<2> xa, xb;
    xa.x[0] = u.c0; fp_neg_nr(xa.x[1], u.c1);
    xb.x[0] = u.c0; xb.x[1] = u.c1;
    SopY<2> ya, yb;
    ya.base[0] = c.at(b); ya.base[1] = c.at(b + 1); ya.stride = NT;   // u0 b0 + (p - u1) b1
    yb.base[0] = c.at(b + 1); yb.base[1] = c.at(b); yb.stride = NT;   // u0 b1 + u1 b0
    Fp r0, r1;
    fp_sop2<2>(r0, r1, xa, ya, xb, yb);
    c.st(d, r0); c.st(d + 1, r1);
}
template <int NT>
LHB_HD LHB_NOINLINE void c2_mul(Col<NT> c, int d, int a, int b) {
    Fp2 u;
    c.ld2(u, a);
    c2_mul_u(c, d, u, b);
}
// d = g * b with g an Fp2 parked in global memory (24 words at stride NT)
template <int NT>
LHB_HD LHB_NOINLINE void c2_mul_g(Col<NT> c, int d, const uint32_t* g, int b) {
    Fp2 u;
#pragma unroll
    for (int i = 0; i < NL; i++) { u.c0.v[i] = g[i * NT]; u.c1.v[i] = g[(NL + i) * NT]; }
    c2_mul_u(c, d, u, b);
}
// d = a^2 :  ((a0 + a1)(a0 - a1), (2 a0) a1)      K = 1, X = 2, Y = 1
template <int NT>
LHB_HD LHB_NOINLINE void c2_sqr(Col<NT> c, int d, int a, int scr) {
    Fp2 u;
    c.ld2(u, a);
    Fp dif;
    fp_sub_inl(dif, u.c0, u.c1);
    c.st(scr, dif);
    SopX<1> xa, xb;
    fp_add_nr(xa.x[0], u.c0, u.c1);
    fp_add_nr(xb.x[0], u.c0, u.c0);
    SopY<1> ya, yb;
    ya.base[0] = c.at(scr); ya.stride = NT;
    yb.base[0] = c.at(a + 1); yb.stride = NT;
    Fp r0, r1;
    fp_sop2<1>(r0, r1, xa, ya, xb, yb);
    c.st(d, r0); c.st(d + 1, r1);
}
// d = a * s, s in Fp (global / constant memory)    K = 1, X = 1, Y = 1
template <int NT>
LHB_HD LHB_NOINLINE void c2_mul_fp(Col<NT> c, int d, int a, const Fp* s) {
    SopX<1> x;
    x.x[0] = *s;
    SopY<1> ya, yb;
    ya.base[0] = c.at(a); ya.stride = NT;
    yb.base[0] = c.at(a + 1); yb.stride = NT;
    Fp r0, r1;
    fp_sop2<1>(r0, r1, x, ya, x, yb);
    c.st(d, r0); c.st(d + 1, r1);
}
template <int NT>
LHB_HD LHB_NOINLINE void c2_add(Col<NT> c, int d, int a, int b) {
    Fp x, y, o;
    c.ld(x, a); c.ld(y, b); fp_add_inl(o, x, y); c.st(d, o);
    c.ld(x, a + 1); c.ld(y, b + 1); fp_add_inl(o, x, y); c.st(d + 1, o);
}
template <int NT>
LHB_HD LHB_NOINLINE void c2_sub(Col<NT> c, int d, int a, int b) {
    Fp x, y, o;
    c.ld(x, a); c.ld(y, b); fp_sub_inl(o, x, y); c.st(d, o);
    c.ld(x, a + 1); c.ld(y, b + 1); fp_sub_inl(o, x, y); c.st(d + 1, o);
}
template <int NT>
LHB_HD LHB_NOINLINE void c2_neg(Col<NT> c, int d, int a) {
    Fp x, o;
    c.ld(x, a); fp_neg(o, x); c.st(d, o);
    c.ld(x, a + 1); fp_neg(o, x); c.st(d + 1, o);
}
template <int NT>
LHB_HD LHB_NOINLINE void c2_half(Col<NT> c, int d, int a) {
    Fp x, o;
    c.ld(x, a); fp_half(o, x); c.st(d, o);
    c.ld(x, a + 1); fp_half(o, x); c.st(d + 1, o);
}
// d = 3 a
template <int NT>
LHB_HD LHB_NOINLINE void c2_triple(Col<NT> c, int d, int a) {
    Fp x, t, o;
    c.ld(x, a); fp_add_inl(t, x, x); fp_add_inl(o, t, x); c.st(d, o);
    c.ld(x, a + 1); fp_add_inl(t, x, x); fp_add_inl(o, t, x); c.st(d + 1, o);
}
template <int NT>
LHB_HD LHB_NOINLINE void c2_copy(Col<NT> c, int d, int a) {
    Fp2 u;
    c.ld2(u, a);
    c.st2(d, u);
}
// xi * (a0 + a1 i) = (a0 - a1) + (a0 + a1) i
LHB_HD LHB_INLINE void fp2_mul_xi_inl(Fp2& r, const Fp2& a) {
    Fp t0, t1;
    fp_sub_inl(t0, a.c0, a.c1);
    fp_add_inl(t1, a.c0, a.c1);
    r.c0 = t0; r.c1 = t1;
}
// d = 12 xi a      (3 b' C of the doubling formulas, b' = 4 xi)
template <int NT>
LHB_HD LHB_NOINLINE void c2_mul_12xi(Col<NT> c, int d, int a) {
    Fp2 u, v;
    c.ld2(u, a);
    fp2_mul_xi_inl(v, u);
    Fp t;
    fp_add_inl(t, v.c0, v.c0); fp_add_inl(t, t, t); fp_add_inl(u.c0, t, t); fp_add_inl(u.c0, u.c0, t);   // 8x + 4x
    fp_add_inl(t, v.c1, v.c1); fp_add_inl(t, t, t); fp_add_inl(u.c1, t, t); fp_add_inl(u.c1, u.c1, t);
    c.st2(d, u);
}

}  // namespace bls
}  // namespace lhb200
