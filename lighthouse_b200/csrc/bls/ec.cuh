// ec.cuh — G1 (over Fp) and G2 (over Fp2) point arithmetic in Jacobian coordinates, a = 0.
// Replaces what crypto/bls/src/impls/blst.rs obtains from blst for: AggregatePublicKey::aggregate (:103),
// subgroup_check (:75), 64-bit random-scalar multiplications inside verify_multiple_aggregate_signatures (:114),
// (de)serialisation (:126-140,:190-194).
#pragma once
#include "fp2.cuh"

namespace lhb200 {
namespace bls {

// field-generic helpers
LHB_HD LHB_INLINE void f_add(Fp& r, const Fp& a, const Fp& b) { fp_add(r, a, b); }
LHB_HD LHB_INLINE void f_sub(Fp& r, const Fp& a, const Fp& b) { fp_sub(r, a, b); }
LHB_HD LHB_INLINE void f_mul(Fp& r, const Fp& a, const Fp& b) { fp_mul(r, a, b); }
LHB_HD LHB_INLINE void f_sqr(Fp& r, const Fp& a) { fp_sqr(r, a); }
LHB_HD LHB_INLINE void f_neg(Fp& r, const Fp& a) { fp_neg(r, a); }
LHB_HD LHB_INLINE bool f_is_zero(const Fp& a) { return fp_is_zero(a); }
LHB_HD LHB_INLINE bool f_eq(const Fp& a, const Fp& b) { return fp_eq(a, b); }
LHB_HD LHB_INLINE void f_set_zero(Fp& a) { fp_set_zero(a); }
LHB_HD LHB_INLINE void f_set_one(Fp& a) { a = FP_ONE; }
LHB_HD LHB_INLINE void f_inv(Fp& r, const Fp& a) { fp_inv(r, a); }
LHB_HD LHB_INLINE void f_add(Fp2& r, const Fp2& a, const Fp2& b) { fp2_add(r, a, b); }
LHB_HD LHB_INLINE void f_sub(Fp2& r, const Fp2& a, const Fp2& b) { fp2_sub(r, a, b); }
LHB_HD LHB_INLINE void f_mul(Fp2& r, const Fp2& a, const Fp2& b) { fp2_mul(r, a, b); }
LHB_HD LHB_INLINE void f_sqr(Fp2& r, const Fp2& a) { fp2_sqr(r, a); }
LHB_HD LHB_INLINE void f_neg(Fp2& r, const Fp2& a) { fp2_neg(r, a); }
LHB_HD LHB_INLINE bool f_is_zero(const Fp2& a) { return fp2_is_zero(a); }
LHB_HD LHB_INLINE bool f_eq(const Fp2& a, const Fp2& b) { return fp2_eq(a, b); }
LHB_HD LHB_INLINE void f_set_zero(Fp2& a) { fp2_set_zero(a); }
LHB_HD LHB_INLINE void f_set_one(Fp2& a) { fp2_set_one(a); }
LHB_HD LHB_INLINE void f_inv(Fp2& r, const Fp2& a) { fp2_inv(r, a); }

template <class F>
struct Affine {
    F x, y;
    uint32_t inf;  // 1 = point at infinity
};
template <class F>
struct Jac {
    F X, Y, Z;  // infinity <=> Z == 0
};
typedef Affine<Fp> G1Affine;
typedef Affine<Fp2> G2Affine;
typedef Jac<Fp> G1Jac;
typedef Jac<Fp2> G2Jac;

template <class F>
LHB_HD LHB_INLINE bool jac_is_inf(const Jac<F>& p) { return f_is_zero(p.Z); }
template <class F>
LHB_HD LHB_INLINE void jac_set_inf(Jac<F>& p) { f_set_one(p.X); f_set_one(p.Y); f_set_zero(p.Z); }
template <class F>
LHB_HD LHB_INLINE void jac_from_affine(Jac<F>& r, const Affine<F>& a) {
    if (a.inf) { jac_set_inf(r); return; }
    r.X = a.x; r.Y = a.y; f_set_one(r.Z);
}
template <class F>
LHB_HD LHB_INLINE void jac_neg(Jac<F>& r, const Jac<F>& p) { r.X = p.X; f_neg(r.Y, p.Y); r.Z = p.Z; }

// dbl-2009-l: 2M + 5S
template <class F>
LHB_HD LHB_NOINLINE void jac_dbl(Jac<F>& r, const Jac<F>& p) {
    F A, B, C, D, E, Fq, t;
    f_sqr(A, p.X);
    f_sqr(B, p.Y);
    f_sqr(C, B);
    f_add(t, p.X, B);
    f_sqr(t, t);
    f_sub(t, t, A);
    f_sub(t, t, C);
    f_add(D, t, t);            // D = 2((X+B)^2 - A - C)
    f_add(E, A, A);
    f_add(E, E, A);            // E = 3A
    f_sqr(Fq, E);
    f_mul(t, p.Y, p.Z);        // YZ (before X/Y are overwritten)
    f_add(r.Z, t, t);          // Z3 = 2YZ
    f_sub(Fq, Fq, D);
    f_sub(r.X, Fq, D);         // X3 = F - 2D
    f_sub(t, D, r.X);
    f_mul(t, E, t);
    f_add(C, C, C);
    f_add(C, C, C);
    f_add(C, C, C);            // 8C
    f_sub(r.Y, t, C);
}

// madd-2007-bl (Jacobian + affine): 7M + 4S; handles P = inf, Q = inf, P = Q, P = -Q.
template <class F>
LHB_HD LHB_NOINLINE void jac_add_affine(Jac<F>& r, const Jac<F>& p, const Affine<F>& q) {
    if (q.inf) { r = p; return; }
    if (jac_is_inf(p)) { r.X = q.x; r.Y = q.y; f_set_one(r.Z); return; }
    F Z1Z1, U2, S2, H, HH, I, J, rr, V, t;
    f_sqr(Z1Z1, p.Z);
    f_mul(U2, q.x, Z1Z1);
    f_mul(S2, q.y, p.Z);
    f_mul(S2, S2, Z1Z1);
    f_sub(H, U2, p.X);
    f_sub(rr, S2, p.Y);
    if (f_is_zero(H)) {
        if (f_is_zero(rr)) {  // same point: double
            Jac<F> qq;
            qq.X = q.x; qq.Y = q.y; f_set_one(qq.Z);
            jac_dbl(r, qq);
        } else {
            jac_set_inf(r);
        }
        return;
    }
    f_add(rr, rr, rr);         // r = 2(S2 - Y1)
    f_sqr(HH, H);
    f_add(I, HH, HH);
    f_add(I, I, I);            // I = 4HH
    f_mul(J, H, I);
    f_mul(V, p.X, I);
    f_add(t, p.Z, H);
    f_sqr(t, t);
    f_sub(t, t, Z1Z1);
    F Z3;
    f_sub(Z3, t, HH);          // Z3 = (Z1+H)^2 - Z1Z1 - HH
    f_sqr(t, rr);
    f_sub(t, t, J);
    f_sub(t, t, V);
    F X3;
    f_sub(X3, t, V);           // X3 = r^2 - J - 2V
    f_sub(t, V, X3);
    f_mul(t, rr, t);
    f_mul(J, p.Y, J);
    f_add(J, J, J);
    f_sub(r.Y, t, J);          // Y3 = r(V - X3) - 2 Y1 J
    r.X = X3;
    r.Z = Z3;
}

// add-2007-bl (Jacobian + Jacobian): 11M + 5S, all special cases handled.
template <class F>
LHB_HD LHB_NOINLINE void jac_add(Jac<F>& r, const Jac<F>& p, const Jac<F>& q) {
    if (jac_is_inf(q)) { r = p; return; }
    if (jac_is_inf(p)) { r = q; return; }
    F Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t;
    f_sqr(Z1Z1, p.Z);
    f_sqr(Z2Z2, q.Z);
    f_mul(U1, p.X, Z2Z2);
    f_mul(U2, q.X, Z1Z1);
    f_mul(S1, p.Y, q.Z);
    f_mul(S1, S1, Z2Z2);
    f_mul(S2, q.Y, p.Z);
    f_mul(S2, S2, Z1Z1);
    f_sub(H, U2, U1);
    f_sub(rr, S2, S1);
    if (f_is_zero(H)) {
        if (f_is_zero(rr)) jac_dbl(r, p);
        else jac_set_inf(r);
        return;
    }
    f_add(rr, rr, rr);
    f_add(I, H, H);
    f_sqr(I, I);               // (2H)^2
    f_mul(J, H, I);
    f_mul(V, U1, I);
    f_add(t, p.Z, q.Z);
    f_sqr(t, t);
    f_sub(t, t, Z1Z1);
    f_sub(t, t, Z2Z2);
    F Z3;
    f_mul(Z3, t, H);
    f_sqr(t, rr);
    f_sub(t, t, J);
    f_sub(t, t, V);
    F X3;
    f_sub(X3, t, V);
    f_sub(t, V, X3);
    f_mul(t, rr, t);
    f_mul(S1, S1, J);
    f_add(S1, S1, S1);
    f_sub(r.Y, t, S1);
    r.X = X3;
    r.Z = Z3;
}

template <class F>
LHB_HD LHB_NOINLINE void jac_to_affine(Affine<F>& r, const Jac<F>& p) {
    if (jac_is_inf(p)) { f_set_zero(r.x); f_set_zero(r.y); r.inf = 1; return; }
    F zi, zi2;
    f_inv(zi, p.Z);
    f_sqr(zi2, zi);
    f_mul(r.x, p.X, zi2);
    f_mul(zi2, zi2, zi);
    f_mul(r.y, p.Y, zi2);
    r.inf = 0;
}

// [k]P, k given as little-endian 32-bit words, `nbits` significant bits (left-to-right double-and-add).
template <class F>
LHB_HD LHB_NOINLINE void jac_mul_affine(Jac<F>& r, const Affine<F>& p, const uint32_t* k, int nbits) {
    Jac<F> acc;
    jac_set_inf(acc);
    for (int i = nbits - 1; i >= 0; i--) {
        jac_dbl(acc, acc);
        if ((k[i >> 5] >> (i & 31)) & 1) jac_add_affine(acc, acc, p);
    }
    r = acc;
}
template <class F>
LHB_HD LHB_NOINLINE void jac_mul(Jac<F>& r, const Jac<F>& p, const uint32_t* k, int nbits) {
    Jac<F> acc;
    jac_set_inf(acc);
    for (int i = nbits - 1; i >= 0; i--) {
        jac_dbl(acc, acc);
        if ((k[i >> 5] >> (i & 31)) & 1) jac_add(acc, acc, p);
    }
    r = acc;
}

// [|x|]P for the curve parameter |x| = 0xd201000000010000 (bits 63,62,60,57,48,16)
constexpr uint64_t BLS_X_ABS = 0xd201000000010000ull;
template <class F>
LHB_HD LHB_NOINLINE void jac_mul_x_abs(Jac<F>& r, const Jac<F>& p) {
    Jac<F> acc = p;
    for (int i = 62; i >= 0; i--) {
        jac_dbl(acc, acc);
        if ((BLS_X_ABS >> i) & 1) jac_add(acc, acc, p);
    }
    r = acc;
}

// on-curve tests (affine)
LHB_HD LHB_INLINE bool g1_on_curve(const G1Affine& p) {
    if (p.inf) return true;
    Fp l, r;
    fp_sqr(l, p.y);
    fp_sqr(r, p.x);
    fp_mul(r, r, p.x);
    fp_add(r, r, G1_B);
    return fp_eq(l, r);
}
LHB_HD LHB_INLINE bool g2_on_curve(const G2Affine& p) {
    if (p.inf) return true;
    Fp2 l, r;
    fp2_sqr(l, p.y);
    fp2_sqr(r, p.x);
    fp2_mul(r, r, p.x);
    fp2_add(r, r, G2_B);
    return fp2_eq(l, r);
}

// psi endomorphism on E2 (untwist-Frobenius-twist), Jacobian form; psi^2 = (x * PSI2_CX, -y)
LHB_HD LHB_INLINE void g2_psi(G2Jac& r, const G2Jac& p) {
    Fp2 t;
    fp2_conj(t, p.X); fp2_mul(r.X, t, PSI_CX);
    fp2_conj(t, p.Y); fp2_mul(r.Y, t, PSI_CY);
    fp2_conj(r.Z, p.Z);
}
LHB_HD LHB_INLINE void g2_psi2(G2Jac& r, const G2Jac& p) {
    fp2_mul_fp(r.X, p.X, PSI2_CX);
    fp2_neg(r.Y, p.Y);
    r.Z = p.Z;
}
// Jacobian equality (projective comparison)
template <class F>
LHB_HD LHB_INLINE bool jac_eq(const Jac<F>& a, const Jac<F>& b) {
    const bool ia = jac_is_inf(a), ib = jac_is_inf(b);
    if (ia || ib) return ia && ib;
    F za, zb, l, r;
    f_sqr(za, a.Z);
    f_sqr(zb, b.Z);
    f_mul(l, a.X, zb);
    f_mul(r, b.X, za);
    if (!f_eq(l, r)) return false;
    f_mul(za, za, a.Z);
    f_mul(zb, zb, b.Z);
    f_mul(l, a.Y, zb);
    f_mul(r, b.Y, za);
    return f_eq(l, r);
}

// G1 membership (key_validate, blst.rs:130-140 -> blst's POINTonE1_in_G1) without the 255-bit [r]P: with
// phi(x, y) = (beta x, y) acting on G1 as multiplication by -x^2, P is in G1 <=> phi(P) == -[x^2]P (Scott 2021, sec. 6);
// two sparse 64-bit multiplications (126 doublings + 10 additions) instead of 255 doublings + ~128 additions.
// The [x]P == P guard rejects the points of E(Fp) on which [x] is the identity before the comparison.
LHB_HD LHB_NOINLINE bool g1_in_subgroup(const G1Affine& a) {
    if (a.inf) return true;
    G1Jac p, t;
    jac_from_affine(p, a);
    jac_mul_x_abs(t, p);                  // [|x|]P
    if (jac_eq(t, p)) return false;
    jac_mul_x_abs(t, t);                  // [x^2]P
    jac_neg(t, t);
    G1Jac e = p;
    fp_mul(e.X, p.X, G1_BETA);            // phi(P)
    return jac_eq(t, e);
}

// [r]P for a 64-bit r, SIMT-friendly: right-to-left over signed base-4 digits d_j in {-2,-1,0,1} with two
// buckets (B1 collects +-4^j P for |d_j| = 1, B2 for |d_j| = 2), result B1 + 2 B2.  Every lane performs the same
// 33 additions at the same program points (only the target bucket / sign differ), so a warp does 33 + 2 additions
// instead of the ~64 a divergent double-and-add costs it.  `on_bit(i, D)` is called with D = 2^i P for i = 0..63
// (k_sig_prepare uses it to collect [|x|]P from the same doublings).
template <class F, class OnBit>
LHB_HD LHB_INLINE void jac_mul_u64_buckets(Jac<F>& out, const Jac<F>& p, uint64_t r, OnBit on_bit) {
    Jac<F> D = p, T, B1, B2;
    jac_set_inf(B1);
    jac_set_inf(B2);
    uint32_t carry = 0;
    for (int j = 0; j <= 32; j++) {
        const uint32_t v = (j < 32 ? (uint32_t)((r >> (2 * j)) & 3u) : 0u) + carry;
        const int d = v >= 2 ? (int)v - 4 : (int)v;
        carry = v >= 2 ? 1u : 0u;
        if (d != 0) {
            T = D;
            if (d < 0) f_neg(T.Y, D.Y);
            Jac<F>* tgt = (d == 2 || d == -2) ? &B2 : &B1;
            jac_add(*tgt, *tgt, T);
        }
        if (j < 32) {
            on_bit(2 * j, D);
            jac_dbl(D, D);
            on_bit(2 * j + 1, D);
            jac_dbl(D, D);
        }
    }
    jac_dbl(B2, B2);
    jac_add(out, B1, B2);
}
struct NoOnBit {
    template <class J>
    LHB_HD void operator()(int, const J&) const {}
};
template <class F>
LHB_HD LHB_NOINLINE void jac_mul_u64(Jac<F>& out, const Jac<F>& p, uint64_t r) {
    jac_mul_u64_buckets(out, p, r, NoOnBit());
}

// One pass computing BOTH [r]P and [|x|]P from the same 64 doublings of P (k_sig_prepare: r*sig for the batch sum,
// [x]sig for the subgroup check).
struct CollectX {
    G2Jac* acc;
    LHB_HD void operator()(int i, const G2Jac& D) const {
        if ((BLS_X_ABS >> i) & 1) jac_add(*acc, *acc, D);
    }
};
LHB_HD LHB_NOINLINE void g2_mul_r_and_x(G2Jac& out_r, G2Jac& out_x, const G2Affine& p, uint64_t r) {
    G2Jac pj;
    jac_from_affine(pj, p);
    jac_set_inf(out_x);
    CollectX cx;
    cx.acc = &out_x;
    jac_mul_u64_buckets(out_r, pj, r, cx);
}

// P in G2  <=>  psi(P) == [x]P = -[|x|]P   (blst.rs:75 subgroup_check).
// The point at infinity passes (Appendix C item 3).
LHB_HD LHB_NOINLINE bool g2_in_subgroup(const G2Affine& p) {
    if (p.inf) return true;
    G2Jac pj, xp, ps;
    jac_from_affine(pj, p);
    jac_mul_x_abs(xp, pj);
    jac_neg(xp, xp);
    g2_psi(ps, pj);
    return jac_eq(ps, xp);
}

// Budroni–Pintore cofactor clearing, equals [h_eff]P (RFC 9380 App. G.3).
LHB_HD LHB_NOINLINE void g2_clear_cofactor(G2Jac& r, const G2Jac& p) {
    G2Jac t1, t2, t3, np;
    jac_mul_x_abs(t1, p);
    jac_neg(t1, t1);            // t1 = [x]P
    g2_psi(t2, p);              // t2 = psi(P)
    jac_dbl(t3, p);
    g2_psi2(t3, t3);            // t3 = psi^2(2P)
    G2Jac n2;
    jac_neg(n2, t2);
    jac_add(t3, t3, n2);        // t3 -= t2
    jac_add(t2, t1, t2);        // t2 = t1 + t2
    jac_mul_x_abs(t2, t2);
    jac_neg(t2, t2);            // t2 = [x] t2
    jac_add(t3, t3, t2);
    jac_neg(n2, t1);
    jac_add(t3, t3, n2);        // t3 -= t1
    jac_neg(np, p);
    jac_add(r, t3, np);         // Q = t3 - P
}

// ------------------------------------------------------------------------------------------------ serialisation
// ZCash format (crypto/bls/src/generic_public_key.rs:12-21, generic_signature.rs:15-26).
enum DecodeStatus : int32_t { DEC_OK = 0, DEC_INFINITY = 1, DEC_BAD = 2 };

// 96-byte uncompressed affine G1 (x || y big-endian), as persisted by validator_pubkey_cache.rs:195-199.
// Encoding checks as blst's P1 deserialize does for the 96-byte form (blst.rs:142-150): the compression and sort flags
// must be clear, an infinity encoding must be all zero otherwise.  No curve/subgroup validation here (keys were
// validated when the cache imported them; k_table_import / lhb200_g1_deserialize_uncompressed add the curve check).
LHB_HD LHB_INLINE int32_t g1_from_uncompressed(G1Affine& r, const uint8_t* b) {
    if (b[0] & 0xa0) return DEC_BAD;
    if (b[0] & 0x40) {
        uint32_t nz = b[0] & 0x3f;
        for (int i = 1; i < 96; i++) nz |= b[i];
        if (nz) return DEC_BAD;
        f_set_zero(r.x); f_set_zero(r.y); r.inf = 1;
        return DEC_INFINITY;
    }
    Fp cx, cy;
    fp_from_be48(cx, b);
    fp_from_be48(cy, b + 48);
    cx.v[NL - 1] &= 0x1fffffffu;
    if (!fp_canon_lt_p(cx) || !fp_canon_lt_p(cy)) return DEC_BAD;
    fp_to_mont(r.x, cx);
    fp_to_mont(r.y, cy);
    r.inf = 0;
    return DEC_OK;
}
LHB_HD LHB_INLINE void g1_to_uncompressed(uint8_t* b, const G1Affine& p) {
    if (p.inf) { for (int i = 0; i < 96; i++) b[i] = 0; b[0] = 0x40; return; }
    Fp c;
    fp_from_mont(c, p.x); fp_to_be48(b, c);
    fp_from_mont(c, p.y); fp_to_be48(b + 48, c);
}
LHB_HD LHB_INLINE void g1_compress(uint8_t* b, const G1Affine& p) {
    if (p.inf) { for (int i = 0; i < 48; i++) b[i] = 0; b[0] = 0xc0; return; }
    Fp cx, cy;
    fp_from_mont(cx, p.x);
    fp_from_mont(cy, p.y);
    fp_to_be48(b, cx);
    b[0] |= 0x80 | (fp_canon_gt_half(cy) ? 0x20 : 0);
}
// 48-byte compressed G1 -> affine (sqrt), on-curve by construction; no subgroup check here.
LHB_HD LHB_NOINLINE int32_t g1_decompress(G1Affine& r, const uint8_t* b) {
    const uint32_t c = b[0] >> 7, inf = (b[0] >> 6) & 1, s = (b[0] >> 5) & 1;
    if (!c) return DEC_BAD;
    Fp cx;
    fp_from_be48(cx, b);
    cx.v[NL - 1] &= 0x1fffffffu;
    if (inf) {
        if (!fp_is_zero(cx) || s) return DEC_BAD;
        f_set_zero(r.x); f_set_zero(r.y); r.inf = 1;
        return DEC_INFINITY;
    }
    if (!fp_canon_lt_p(cx)) return DEC_BAD;
    fp_to_mont(r.x, cx);
    Fp rhs, y;
    fp_sqr(rhs, r.x);
    fp_mul(rhs, rhs, r.x);
    fp_add(rhs, rhs, G1_B);
    if (!fp_sqrt(y, rhs)) return DEC_BAD;
    Fp cy;
    fp_from_mont(cy, y);
    if ((fp_canon_gt_half(cy) ? 1u : 0u) != s) fp_neg(y, y);
    r.y = y;
    r.inf = 0;
    return DEC_OK;
}
LHB_HD LHB_INLINE bool fp2_lex_larger(const Fp2& y) {  // ZCash sign of y in Fp2: compare c1 first, then c0
    Fp c0, c1;
    fp_from_mont(c0, y.c0);
    fp_from_mont(c1, y.c1);
    return fp_is_zero(c1) ? fp_canon_gt_half(c0) : fp_canon_gt_half(c1);
}
LHB_HD LHB_INLINE void g2_compress(uint8_t* b, const G2Affine& p) {
    if (p.inf) { for (int i = 0; i < 96; i++) b[i] = 0; b[0] = 0xc0; return; }
    Fp c;
    fp_from_mont(c, p.x.c1); fp_to_be48(b, c);
    fp_from_mont(c, p.x.c0); fp_to_be48(b + 48, c);
    b[0] |= 0x80 | (fp2_lex_larger(p.y) ? 0x20 : 0);
}
// 96-byte compressed G2 (x.c1 || x.c0) -> affine; no subgroup check (Signature::deserialize, blst.rs:192-194).
LHB_HD LHB_NOINLINE int32_t g2_decompress(G2Affine& r, const uint8_t* b) {
    const uint32_t c = b[0] >> 7, inf = (b[0] >> 6) & 1, s = (b[0] >> 5) & 1;
    if (!c) return DEC_BAD;
    Fp c1, c0;
    fp_from_be48(c1, b);
    fp_from_be48(c0, b + 48);
    c1.v[NL - 1] &= 0x1fffffffu;
    if (inf) {
        if (!fp_is_zero(c1) || !fp_is_zero(c0) || s) return DEC_BAD;
        f_set_zero(r.x); f_set_zero(r.y); r.inf = 1;
        return DEC_INFINITY;
    }
    if (!fp_canon_lt_p(c1) || !fp_canon_lt_p(c0)) return DEC_BAD;
    fp_to_mont(r.x.c0, c0);
    fp_to_mont(r.x.c1, c1);
    Fp2 rhs, y;
    fp2_sqr(rhs, r.x);
    fp2_mul(rhs, rhs, r.x);
    fp2_add(rhs, rhs, G2_B);
    if (!fp2_sqrt(y, rhs)) return DEC_BAD;
    if ((fp2_lex_larger(y) ? 1u : 0u) != s) fp2_neg(y, y);
    r.y = y;
    r.inf = 0;
    return DEC_OK;
}

}  // namespace bls
}  // namespace lhb200
