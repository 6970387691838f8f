// h2c.cuh — hash_to_G2 for BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_ (RFC 9380 §8.8.2), the message hashing
// blst performs inside verify_multiple_aggregate_signatures (crypto/bls/src/impls/blst.rs:114, DST at :15).
//
// Inversion-free SSWU: g(x1) is kept as a fraction N/D; one Fp exponentiation t = norm(N D)^((p-3)/4) yields the
// quadratic-residue test, the norm root for the Fp2 square root AND 1/norm (as t^4 * norm), so the affine (x, y)
// that sgn0 needs costs two exponentiations per map and no separate inversion.
#pragma once
#include "../sha256.cuh"
#include "ec.cuh"

namespace lhb200 {
namespace bls {

// DST' = DST || len(DST)  (43 + 1 bytes)
LHB_HD LHB_INLINE void dst_prime(uint8_t* d) {
    const char* s = "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_";
    for (int i = 0; i < 43; i++) d[i] = (uint8_t)s[i];
    d[43] = 43;
}

// expand_message_xmd(msg32, DST, 256) -> 256 uniform bytes
LHB_HD LHB_NOINLINE void expand_message_xmd_256(const uint8_t* msg32, uint8_t* out256) {
    uint8_t buf[144], b0[32];
    // b0 = H(Z_pad(64) || msg || I2OSP(256, 2) || 0x00 || DST')
    for (int i = 0; i < 64; i++) buf[i] = 0;
    for (int i = 0; i < 32; i++) buf[64 + i] = msg32[i];
    buf[96] = 0x01; buf[97] = 0x00; buf[98] = 0x00;
    dst_prime(buf + 99);
    sha256_short(buf, 143, b0);
    // b1 = H(b0 || 0x01 || DST'), b_i = H((b0 ^ b_{i-1}) || i || DST')
    uint8_t m[80];
    dst_prime(m + 33);
    for (int i = 0; i < 32; i++) m[i] = b0[i];
    m[32] = 1;
    sha256_short(m, 77, out256);
    for (int k = 2; k <= 8; k++) {
        for (int i = 0; i < 32; i++) m[i] = b0[i] ^ out256[32 * (k - 2) + i];
        m[32] = (uint8_t)k;
        sha256_short(m, 77, out256 + 32 * (k - 1));
    }
}

// 64 big-endian bytes -> Fp (Montgomery) = int(bytes) mod p
LHB_HD LHB_INLINE void fp_from_be64_mod(Fp& r, const uint8_t* b) {
    Fp hi, lo, t;
    fp_set_zero(hi);
    fp_set_zero(lo);
    for (int i = 0; i < 8; i++) {
        const uint8_t* qh = b + 4 * (7 - i);
        const uint8_t* ql = b + 32 + 4 * (7 - i);
        hi.v[i] = ((uint32_t)qh[0] << 24) | ((uint32_t)qh[1] << 16) | ((uint32_t)qh[2] << 8) | qh[3];
        lo.v[i] = ((uint32_t)ql[0] << 24) | ((uint32_t)ql[1] << 16) | ((uint32_t)ql[2] << 8) | ql[3];
    }
    fp_mul(t, hi, FP_R2_256);  // hi * 2^256 * R
    fp_mul(lo, lo, FP_R2);     // lo * R
    fp_add(r, t, lo);
}

// hash_to_field(msg, count = 2) over Fp2
LHB_HD LHB_NOINLINE void hash_to_field_fp2(Fp2& u0, Fp2& u1, const uint8_t* msg32) {
    uint8_t uni[256];
    expand_message_xmd_256(msg32, uni);
    fp_from_be64_mod(u0.c0, uni);
    fp_from_be64_mod(u0.c1, uni + 64);
    fp_from_be64_mod(u1.c0, uni + 128);
    fp_from_be64_mod(u1.c1, uni + 192);
}

LHB_HD LHB_INLINE void fp2_norm(Fp& n, const Fp2& a) {
    Fp t;
    fp_sqr(n, a.c0);
    fp_sqr(t, a.c1);
    fp_add(n, n, t);
}

// Simplified SWU onto E2': y^2 = x^3 + A'x + B'.  Output affine (x, y).
LHB_HD LHB_NOINLINE void map_to_curve_sswu(Fp2& x, Fp2& y, const Fp2& u, Fp2* trace = nullptr) {
    Fp2 tv1, tv2, x1n, x1d, xd2, D, N, t, a;
    fp2_sqr(tv1, u);
    fp2_mul(tv1, tv1, SSWU_Z);           // Z u^2
    fp2_sqr(tv2, tv1);
    fp2_add(tv2, tv2, tv1);              // Z^2 u^4 + Z u^2
    Fp2 one;
    fp2_set_one(one);
    fp2_add(x1n, tv2, one);
    fp2_mul(x1n, x1n, SSWU_B);           // B (tv2 + 1)
    if (fp2_is_zero(tv2)) {
        x1d = SSWU_ZA;                   // exceptional case: x1 = B / (Z A)
    } else {
        fp2_mul(x1d, tv2, SSWU_A);
        fp2_neg(x1d, x1d);               // -A tv2
    }
    fp2_sqr(xd2, x1d);
    fp2_mul(D, xd2, x1d);                // D = x1d^3
    fp2_sqr(N, x1n);
    fp2_mul(t, xd2, SSWU_A);
    fp2_add(N, N, t);
    fp2_mul(N, N, x1n);
    fp2_mul(t, D, SSWU_B);
    fp2_add(N, N, t);                    // N = x1n^3 + A x1n x1d^2 + B x1d^3
    fp2_mul(a, N, D);                    // g(x1) = N/D is a square  <=>  a = N D is
    Fp na, t1, s, chk;
    fp2_norm(na, a);
    fp_pow_pm3d4(t1, na);
    fp_mul(s, na, t1);                   // na^((p+1)/4)
    fp_sqr(chk, s);
    const bool is_sq = fp_eq(chk, na);   // a square in Fp2 <=> norm(a) square in Fp
    Fp2 target = a;
    if (!is_sq) {
        fp2_mul(target, a, SSWU_Z);      // Z a is a square; norm(Z a) = 5 na, sqrt = sqrt(-5) * sqrt(-na)
        fp_mul(s, s, FP_SQRT_M5);
    }
    Fp2 y0;
    fp2_sqrt_with_norm_root(y0, target, s);
    // 1/D = conj(D) * norm(N) / norm(a),  1/norm(a) = t1^4 * na
    Fp inv_na, nN;
    fp_sqr(inv_na, t1);
    fp_sqr(inv_na, inv_na);
    fp_mul(inv_na, inv_na, na);
    fp2_norm(nN, N);
    fp_mul(inv_na, inv_na, nN);
    Fp2 invD;
    fp2_conj(invD, D);
    fp2_mul_fp(invD, invD, inv_na);
    Fp2 y1, x1;
    fp2_mul(y1, y0, invD);               // sqrt(N/D) or sqrt(Z N/D)
    fp2_mul(x1, x1n, xd2);
    fp2_mul(x1, x1, invD);               // x1n / x1d
    if (is_sq) {
        x = x1;
        y = y1;
    } else {
        fp2_mul(x, tv1, x1);             // x2 = Z u^2 x1
        fp2_mul(y, tv1, u);
        fp2_mul(y, y, y1);               // y2 = Z u^3 sqrt(Z g(x1))
    }
    if (trace) {  // stage probe for the parity tests
        trace[0] = tv1; trace[1] = tv2; trace[2] = x1n; trace[3] = x1d; trace[4] = N; trace[5] = D; trace[6] = a;
        trace[7].c0 = na; trace[7].c1 = t1; trace[8].c0 = s; trace[8].c1 = inv_na; trace[9] = target; trace[10] = y0;
        trace[11] = invD; trace[12] = y1; trace[13] = x1; trace[14] = x; trace[15] = y;
    }
    if (fp2_sgn0(u) != fp2_sgn0(y)) fp2_neg(y, y);
}

// 3-isogeny E2' -> E2, output Jacobian with Z = xden * yden (no inversion).
LHB_HD LHB_NOINLINE void iso_map_g2(G2Jac& r, const Fp2& x, const Fp2& y) {
    Fp2 xn, xd, yn, yd, t;
    // Horner
    fp2_mul(xn, ISO_XNUM[3], x); fp2_add(xn, xn, ISO_XNUM[2]);
    fp2_mul(xn, xn, x); fp2_add(xn, xn, ISO_XNUM[1]);
    fp2_mul(xn, xn, x); fp2_add(xn, xn, ISO_XNUM[0]);
    fp2_add(xd, x, ISO_XDEN[1]);                     // monic degree 2
    fp2_mul(xd, xd, x); fp2_add(xd, xd, ISO_XDEN[0]);
    fp2_mul(yn, ISO_YNUM[3], x); fp2_add(yn, yn, ISO_YNUM[2]);
    fp2_mul(yn, yn, x); fp2_add(yn, yn, ISO_YNUM[1]);
    fp2_mul(yn, yn, x); fp2_add(yn, yn, ISO_YNUM[0]);
    fp2_add(yd, x, ISO_YDEN[2]);                     // monic degree 3
    fp2_mul(yd, yd, x); fp2_add(yd, yd, ISO_YDEN[1]);
    fp2_mul(yd, yd, x); fp2_add(yd, yd, ISO_YDEN[0]);
    // Z = xd yd ; X = xn xd yd^2 ; Y = y yn xd^3 yd^2
    Fp2 yd2, xd2;
    fp2_mul(r.Z, xd, yd);
    fp2_sqr(yd2, yd);
    fp2_mul(t, xn, xd);
    fp2_mul(r.X, t, yd2);
    fp2_sqr(xd2, xd);
    fp2_mul(t, xd2, xd);
    fp2_mul(t, t, yd2);
    fp2_mul(t, t, yn);
    fp2_mul(r.Y, t, y);
}

// hash_to_curve(msg) in Jacobian coordinates
LHB_HD LHB_NOINLINE void hash_to_g2_jac(G2Jac& r, const uint8_t* msg32) {
    Fp2 u0, u1, x, y;
    hash_to_field_fp2(u0, u1, msg32);
    G2Jac q0, q1;
    map_to_curve_sswu(x, y, u0);
    iso_map_g2(q0, x, y);
    map_to_curve_sswu(x, y, u1);
    iso_map_g2(q1, x, y);
    jac_add(q0, q0, q1);
    g2_clear_cofactor(r, q0);
}

}  // namespace bls
}  // namespace lhb200
