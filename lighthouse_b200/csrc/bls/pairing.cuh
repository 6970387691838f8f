// pairing.cuh — optimal ate Miller loop and final exponentiation on BLS12-381
// (the pairing work inside blst's verify_multiple_aggregate_signatures, crypto/bls/src/impls/blst.rs:114-118).
//
// Line functions are kept in the sparse form  l = c0 + c1 v + c4 (v w)  ("014") obtained by multiplying the
// untwisted line by w^3 (an element of the proper subfield Fp4, killed by the final exponentiation):
//     l(P) = (lam x_T - y_T) - lam x_P w^2 + y_P w^3 ,   w^2 = v, w^3 = v w.
// T is tracked in Jacobian coordinates; all denominators are Fp2 factors, likewise killed.
// The G1 argument may be given projectively as (px, py, pz) = (X Z, Y, Z^3) of a Jacobian point so that no
// field inversion is needed after the 64-bit scalar multiplication r * apk.
#pragma once
#include "ec.cuh"

namespace lhb200 {
namespace bls {

struct G1Proj3 {  // evaluation point scaled by Z^3:  x_P = px / pz, y_P = py / pz  (pz in Fp, killed by final exp)
    Fp px, py, pz;
};
LHB_HD LHB_INLINE void g1proj3_from_jac(G1Proj3& r, const G1Jac& p) {
    Fp z2;
    fp_mul(r.px, p.X, p.Z);
    r.py = p.Y;
    fp_sqr(z2, p.Z);
    fp_mul(r.pz, z2, p.Z);
}
LHB_HD LHB_INLINE void g1proj3_from_affine(G1Proj3& r, const G1Affine& p) {
    r.px = p.x; r.py = p.y; r.pz = FP_ONE;
}

// Doubling step: T <- 2T, returns line coefficients (scaled by 2 Y Z^3 and by pz).
LHB_HD LHB_NOINLINE void miller_dbl_step(G2Jac& T, Fp2& c0, Fp2& c1, Fp2& c4, const G1Proj3& P) {
    Fp2 A, B, C, D, E, Fq, ZZ, t, Z3;
    fp2_sqr(A, T.X);
    fp2_sqr(B, T.Y);
    fp2_sqr(ZZ, T.Z);
    fp2_sqr(C, B);
    fp2_add(t, T.X, B);
    fp2_sqr(t, t);
    fp2_sub(t, t, A);
    fp2_sub(t, t, C);
    fp2_add(D, t, t);
    fp2_add(E, A, A);
    fp2_add(E, E, A);                 // 3 X^2
    fp2_sqr(Fq, E);
    fp2_add(t, T.Y, T.Z);
    fp2_sqr(t, t);
    fp2_sub(t, t, B);
    fp2_sub(Z3, t, ZZ);               // 2 Y Z
    // line: c0 = 3X^3 - 2Y^2 ; c1 = -(3X^2 Z^2) x_P ; c4 = (Z3 Z^2) y_P   [times pz for a projective P]
    fp2_mul(c0, E, T.X);
    fp2_sub(c0, c0, B);
    fp2_sub(c0, c0, B);
    fp2_mul(c1, E, ZZ);
    fp2_neg(c1, c1);
    fp2_mul(c4, Z3, ZZ);
    fp2_mul_fp(c0, c0, P.pz);
    fp2_mul_fp(c1, c1, P.px);
    fp2_mul_fp(c4, c4, P.py);
    // point
    fp2_sub(Fq, Fq, D);
    fp2_sub(T.X, Fq, D);
    fp2_sub(t, D, T.X);
    fp2_mul(t, E, t);
    fp2_add(C, C, C);
    fp2_add(C, C, C);
    fp2_add(C, C, C);
    fp2_sub(T.Y, t, C);
    T.Z = Z3;
}

// Addition step: T <- T + Q (Q affine, Q != +-T), line through T and Q (scaled by Z3 = 2 Z H and by pz).
LHB_HD LHB_NOINLINE void miller_add_step(G2Jac& T, Fp2& c0, Fp2& c1, Fp2& c4, const G2Affine& Q, const G1Proj3& P) {
    Fp2 Z1Z1, U2, S2, H, HH, I, J, rr, V, t, Z3, X3;
    fp2_sqr(Z1Z1, T.Z);
    fp2_mul(U2, Q.x, Z1Z1);
    fp2_mul(S2, Q.y, T.Z);
    fp2_mul(S2, S2, Z1Z1);
    fp2_sub(H, U2, T.X);
    fp2_sub(rr, S2, T.Y);
    fp2_add(rr, rr, rr);
    fp2_sqr(HH, H);
    fp2_add(I, HH, HH);
    fp2_add(I, I, I);
    fp2_mul(J, H, I);
    fp2_mul(V, T.X, I);
    fp2_add(t, T.Z, H);
    fp2_sqr(t, t);
    fp2_sub(t, t, Z1Z1);
    fp2_sub(Z3, t, HH);
    fp2_sqr(t, rr);
    fp2_sub(t, t, J);
    fp2_sub(t, t, V);
    fp2_sub(X3, t, V);
    fp2_sub(t, V, X3);
    fp2_mul(t, rr, t);
    fp2_mul(J, T.Y, J);
    fp2_add(J, J, J);
    fp2_sub(T.Y, t, J);
    T.X = X3;
    T.Z = Z3;
    // line: c0 = r x_Q - y_Q Z3 ; c1 = -r x_P ; c4 = Z3 y_P
    fp2_mul(c0, rr, Q.x);
    fp2_mul(t, Q.y, Z3);
    fp2_sub(c0, c0, t);
    fp2_neg(c1, rr);
    c4 = Z3;
    fp2_mul_fp(c0, c0, P.pz);
    fp2_mul_fp(c1, c1, P.px);
    fp2_mul_fp(c4, c4, P.py);
}

// f_{|x|,Q}(P), conjugated because x < 0.  Q affine and not at infinity; P not at infinity.
LHB_HD LHB_NOINLINE void miller_loop(Fp12& f, const G1Proj3& P, const G2Affine& Q) {
    G2Jac T;
    jac_from_affine(T, Q);
    Fp2 c0, c1, c4;
    fp12_set_one(f);
    bool first = true;
    for (int i = 62; i >= 0; i--) {
        if (!first) fp12_sqr(f, f);
        miller_dbl_step(T, c0, c1, c4, P);
        if (first) {  // f = 1: f^2 * l = l
            fp6_set_zero(f.c0); fp6_set_zero(f.c1);
            f.c0.c0 = c0; f.c0.c1 = c1; f.c1.c1 = c4;
            first = false;
        } else {
            fp12_mul_by_014(f, f, c0, c1, c4);
        }
        if ((BLS_X_ABS >> i) & 1) {
            miller_add_step(T, c0, c1, c4, Q, P);
            fp12_mul_by_014(f, f, c0, c1, c4);
        }
    }
    fp12_conj(f, f);
}

// Addition step with Q in JACOBIAN coordinates (X2, Y2, Z2): T <- T + Q (general addition, add-2007-bl) and the line
// through T and Q.  With lambda = r / Z3 (r = 2(S2 - S1), Z3 = 2 Z1 Z2 H) and x_Q = X2/Z2^2, y_Q = Y2/Z2^3, the line
// scaled by Z3 Z2^3 (an Fp2 factor, killed by the final exponentiation) is
//     c0 = r X2 Z2 - Y2 Z3 ,   c1 = -(r Z2^3) x_P ,   c4 = (Z3 Z2^3) y_P          [times pz for a projective P].
// This lets hash_to_G2 hand its result over without the to-affine inversion (one Fp exponentiation per set).
LHB_HD LHB_NOINLINE void miller_add_step_jac(G2Jac& T, Fp2& c0, Fp2& c1, Fp2& c4, const G2Jac& Q, const G1Proj3& P) {
    Fp2 Z1Z1, Z2Z2, Z2c, U1, U2, S1, S2, H, I, J, rr, V, t, Z3, X3;
    fp2_sqr(Z1Z1, T.Z);
    fp2_sqr(Z2Z2, Q.Z);
    fp2_mul(Z2c, Z2Z2, Q.Z);          // Z2^3
    fp2_mul(U1, T.X, Z2Z2);
    fp2_mul(U2, Q.X, Z1Z1);
    fp2_mul(S1, T.Y, Z2c);
    fp2_mul(S2, Q.Y, T.Z);
    fp2_mul(S2, S2, Z1Z1);
    fp2_sub(H, U2, U1);
    fp2_sub(rr, S2, S1);
    fp2_add(rr, rr, rr);
    fp2_add(I, H, H);
    fp2_sqr(I, I);
    fp2_mul(J, H, I);
    fp2_mul(V, U1, I);
    fp2_add(t, T.Z, Q.Z);
    fp2_sqr(t, t);
    fp2_sub(t, t, Z1Z1);
    fp2_sub(t, t, Z2Z2);
    fp2_mul(Z3, t, H);
    fp2_sqr(t, rr);
    fp2_sub(t, t, J);
    fp2_sub(t, t, V);
    fp2_sub(X3, t, V);
    fp2_sub(t, V, X3);
    fp2_mul(t, rr, t);
    fp2_mul(J, S1, J);
    fp2_add(J, J, J);
    fp2_sub(T.Y, t, J);
    T.X = X3;
    T.Z = Z3;
    // line
    fp2_mul(t, Q.X, Q.Z);
    fp2_mul(c0, rr, t);
    fp2_mul(t, Q.Y, Z3);
    fp2_sub(c0, c0, t);
    fp2_mul(c1, rr, Z2c);
    fp2_neg(c1, c1);
    fp2_mul(c4, Z3, Z2c);
    fp2_mul_fp(c0, c0, P.pz);
    fp2_mul_fp(c1, c1, P.px);
    fp2_mul_fp(c4, c4, P.py);
}

// Multi-pairing Miller loop: prod_j f_{|x|,Q_j}(P_j) for m <= MILLER_KMAX pairs with ONE Fp12 squaring per bit shared by
// all of them ((prod f_j)^2 prod l_j == prod (f_j^2 l_j) exactly in Fp12) — the shape of blst's miller_loop_n behind
// verify_multiple_aggregate_signatures (crypto/bls/src/impls/blst.rs:114-118).  P and Q are read in place; Q is
// Jacobian (what hash_to_G2 produces), so the value equals the affine loop's up to Fp2 factors the final
// exponentiation kills.
constexpr int MILLER_KMAX = 4;
LHB_HD LHB_NOINLINE void miller_loop_multi(Fp12& f, const G1Proj3* P, const G2Jac* Q, const uint32_t* idx, int m) {
    G2Jac T[MILLER_KMAX];
#pragma unroll 1
    for (int j = 0; j < m; j++) T[j] = Q[idx[j]];
    Fp2 c0, c1, c4;
#pragma unroll 1
    for (int i = 62; i >= 0; i--) {
        if (i != 62) fp12_sqr(f, f);
#pragma unroll 1
        for (int j = 0; j < m; j++) {
            miller_dbl_step(T[j], c0, c1, c4, P[idx[j]]);
            if (i == 62 && j == 0) {  // f = 1: f^2 * l = l
                fp6_set_zero(f.c0); fp6_set_zero(f.c1);
                f.c0.c0 = c0; f.c0.c1 = c1; f.c1.c1 = c4;
            } else {
                fp12_mul_by_014(f, f, c0, c1, c4);
            }
        }
        if ((BLS_X_ABS >> i) & 1) {
#pragma unroll 1
            for (int j = 0; j < m; j++) {
                miller_add_step_jac(T[j], c0, c1, c4, Q[idx[j]], P[idx[j]]);
                fp12_mul_by_014(f, f, c0, c1, c4);
            }
        }
    }
    fp12_conj(f, f);
}

// f^|x| for f in the cyclotomic subgroup
LHB_HD LHB_NOINLINE void fp12_cyc_pow_x_abs(Fp12& r, const Fp12& a) {
    Fp12 acc = a;
    for (int i = 62; i >= 0; i--) {
        fp12_cyclotomic_sqr(acc, acc);
        if ((BLS_X_ABS >> i) & 1) fp12_mul(acc, acc, a);
    }
    r = acc;
}
// f^x (x negative): conjugate of f^|x| (inverse = conjugate in the cyclotomic subgroup)
LHB_HD LHB_INLINE void fp12_cyc_pow_x(Fp12& r, const Fp12& a) {
    fp12_cyc_pow_x_abs(r, a);
    fp12_conj(r, r);
}

// Final exponentiation.  Returns f^(3 (p^12-1)/r): the hard part uses
//   3 (p^4 - p^2 + 1)/r = (x-1)^2 (x+p) (x^2+p^2-1) + 3      (checked in scripts/gen_bls_consts.py's test)
// The factor 3 is coprime to r, so "== 1" is unchanged; tests compare against oracle_value^3.
LHB_HD LHB_NOINLINE void final_exp(Fp12& r, const Fp12& f_in) {
    Fp12 f, t0, t1, t2;
    // easy part: f^((p^6-1)(p^2+1))
    fp12_conj(t0, f_in);
    fp12_inv(t1, f_in);
    fp12_mul(f, t0, t1);
    fp12_frob2(t0, f);
    fp12_mul(f, t0, f);
    // hard part
    fp12_cyc_pow_x(t0, f);          // f^x
    fp12_conj(t1, f);
    fp12_mul(t0, t0, t1);           // a = f^(x-1)
    fp12_cyc_pow_x(t1, t0);
    fp12_conj(t2, t0);
    fp12_mul(t0, t1, t2);           // a = f^((x-1)^2)
    fp12_cyc_pow_x(t1, t0);         // a^x
    fp12_frob(t2, t0);              // a^p
    fp12_mul(t0, t1, t2);           // b = a^(x+p)
    fp12_cyc_pow_x(t1, t0);
    fp12_cyc_pow_x(t1, t1);         // b^(x^2)
    fp12_frob2(t2, t0);             // b^(p^2)
    fp12_mul(t1, t1, t2);
    fp12_conj(t2, t0);              // b^-1
    fp12_mul(t1, t1, t2);           // c = b^(x^2 + p^2 - 1)
    fp12_cyclotomic_sqr(t2, f);
    fp12_mul(t2, t2, f);            // f^3
    fp12_mul(r, t1, t2);
}

}  // namespace bls
}  // namespace lhb200
