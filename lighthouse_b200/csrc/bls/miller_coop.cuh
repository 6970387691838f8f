// miller_coop.cuh — cooperative multi-pairing Miller loop with ALL state in shared memory (no local-memory stack).
//
// The pairing work behind blst's verify_multiple_aggregate_signatures (crypto/bls/src/impls/blst.rs:114-118):
// prod_i f_{|x|,Q_i}(P_i), conjugated (x < 0).  Round 1 ran one thread per group of sets with a 4.5 KB/thread stack
// that spilled 74 GB per 100 k-set launch to DRAM.  Here
//   * a GROUP of six lanes shares one Fp12 accumulator f = sum a_k w^k (Fp12 = Fp2[w]/(w^6 - xi)); lane t owns the Fp2
//     coefficient a_t.  f^2 and the sparse products f * l are SCHOOLBOOK in the w-basis: lane t computes its own
//     output coefficient as one fused sum of Fp2 products (sop.cuh: lazy reduction, no Karatsuba glue, uniform code
//     across the six lanes);
//   * every lane owns one SignatureSet per round: its point T (homogeneous projective, Costello-Lange-Naehrig doubling:
//     3 M + 6 S in Fp2), the line through it, and the temporaries live in the lane's private shared-memory column;
//   * a block of NT = 96 lanes (16 groups) runs `rounds` sets per lane; between rounds T is parked in global memory
//     (L2), so the shared-memory footprint is 1152 B per lane whatever the batch size, and one squaring of f is shared
//     by 6 * rounds sets.
// Line convention as in pairing.cuh:  l = c0 + c1 w^2 + c4 w^3 (scaled by Fp2 factors the final exponentiation kills):
//   doubling at T = (X, Y, Z):  c0 = (Y^2 - 3 b' Z^2) pz,  c1 = -3 X^2 px,  c4 = 2 Y Z py        (b' = 4 xi)
//   addition of Q to T:         c0 = (u X2 - v Y2) pz,  c1 = -u Z2 px,  c4 = v Z2 py,  u = Y2 Z1 - Y1 Z2, v = X2 Z1 - X1 Z2
// with the G1 argument given projectively as (px, py, pz) = (x_P pz, y_P pz, pz)  (pairing.cuh G1Proj3).
//
// The program is written once as a sequence of PHASES separated by block barriers (MC_PHASE); the host build
// (tests/hostsim) runs the same phases lane by lane, so the whole cooperative algorithm is checked limb-exactly on the
// CPU against the oracle before it ever runs on a GPU.
#pragma once
#include "pairing.cuh"
#include "sop.cuh"

namespace lhb200 {
namespace bls {
namespace mc {

constexpr int S_TX = 0, S_TY = 2, S_TZ = 4;      // T
constexpr int S_F0 = 6, S_F1 = 7, S_NF1 = 8;     // own coefficient of f: (re, im, p - im)
constexpr int S_L0 = 9, S_L1 = 11, S_L4 = 13;    // line coefficients
constexpr int S_T0 = 15, S_T1 = 17, S_T2 = 19, S_T3 = 21;
constexpr int S_SCR = 23;
constexpr int NSLOT = 24;
constexpr int TWORDS = 6 * NL;                    // one projective G2 point

template <int NT>
constexpr size_t smem_bytes() { return (size_t)NSLOT * NL * NT * 4 + NT; }

// squaring schedule of lane t: four terms  a_i * a_j  (x = a_i times 2 and/or xi, y = a_j);  byte = i | j << 3 |
// dbl << 6 | xi << 7, 0xff = no term.  r_t = sum_{i+j = t} a_i a_j + xi sum_{i+j = t+6} a_i a_j.
#define LHB_SQT(i, j, d, x) ((uint8_t)((i) | ((j) << 3) | ((d) << 6) | ((x) << 7)))
LHB_CONST uint8_t SQR_TERMS[6][4] = {
    {LHB_SQT(0, 0, 0, 0), LHB_SQT(1, 5, 1, 1), LHB_SQT(2, 4, 1, 1), LHB_SQT(3, 3, 0, 1)},
    {LHB_SQT(0, 1, 1, 0), LHB_SQT(2, 5, 1, 1), LHB_SQT(3, 4, 1, 1), 0xff},
    {LHB_SQT(0, 2, 1, 0), LHB_SQT(1, 1, 0, 0), LHB_SQT(3, 5, 1, 1), LHB_SQT(4, 4, 0, 1)},
    {LHB_SQT(0, 3, 1, 0), LHB_SQT(1, 2, 1, 0), LHB_SQT(4, 5, 1, 1), 0xff},
    {LHB_SQT(0, 4, 1, 0), LHB_SQT(1, 3, 1, 0), LHB_SQT(2, 2, 0, 0), LHB_SQT(5, 5, 0, 1)},
    {LHB_SQT(0, 5, 1, 0), LHB_SQT(1, 4, 1, 0), LHB_SQT(2, 3, 1, 0), 0xff},
};

template <int NT>
struct Lane {
    Col<NT> c;
    int tid, t;          // t = lane within its group of six
    uint8_t* act;        // active flags of the current round, one per lane of the block
    bool active;         // this lane's set of the current round takes part
    bool extra;          // ... and it is the appended pair (-g1, sum r sig)
    uint32_t set;
    Fp ra, rb;           // result coefficient in flight between a compute phase and its store phase
};

struct Args {
    const G1Proj3* P;
    const G2Jac* H;
    const uint8_t* status;
    uint32_t n;                // sets in P / H / status
    const G2Jac* extra_q;      // nullable: the pair (*extra_p, *extra_q) = (-g1, sum r sig) is appended as set n
    const G1Proj3* extra_p;
    uint32_t lo, hi;           // this block's sets [lo, hi)
    uint32_t* scratch;         // this block's parking area: rounds * 2 * TWORDS * NT words (T and Q per set)
    Fp12* out;                 // this block's NT / 6 group products
};

template <int NT>
LHB_HD LHB_INLINE const G1Proj3* lane_p(const Lane<NT>& L, const Args& a) { return L.extra ? a.extra_p : a.P + L.set; }
template <int NT>
LHB_HD LHB_INLINE void ld_words(const Col<NT>& c, int slot, const uint32_t* g, int nwords) {  // g[w * NT] -> column
    uint32_t* q = c.base() + slot * NL * NT;
#pragma unroll 8
    for (int w = 0; w < nwords; w++) q[w * NT] = g[w * NT];
}
template <int NT>
LHB_HD LHB_INLINE void st_words(uint32_t* g, const Col<NT>& c, int slot, int nwords) {
    const uint32_t* q = c.base() + slot * NL * NT;
#pragma unroll 8
    for (int w = 0; w < nwords; w++) g[w * NT] = q[w * NT];
}
// bind the lane to its set of round r
template <int NT>
LHB_HD LHB_INLINE void lane_select(Lane<NT>& L, const Args& a, uint32_t r) {
    const uint32_t n_total = a.n + (a.extra_q ? 1u : 0u);
    L.set = a.lo + r * NT + L.tid;
    L.extra = false;
    bool act = L.set < a.hi && L.set < n_total;
    if (act) {
        if (L.set >= a.n) { L.extra = true; act = !jac_is_inf(*a.extra_q); }
        else act = a.status[L.set] == 0 && !jac_is_inf(a.H[L.set]);
    }
    L.active = act;
    L.act[L.tid] = act ? 1 : 0;
}
template <int NT>
LHB_HD LHB_INLINE uint32_t* park_t(const Lane<NT>& L, const Args& a, uint32_t r) { return a.scratch + (size_t)(2 * r) * TWORDS * NT + L.tid; }
template <int NT>
LHB_HD LHB_INLINE uint32_t* park_q(const Lane<NT>& L, const Args& a, uint32_t r) { return a.scratch + (size_t)(2 * r + 1) * TWORDS * NT + L.tid; }

// ---- prologue of a round: Q = H(m) Jacobian (X, Y, Z) -> projective (X Z, Y, Z^3) = T; parked copy for the additions
template <int NT>
LHB_HD LHB_INLINE void phase_init_point(Lane<NT>& L, const Args& a, uint32_t* q_park) {
    if (!L.active) return;
    const Col<NT>& c = L.c;
    const G2Jac& h = L.extra ? *a.extra_q : a.H[L.set];
    c.st2(S_T0, h.X); c.st2(S_TY, h.Y); c.st2(S_T1, h.Z);
    c2_mul(c, S_TX, S_T0, S_T1);
    c2_sqr(c, S_TZ, S_T1, S_SCR);
    c2_mul(c, S_TZ, S_TZ, S_T1);
    st_words(q_park, c, S_TX, TWORDS);
}

// ---- doubling step: T <- 2T, line -> L slots.  54 multiply units (3 M + 6 S + 3 scalings), see the header.
template <int NT>
LHB_HD LHB_INLINE void phase_dbl(Lane<NT>& L, const Args& a) {
    if (!L.active) return;
    const Col<NT>& c = L.c;
    const G1Proj3* P = lane_p(L, a);
    c2_mul(c, S_T0, S_TX, S_TY); c2_half(c, S_T0, S_T0);                           // A = X Y / 2
    c2_sqr(c, S_T1, S_TY, S_SCR);                                                  // B = Y^2
    c2_add(c, S_T3, S_TY, S_TZ); c2_sqr(c, S_T3, S_T3, S_SCR);                     // (Y + Z)^2
    c2_sqr(c, S_T2, S_TZ, S_SCR);                                                  // C = Z^2
    c2_sub(c, S_T3, S_T3, S_T1); c2_sub(c, S_T3, S_T3, S_T2);                      // H = 2 Y Z
    c2_sqr(c, S_L1, S_TX, S_SCR); c2_triple(c, S_L1, S_L1); c2_neg(c, S_L1, S_L1); // -3 X^2
    c2_mul_fp(c, S_L1, S_L1, &P->px);
    c2_mul_12xi(c, S_T2, S_T2);                                                    // E = 3 b' C
    c2_sub(c, S_L0, S_T1, S_T2);                                                   // B - E
    c2_mul_fp(c, S_L0, S_L0, &P->pz);
    c2_mul_fp(c, S_L4, S_T3, &P->py);                                              // H py
    c2_triple(c, S_TZ, S_T2);                                                      // F = 3 E   (Z is dead)
    c2_sub(c, S_TX, S_T1, S_TZ); c2_mul(c, S_TX, S_T0, S_TX);                      // X3 = A (B - F)
    c2_add(c, S_TY, S_T1, S_TZ); c2_half(c, S_TY, S_TY); c2_sqr(c, S_TY, S_TY, S_SCR);   // G^2, G = (B + F)/2
    c2_sqr(c, S_T0, S_T2, S_SCR); c2_triple(c, S_T0, S_T0); c2_sub(c, S_TY, S_TY, S_T0); // Y3 = G^2 - 3 E^2
    c2_mul(c, S_TZ, S_T1, S_T3);                                                   // Z3 = B H
}

// ---- addition step: T <- T + Q (projective, add-1998-cmo-2), line through T and Q -> L slots
template <int NT>
LHB_HD LHB_INLINE void phase_add(Lane<NT>& L, const Args& a, const uint32_t* q_park) {
    if (!L.active) return;
    const Col<NT>& c = L.c;
    const G1Proj3* P = lane_p(L, a);
    const uint32_t *qx = q_park, *qy = q_park + 2 * NL * NT, *qz = q_park + 4 * NL * NT;   // X2, Y2, Z2
    c2_mul_g(c, S_T0, qz, S_TY);                                                   // Y1 Z2
    c2_mul_g(c, S_T1, qz, S_TX);                                                   // X1 Z2
    c2_mul_g(c, S_T2, qz, S_TZ);                                                   // Z1 Z2
    c2_mul_g(c, S_T3, qy, S_TZ); c2_sub(c, S_T3, S_T3, S_T0);                      // u = Y2 Z1 - Y1 Z2
    c2_mul_g(c, S_TZ, qx, S_TZ); c2_sub(c, S_TZ, S_TZ, S_T1);                      // v = X2 Z1 - X1 Z2
    c2_sqr(c, S_TX, S_T3, S_SCR);                                                  // uu
    c2_sqr(c, S_TY, S_TZ, S_SCR);                                                  // vv
    c2_mul(c, S_L0, S_TZ, S_TY);                                                   // vvv
    c2_mul(c, S_T1, S_TY, S_T1);                                                   // R = vv X1Z2
    c2_mul(c, S_TX, S_TX, S_T2);                                                   // uu Z1Z2
    c2_sub(c, S_TX, S_TX, S_L0); c2_sub(c, S_TX, S_TX, S_T1); c2_sub(c, S_TX, S_TX, S_T1);   // A
    c2_mul(c, S_L1, S_TZ, S_TX);                                                   // X3 = v A
    c2_sub(c, S_TY, S_T1, S_TX);
    c2_mul(c, S_TY, S_T3, S_TY);                                                   // u (R - A)
    c2_mul(c, S_T0, S_L0, S_T0);                                                   // vvv Y1Z2
    c2_sub(c, S_TY, S_TY, S_T0);                                                   // Y3
    c2_mul(c, S_T2, S_L0, S_T2);                                                   // Z3 = vvv Z1Z2 (kept in T2)
    c2_copy(c, S_TX, S_L1);
    // line (u in T3, v in TZ)
    c2_mul_g(c, S_L0, qx, S_T3);                                                   // u X2
    c2_mul_g(c, S_T0, qy, S_TZ);                                                   // v Y2
    c2_sub(c, S_L0, S_L0, S_T0);
    c2_mul_fp(c, S_L0, S_L0, &P->pz);
    c2_mul_g(c, S_L1, qz, S_T3); c2_neg(c, S_L1, S_L1);                            // -u Z2
    c2_mul_fp(c, S_L1, S_L1, &P->px);
    c2_mul_g(c, S_L4, qz, S_TZ);                                                   // v Z2
    c2_mul_fp(c, S_L4, S_L4, &P->py);
    c2_copy(c, S_TZ, S_T2);
}

// ---- f <- f^2: lane t's coefficient.  K = 8 (four Fp2 terms), X = Y = 1
template <int NT>
LHB_HD LHB_INLINE void phase_sqr_compute(Lane<NT>& L) {
    const Col<NT> g = L.c.lane(-L.t);   // column of the group's lane 0
    SopX<8> x;
    SopY<8> ya, yb;
    ya.stride = NT; yb.stride = NT;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t e = SQR_TERMS[L.t][k];
        const bool valid = e != 0xff;
        const int i = valid ? (e & 7) : 0, j = valid ? ((e >> 3) & 7) : 0;
        Fp2 u;
        g.lane(i).ld2(u, S_F0);
        if (e & 0x40) { fp_add_inl(u.c0, u.c0, u.c0); fp_add_inl(u.c1, u.c1, u.c1); }
        if (e & 0x80) fp2_mul_xi_inl(u, u);
        if (!valid) { fp_set_zero(u.c0); fp_set_zero(u.c1); }
        x.x[2 * k] = u.c0; x.x[2 * k + 1] = u.c1;
        const Col<NT> cj = g.lane(j);
        ya.base[2 * k] = cj.at(S_F0); ya.base[2 * k + 1] = cj.at(S_NF1);   // u0 a0 + u1 (-a1)
        yb.base[2 * k] = cj.at(S_F1); yb.base[2 * k + 1] = cj.at(S_F0);    // u0 a1 + u1 a0
    }
    fp_sop2<8>(L.ra, L.rb, x, ya, x, yb);
}
// ---- f <- f * l(owner): lane t's coefficient  a_t c0 + xi^[t<2] a_{t-2} c1 + xi^[t<3] a_{t-3} c4.  K = 6
template <int NT>
LHB_HD LHB_INLINE void phase_sparse_compute(Lane<NT>& L, int owner) {
    const Col<NT> g = L.c.lane(-L.t);
    if (!L.act[L.tid - L.t + owner]) return;   // uniform within the group
    const Col<NT> oc = g.lane(owner);
    Fp2 l0, l1, l4;
    oc.ld2(l0, S_L0); oc.ld2(l1, S_L1); oc.ld2(l4, S_L4);
    if (L.t < 2) fp2_mul_xi_inl(l1, l1);
    if (L.t < 3) fp2_mul_xi_inl(l4, l4);
    SopX<6> x;
    x.x[0] = l0.c0; x.x[1] = l0.c1; x.x[2] = l1.c0; x.x[3] = l1.c1; x.x[4] = l4.c0; x.x[5] = l4.c1;
    const int i1 = L.t >= 2 ? L.t - 2 : L.t + 4, i2 = L.t >= 3 ? L.t - 3 : L.t + 3;
    const Col<NT> c0 = L.c, c1 = g.lane(i1), c2 = g.lane(i2);
    SopY<6> ya, yb;
    ya.stride = NT; yb.stride = NT;
    ya.base[0] = c0.at(S_F0); ya.base[1] = c0.at(S_NF1); ya.base[2] = c1.at(S_F0); ya.base[3] = c1.at(S_NF1);
    ya.base[4] = c2.at(S_F0); ya.base[5] = c2.at(S_NF1);
    yb.base[0] = c0.at(S_F1); yb.base[1] = c0.at(S_F0); yb.base[2] = c1.at(S_F1); yb.base[3] = c1.at(S_F0);
    yb.base[4] = c2.at(S_F1); yb.base[5] = c2.at(S_F0);
    fp_sop2<6>(L.ra, L.rb, x, ya, x, yb);
}
template <int NT>
LHB_HD LHB_INLINE void phase_store_f(Lane<NT>& L) {
    Fp n;
    fp_neg(n, L.rb);
    L.c.st(S_F0, L.ra); L.c.st(S_F1, L.rb); L.c.st(S_NF1, n);
}
template <int NT>
LHB_HD LHB_INLINE void phase_store_f_if(Lane<NT>& L, int owner) {
    if (L.act[L.tid - L.t + owner]) phase_store_f(L);
}

// The whole program of one block: PHASES separated by block barriers.  MC_PHASE(stmts) runs stmts on the lane `L` and
// synchronises (device), or runs them on every lane of `ex.lanes` in turn (host simulation).
#ifdef LHB_HOSTSIM
#define MC_PHASE(...) do { for (auto& L : ex.lanes) { __VA_ARGS__; } } while (0)
#else
#define MC_PHASE(...) do { Lane<NT>& L = ex.L; { __VA_ARGS__; } __syncthreads(); } while (0)
#endif
template <int NT, class Exec>
LHB_HD LHB_INLINE void miller_program(Exec& ex, const Args& a) {
    const uint32_t cnt = a.hi - a.lo;
    const uint32_t rounds = (cnt + NT - 1) / NT;
    // f = 1; points
    MC_PHASE(Fp z; fp_set_zero(z); Fp one = FP_ONE;
             L.c.st(S_F0, L.t == 0 ? one : z); L.c.st(S_F1, z); L.c.st(S_NF1, z));
    for (uint32_t r = 0; r < rounds; r++)
        MC_PHASE(lane_select(L, a, r);
                 phase_init_point(L, a, park_q(L, a, r));
                 if (rounds > 1 && L.active) st_words(park_t(L, a, r), L.c, S_TX, TWORDS));
#pragma unroll 1
    for (int it = 62; it >= 0; it--) {
        MC_PHASE(phase_sqr_compute(L));
        MC_PHASE(phase_store_f(L));
        const int steps = ((BLS_X_ABS >> it) & 1) ? 2 : 1;
#pragma unroll 1
        for (uint32_t r = 0; r < rounds; r++) {
#pragma unroll 1
            for (int step = 0; step < steps; step++) {
                if (step == 0) {
                    MC_PHASE(if (rounds > 1) { lane_select(L, a, r); if (L.active) ld_words(L.c, S_TX, park_t(L, a, r), TWORDS); }
                             phase_dbl(L, a);
                             if (rounds > 1 && steps == 1 && L.active) st_words(park_t(L, a, r), L.c, S_TX, TWORDS));
                } else {
                    MC_PHASE(phase_add(L, a, park_q(L, a, r));
                             if (rounds > 1 && L.active) st_words(park_t(L, a, r), L.c, S_TX, TWORDS));
                }
#pragma unroll 1
                for (int owner = 0; owner < 6; owner++) {
                    MC_PHASE(phase_sparse_compute(L, owner));
                    MC_PHASE(phase_store_f_if(L, owner));
                }
            }
        }
    }
    // conjugate (x < 0): negate the odd powers of w; write the group's product in tower layout
    MC_PHASE(Fp2 v; L.c.ld2(v, S_F0);
             if (L.t & 1) { fp_neg(v.c0, v.c0); fp_neg(v.c1, v.c1); }
             Fp12& o = a.out[L.tid / 6];
             Fp2& dst = L.t == 0 ? o.c0.c0 : L.t == 1 ? o.c1.c0 : L.t == 2 ? o.c0.c1 : L.t == 3 ? o.c1.c1 : L.t == 4 ? o.c0.c2 : o.c1.c2;
             dst = v);
}

#ifndef LHB_HOSTSIM
template <int NT>
struct ExecDev {
    Lane<NT>& L;
};

// One block = NT lanes = NT / 6 groups; block b owns sets [b * sets_per_block, (b + 1) * sets_per_block) of the n (+1)
// pairs and writes NT / 6 group products to out_f[b * NT / 6 ...].  scratch: per block sets_per_block rounded up to
// whole rounds, 2 * TWORDS words per set.
template <int NT>
__global__ void __launch_bounds__(NT) k_miller_coop(const G1Proj3* __restrict__ P, const G2Jac* __restrict__ H,
                                                     const uint8_t* __restrict__ status, uint32_t n,
                                                     const G2Jac* __restrict__ extra_q,
                                                     const G1Proj3* __restrict__ extra_p, uint32_t sets_per_block,
                                                     uint32_t* __restrict__ scratch, Fp12* __restrict__ out_f) {
    uint32_t* const mc_smem = lhb_dyn_smem;
    Lane<NT> L;
    L.tid = threadIdx.x;
    L.t = threadIdx.x % 6;
    L.c = Col<NT>::make(mc_smem, threadIdx.x);
    L.act = reinterpret_cast<uint8_t*>(mc_smem + NSLOT * NL * NT);
    L.active = false; L.extra = false; L.set = 0;
    const uint32_t n_total = n + (extra_q ? 1u : 0u);
    Args a;
    a.P = P; a.H = H; a.status = status; a.n = n; a.extra_q = extra_q; a.extra_p = extra_p;
    a.lo = blockIdx.x * sets_per_block;
    a.hi = min(n_total, a.lo + sets_per_block);
    if (a.lo > a.hi) a.lo = a.hi;
    const uint32_t rounds_cap = (sets_per_block + NT - 1) / NT;
    a.scratch = scratch + (size_t)blockIdx.x * rounds_cap * 2 * TWORDS * NT;
    a.out = out_f + (size_t)blockIdx.x * (NT / 6);
    ExecDev<NT> ex{L};
    miller_program<NT>(ex, a);
}
#endif

}  // namespace mc
}  // namespace bls
}  // namespace lhb200
