// kernels.cuh — CUDA kernels of the batch BLS verification path (sm_100a).  One thread owns one SignatureSet
// through each stage (the work is long, branch-light and identical across sets; parallelism comes from the
// batch).  Stage outputs go through HBM in AoS form: they are tiny next to the arithmetic
// (1.2 KB per set against ~20k Montgomery multiplications).
//
// Stage map:
//   k_sig_prepare ...... K10 g2_decompress + K4 subgroup check + K7 r*sig        (blst.rs:73-83, :114)
//   k_pk_aggregate ..... K5 segmented G1 sum over CSR offsets + K7 r*apk          (blst.rs:86-106, :114)
//   k_hash_to_g2 ....... K6 hash_to_curve                                         (blst.rs:114, DST :15)
//   k_miller_multi ..... K8 Miller loops, k sets per thread sharing the Fp12 squarings
//   k_fp12_reduce / k_g2_reduce ... product / sum trees
//   k_final ............ K8 Miller loop for (-g1, sum r*sig) + K9 final exponentiation and == 1
#pragma once
#include "pairing.cuh"
#include "miller_coop.cuh"
#include "h2c.cuh"

namespace lhb200 {
namespace bls {

enum SetStatus : uint8_t {
    SET_OK = 0,
    SET_EMPTY_SIG = 1,      // all-zero "empty" signature (generic_signature.rs:26) -> batch false (blst.rs:79-82)
    SET_SIG_DECODE = 2,     // malformed compressed G2
    SET_SIG_SUBGROUP = 3,   // blst.rs:75-77
    SET_NO_KEYS = 4,        // blst.rs:86-89
    SET_APK_INFINITY = 5,   // aggregate key at infinity (Appendix C item 5)
    SET_PK_DECODE = 6,      // malformed uncompressed G1 key
};

constexpr int BLS_BLOCK = 64;
#ifndef LHB_MILLER_BLOCK
#define LHB_MILLER_BLOCK 64
#endif
constexpr int MILLER_BLOCK = LHB_MILLER_BLOCK;  // k_miller_multi's block size (its 4.5 KB/thread stack vs the 126 MB L2)

__device__ __forceinline__ void load_bytes16(uint8_t* dst, const uint8_t* src, int nbytes) {
    // src is 16-byte aligned; nbytes multiple of 16
    for (int i = 0; i < nbytes; i += 16) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(src + i));
        *reinterpret_cast<uint4*>(dst + i) = v;
    }
}

__global__ void __launch_bounds__(BLS_BLOCK) k_sig_prepare(const uint8_t* __restrict__ sigs,
                                                            const uint64_t* __restrict__ rands, uint32_t n,
                                                            G2Jac* __restrict__ sig_r, uint8_t* __restrict__ status,
                                                            uint32_t* __restrict__ fail) {
    // grid-stride: the host caps resident CTAs per SM so the per-thread stacks stay cache-resident
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        __align__(16) uint8_t b[96];
        load_bytes16(b, sigs + 96ull * i, 96);
        uint32_t nz = 0;
        for (int k = 0; k < 96; k++) nz |= b[k];
        G2Jac out;
        jac_set_inf(out);
        uint8_t st = SET_OK;
        if (nz == 0) {
            st = SET_EMPTY_SIG;
        } else {
            G2Affine a;
            const int32_t rc = g2_decompress(a, b);
            if (rc == DEC_BAD) st = SET_SIG_DECODE;
            else if (rc == DEC_OK) {
                // [r]sig and [|x|]sig in one pass; in G2  <=>  psi(sig) == -[|x|]sig   (blst.rs:75)
                G2Jac xs, ps, aj;
                g2_mul_r_and_x(out, xs, a, rands[i]);
                jac_neg(xs, xs);
                jac_from_affine(aj, a);
                g2_psi(ps, aj);
                if (!jac_eq(ps, xs)) { st = SET_SIG_SUBGROUP; jac_set_inf(out); }
            }
            // DEC_INFINITY: the infinity signature passes the subgroup check and contributes nothing to the sum
        }
        sig_r[i] = out;
        if (st != SET_OK) { status[i] = st; atomicOr(fail, 1u); }
    }
}

__global__ void __launch_bounds__(BLS_BLOCK) k_pk_aggregate(const uint8_t* __restrict__ pks,
                                                             const uint32_t* __restrict__ offsets,
                                                             const uint64_t* __restrict__ rands, uint32_t n,
                                                             G1Proj3* __restrict__ out_p, uint8_t* __restrict__ status,
                                                             uint32_t* __restrict__ fail) {
    // grid-stride: the host caps resident CTAs per SM so the per-thread stacks stay cache-resident
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t lo = offsets[i], hi = offsets[i + 1];
        uint8_t st = SET_OK;
        G1Jac acc;
        jac_set_inf(acc);
        if (hi <= lo) st = SET_NO_KEYS;
        for (uint32_t j = lo; j < hi && st == SET_OK; j++) {
            __align__(16) uint8_t b[96];
            load_bytes16(b, pks + 96ull * j, 96);
            G1Affine a;
            if (g1_from_uncompressed(a, b) == DEC_BAD) { st = SET_PK_DECODE; break; }
            jac_add_affine(acc, acc, a);
        }
        if (st == SET_OK && jac_is_inf(acc)) st = SET_APK_INFINITY;
        G1Proj3 P;
        if (st == SET_OK) {
            G1Jac ra;
            jac_mul_u64(ra, acc, rands[i]);
            g1proj3_from_jac(P, ra);
        } else {
            P.px = FP_ONE; P.py = FP_ONE; P.pz = FP_ONE;
        }
        out_p[i] = P;
        if (st != SET_OK) { status[i] = st; atomicOr(fail, 1u); }
    }
}

// Device-resident pubkey table (mirror of ValidatorPubkeyCache, beacon_chain/src/validator_pubkey_cache.rs:20-25):
// entries are affine G1 in Montgomery form, converted once at import, so per-set aggregation needs no decoding.
struct G1Mont {
    Fp x, y;
};
__global__ void __launch_bounds__(BLS_BLOCK) k_table_import(const uint8_t* __restrict__ pks96, uint32_t n,
                                                             G1Mont* __restrict__ out, uint32_t* __restrict__ n_bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    __align__(16) uint8_t b[96];
    load_bytes16(b, pks96 + 96ull * i, 96);
    G1Affine a;
    const int32_t rc = g1_from_uncompressed(a, b);
    G1Mont m;
    m.x = a.x; m.y = a.y;
    if (rc != DEC_OK || !g1_on_curve(a)) {  // infinity is rejected at import like generic_public_key.rs:87-88
        atomicAdd(n_bad, 1u);
        fp_set_zero(m.x); fp_set_zero(m.y);
    }
    out[i] = m;
}

__global__ void __launch_bounds__(BLS_BLOCK) k_pk_aggregate_indexed(const G1Mont* __restrict__ table, uint32_t table_len,
                                                                     const uint32_t* __restrict__ indices,
                                                                     const uint32_t* __restrict__ offsets,
                                                                     const uint64_t* __restrict__ rands, uint32_t n,
                                                                     G1Proj3* __restrict__ out_p,
                                                                     uint8_t* __restrict__ status,
                                                                     uint32_t* __restrict__ fail) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t lo = offsets[i], hi = offsets[i + 1];
        uint8_t st = SET_OK;
        G1Jac acc;
        jac_set_inf(acc);
        if (hi <= lo) st = SET_NO_KEYS;
        for (uint32_t j = lo; j < hi && st == SET_OK; j++) {
            const uint32_t idx = __ldg(indices + j);
            if (idx >= table_len) { st = SET_PK_DECODE; break; }
            G1Affine a;
            const G1Mont& m = table[idx];
            a.x = m.x; a.y = m.y; a.inf = 0;
            jac_add_affine(acc, acc, a);
        }
        if (st == SET_OK && jac_is_inf(acc)) st = SET_APK_INFINITY;
        G1Proj3 P;
        if (st == SET_OK) {
            G1Jac ra;
            jac_mul_u64(ra, acc, rands[i]);
            g1proj3_from_jac(P, ra);
        } else {
            P.px = FP_ONE; P.py = FP_ONE; P.pz = FP_ONE;
        }
        out_p[i] = P;
        if (st != SET_OK) { status[i] = st; atomicOr(fail, 1u); }
    }
}

__global__ void __launch_bounds__(BLS_BLOCK) k_hash_to_g2(const uint8_t* __restrict__ msgs, uint32_t n,
                                                           G2Jac* __restrict__ out_h) {
    // grid-stride: the host caps resident CTAs per SM so the per-thread stacks stay cache-resident
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        __align__(16) uint8_t m[32];
        load_bytes16(m, msgs + 32ull * i, 32);
        G2Jac j;
        hash_to_g2_jac(j, m);
        out_h[i] = j;   // stays Jacobian: the Miller loop's addition steps take a projective Q (no inversion here)
    }
}

// Group g = sets [g*k, (g+1)*k): one thread runs their Miller loops with shared squarings; out_f[g] = the product.
// The host picks k = ceil(n / resident threads) so that every resident thread gets one group (no partial last wave).
__global__ void __launch_bounds__(MILLER_BLOCK) k_miller_multi(const G1Proj3* __restrict__ P, const G2Jac* __restrict__ H,
                                                             const uint8_t* __restrict__ status, uint32_t n, uint32_t k,
                                                             uint32_t n_groups, Fp12* __restrict__ out_f) {
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += gridDim.x * blockDim.x) {
        uint32_t idx[MILLER_KMAX];
        int m = 0;
        for (uint32_t j = 0; j < k; j++) {
            const uint32_t i = g * k + j;
            if (i < n && status[i] == SET_OK && !jac_is_inf(H[i])) idx[m++] = i;
        }
        Fp12 f;
        if (m == 0) fp12_set_one(f);
        else miller_loop_multi(f, P, H, idx, m);
        out_f[g] = f;
    }
}

// out[t] = prod in[t*chunk .. min(n,(t+1)*chunk))
__global__ void __launch_bounds__(BLS_BLOCK) k_fp12_reduce(const Fp12* __restrict__ in, uint32_t n, uint32_t chunk,
                                                            Fp12* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t lo = (uint64_t)t * chunk;
    if (lo >= n) return;
    const uint32_t hi = (uint32_t)min((uint64_t)n, lo + chunk);
    Fp12 acc = in[lo];
    for (uint32_t j = (uint32_t)lo + 1; j < hi; j++) {
        Fp12 x = in[j];
        fp12_mul(acc, acc, x);
    }
    out[t] = acc;
}
__global__ void __launch_bounds__(BLS_BLOCK) k_g2_reduce(const G2Jac* __restrict__ in, uint32_t n, uint32_t chunk,
                                                          G2Jac* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t lo = (uint64_t)t * chunk;
    if (lo >= n) return;
    const uint32_t hi = (uint32_t)min((uint64_t)n, lo + chunk);
    G2Jac acc = in[lo];
    for (uint32_t j = (uint32_t)lo + 1; j < hi; j++) {
        G2Jac x = in[j];
        jac_add(acc, acc, x);
    }
    out[t] = acc;
}

// The G1 argument of the aggregated-signature pair e(-g1, sum r sig), in the Miller kernels' projective form.
__global__ void k_init_neg_g1(G1Proj3* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G1Proj3 p;
    p.px = G1_GEN_X;
    fp_neg(p.py, G1_GEN_Y);
    p.pz = FP_ONE;
    *out = p;
}

// f_last = Miller(-g1, S) for the aggregated signature term; runs concurrently with k_miller_multi.
__global__ void k_last_miller(const G2Jac* __restrict__ sig_sum, Fp12* __restrict__ out_f) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G2Jac s = *sig_sum;
    Fp12 f;
    if (jac_is_inf(s)) {
        fp12_set_one(f);
    } else {
        G2Affine q;
        jac_to_affine(q, s);
        G1Proj3 p;
        p.px = G1_GEN_X;
        fp_neg(p.py, G1_GEN_Y);
        p.pz = FP_ONE;
        miller_loop(f, p, q);
    }
    *out_f = f;
}

// verdict = !fail && final_exp(prod * f_last) == 1 ; also exposes the GT value for tests
__global__ void k_final(const Fp12* __restrict__ prod, const Fp12* __restrict__ f_last, const uint32_t* __restrict__ fail,
                        uint8_t* __restrict__ ok, Fp12* __restrict__ gt_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (*fail) { *ok = 0; return; }
    Fp12 a = *prod, b = *f_last;
    fp12_mul(a, a, b);
    final_exp(a, a);
    if (gt_out) *gt_out = a;
    *ok = fp12_is_one(a) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------
// Key-side kernels (SecretKey surface of crypto/bls: sk -> pk, sign) — also the synthetic-workload generators.
// sk: 32-byte big-endian scalars (already reduced mod r).
__global__ void __launch_bounds__(BLS_BLOCK) k_sk_to_pk(const uint8_t* __restrict__ sks, uint32_t n,
                                                         uint8_t* __restrict__ pk48, uint8_t* __restrict__ pk96) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k[8];
    const uint8_t* s = sks + 32ull * i;
    for (int w = 0; w < 8; w++) {
        const uint8_t* q = s + 4 * (7 - w);
        k[w] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    G1Affine g;
    g.x = G1_GEN_X; g.y = G1_GEN_Y; g.inf = 0;
    G1Jac j;
    jac_mul_affine(j, g, k, 255);
    G1Affine a;
    jac_to_affine(a, j);
    uint8_t b[96];
    if (pk48) { g1_compress(b, a); for (int t = 0; t < 48; t++) pk48[48ull * i + t] = b[t]; }
    if (pk96) { g1_to_uncompressed(b, a); for (int t = 0; t < 96; t++) pk96[96ull * i + t] = b[t]; }
}

__global__ void __launch_bounds__(BLS_BLOCK) k_sign(const uint8_t* __restrict__ sks, const uint8_t* __restrict__ msgs,
                                                     uint32_t n, uint8_t* __restrict__ sig96) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k[8];
    const uint8_t* s = sks + 32ull * i;
    for (int w = 0; w < 8; w++) {
        const uint8_t* q = s + 4 * (7 - w);
        k[w] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    __align__(16) uint8_t m[32];
    for (int t = 0; t < 32; t++) m[t] = msgs[32ull * i + t];
    G2Jac h, r;
    hash_to_g2_jac(h, m);
    jac_mul(r, h, k, 255);
    G2Affine a;
    jac_to_affine(a, r);
    uint8_t b[96];
    g2_compress(b, a);
    for (int t = 0; t < 96; t++) sig96[96ull * i + t] = b[t];
}


// ---------------------------------------------------------------------------------------------------------
// Aggregation surface of a crypto/bls backend (TAggregateSignature::add_assign / add_assign_aggregate
// blst.rs:230-237, TAggregatePublicKey::aggregate blst.rs:178-184, deserialize_uncompressed blst.rs:142-150).
// sum tree over G1 points (k_g2_reduce's twin)
__global__ void __launch_bounds__(BLS_BLOCK) k_g1_reduce(const G1Jac* __restrict__ in, uint32_t n, uint32_t chunk,
                                                          G1Jac* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t lo = (uint64_t)t * chunk;
    if (lo >= n) return;
    const uint32_t hi = (uint32_t)min((uint64_t)n, lo + chunk);
    G1Jac acc = in[lo];
    for (uint32_t j = (uint32_t)lo + 1; j < hi; j++) {
        G1Jac x = in[j];
        jac_add(acc, acc, x);
    }
    out[t] = acc;
}
// compressed signatures -> Jacobian points (infinity = identity); any malformed encoding raises *n_bad
__global__ void __launch_bounds__(BLS_BLOCK) k_g2_load_points(const uint8_t* __restrict__ sig96, uint32_t n,
                                                               G2Jac* __restrict__ out, uint32_t* __restrict__ n_bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t b[96];
    for (int t = 0; t < 96; t++) b[t] = sig96[96ull * i + t];
    G2Affine a;
    G2Jac j;
    jac_set_inf(j);
    const int32_t rc = g2_decompress(a, b);
    if (rc == DEC_BAD) atomicAdd(n_bad, 1u);
    else if (rc == DEC_OK) jac_from_affine(j, a);
    out[i] = j;
}
__global__ void k_g2_store_point(const G2Jac* __restrict__ in, uint8_t* __restrict__ out96) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G2Jac j = *in;
    G2Affine a;
    jac_to_affine(a, j);
    uint8_t b[96];
    g2_compress(b, a);
    for (int t = 0; t < 96; t++) out96[t] = b[t];
}
// uncompressed keys -> Jacobian points; status as lhb200_g1_deserialize_uncompressed (0 ok, 1 infinity, 2 bad)
__global__ void __launch_bounds__(BLS_BLOCK) k_g1_load_points(const uint8_t* __restrict__ pk96, uint32_t n,
                                                               G1Jac* __restrict__ out, uint8_t* __restrict__ pk48,
                                                               uint8_t* __restrict__ st, uint32_t* __restrict__ n_bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    __align__(16) uint8_t b[96];
    for (int t = 0; t < 96; t++) b[t] = pk96[96ull * i + t];
    G1Affine a;
    int32_t rc = g1_from_uncompressed(a, b);
    if (rc == DEC_OK && !g1_on_curve(a)) rc = DEC_BAD;
    if (rc == DEC_BAD) { atomicAdd(n_bad, 1u); a.inf = 1; }
    if (out) { G1Jac j; jac_from_affine(j, a); out[i] = j; }
    if (st) st[i] = (uint8_t)rc;
    if (pk48) {
        uint8_t c[48];
        if (rc == DEC_OK) g1_compress(c, a);
        else for (int t = 0; t < 48; t++) c[t] = 0;
        for (int t = 0; t < 48; t++) pk48[48ull * i + t] = c[t];
    }
}
__global__ void k_g1_store_point(const G1Jac* __restrict__ in, uint8_t* __restrict__ out48, uint8_t* __restrict__ out96) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G1Jac j = *in;
    G1Affine a;
    jac_to_affine(a, j);
    uint8_t b[96];
    if (out48) { g1_compress(b, a); for (int t = 0; t < 48; t++) out48[t] = b[t]; }
    if (out96) { g1_to_uncompressed(b, a); for (int t = 0; t < 96; t++) out96[t] = b[t]; }
}

// PublicKey::deserialize + key_validate (blst.rs:130-140): decompress, reject infinity, subgroup check.
// status: 0 ok, 1 infinity, 2 bad encoding / not on curve, 3 not in subgroup.
__global__ void __launch_bounds__(BLS_BLOCK) k_g1_decompress_validate(const uint8_t* __restrict__ pk48, uint32_t n,
                                                                       uint8_t* __restrict__ pk96,
                                                                       uint8_t* __restrict__ st) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t b[96];
    for (int t = 0; t < 48; t++) b[t] = pk48[48ull * i + t];
    G1Affine a;
    const int32_t rc = g1_decompress(a, b);
    uint8_t s = 0;
    if (rc == DEC_BAD) s = 2;
    else if (rc == DEC_INFINITY) s = 1;
    else {
        // [r]P == inf  (r = group order, 255 bits)
        const uint32_t R_ORDER[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                     0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        G1Jac j;
        jac_mul_affine(j, a, R_ORDER, 255);
        if (!jac_is_inf(j)) s = 3;
    }
    g1_to_uncompressed(b, a);
    for (int t = 0; t < 96; t++) pk96[96ull * i + t] = (s == 0) ? b[t] : 0;
    st[i] = s;
}

// Signature::deserialize (blst.rs:192-194): decompress only (no subgroup check); out 192-byte affine x.c1|x.c0|y.c1|y.c0
__global__ void __launch_bounds__(BLS_BLOCK) k_g2_decompress(const uint8_t* __restrict__ sig96, uint32_t n,
                                                              uint8_t* __restrict__ out192, uint8_t* __restrict__ st) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t b[96];
    for (int t = 0; t < 96; t++) b[t] = sig96[96ull * i + t];
    G2Affine a;
    const int32_t rc = g2_decompress(a, b);
    uint8_t o[192];
    for (int t = 0; t < 192; t++) o[t] = 0;
    if (rc == DEC_OK) {
        Fp c;
        fp_from_mont(c, a.x.c1); fp_to_be48(o, c);
        fp_from_mont(c, a.x.c0); fp_to_be48(o + 48, c);
        fp_from_mont(c, a.y.c1); fp_to_be48(o + 96, c);
        fp_from_mont(c, a.y.c0); fp_to_be48(o + 144, c);
    } else if (rc == DEC_INFINITY) {
        o[0] = 0x40;
    }
    for (int t = 0; t < 192; t++) out192[192ull * i + t] = o[t];
    st[i] = (uint8_t)rc;
}

}  // namespace bls
}  // namespace lhb200
