// sha256.cuh — SHA-256 device primitives for SSZ merkleization on sm_100a.
//
// Replaces ethereum_hashing::{hash32_concat, hash} (crate not vendored; call sites
// /root/reference/consensus/merkle_proof/src/lib.rs:1,91,149,380-384 and every tree_hash user).
//
// A Merkle node is SHA256(left || right) over exactly 64 bytes: one data block plus one CONSTANT
// padding block (0x80, zeros, bit length 512).  The padding block's whole 64-word message schedule is
// a compile-time constant, so K[t]+W[t] folds into one immediate per round and the second compression
// does no schedule work.  Everything is 32-bit ALU work (LOP3 / SHF / IADD3): no tensor-core content.
#pragma once
#include <stdint.h>

// LHB_HOSTSIM: the same code compiled for the host by the tests (never part of the product library).
#ifdef LHB_HOSTSIM
#define LHB_SHA_CONSTEXPR constexpr
#define LHB_SHA_FN static inline
#define LHB_SHA_NOINLINE static __attribute__((noinline))
#else
#define LHB_SHA_CONSTEXPR __host__ __device__ constexpr
#define LHB_SHA_FN __device__ __forceinline__
#define LHB_SHA_NOINLINE static __device__ __noinline__
#endif

namespace lhb200 {

LHB_SHA_CONSTEXPR uint32_t c_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

struct Sha256Consts {
    uint32_t k[64];
    uint32_t kw_pad[64];  // K[t] + W[t] for the constant 64-byte-message padding block
};

LHB_SHA_CONSTEXPR Sha256Consts make_sha256_consts() {
    Sha256Consts c{};
    const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
        0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
        0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
        0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
        0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
        0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
        0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t w[64] = {};
    w[0] = 0x80000000u;
    w[15] = 512;
    for (int t = 16; t < 64; t++) {
        uint32_t s0 = c_rotr(w[t - 15], 7) ^ c_rotr(w[t - 15], 18) ^ (w[t - 15] >> 3);
        uint32_t s1 = c_rotr(w[t - 2], 17) ^ c_rotr(w[t - 2], 19) ^ (w[t - 2] >> 10);
        w[t] = w[t - 16] + s0 + w[t - 7] + s1;
    }
    for (int t = 0; t < 64; t++) {
        c.k[t] = K[t];
        c.kw_pad[t] = K[t] + w[t];
    }
    return c;
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
#else
static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static inline uint32_t bswap32(uint32_t x) { return __builtin_bswap32(x); }
#endif

#define LHB_SHA_ROUND(a, b, c, d, e, f, g, h, kw)                         \
    {                                                                     \
        uint32_t t1 = (h) + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) +    \
                      (((e) & (f)) ^ (~(e) & (g))) + (kw);                \
        uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) +          \
                      (((a) & (b)) ^ ((a) & (c)) ^ ((b) & (c)));          \
        (d) += t1;                                                        \
        (h) = t1 + t2;                                                    \
    }

// One compression of a data block held in w[16] (big-endian-decoded words).  w is clobbered.
LHB_SHA_FN void sha256_compress(uint32_t st[8], uint32_t w[16]) {
    constexpr Sha256Consts C = make_sha256_consts();
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int t = 0; t < 64; t += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int i = t + j;
            if (i >= 16) {
                uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
                uint32_t s0 = rotr(w15, 7) ^ rotr(w15, 18) ^ (w15 >> 3);
                uint32_t s1 = rotr(w2, 17) ^ rotr(w2, 19) ^ (w2 >> 10);
                w[i & 15] = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
            }
        }
        LHB_SHA_ROUND(a, b, c, d, e, f, g, h, C.k[t + 0] + w[(t + 0) & 15]);
        LHB_SHA_ROUND(h, a, b, c, d, e, f, g, C.k[t + 1] + w[(t + 1) & 15]);
        LHB_SHA_ROUND(g, h, a, b, c, d, e, f, C.k[t + 2] + w[(t + 2) & 15]);
        LHB_SHA_ROUND(f, g, h, a, b, c, d, e, C.k[t + 3] + w[(t + 3) & 15]);
        LHB_SHA_ROUND(e, f, g, h, a, b, c, d, C.k[t + 4] + w[(t + 4) & 15]);
        LHB_SHA_ROUND(d, e, f, g, h, a, b, c, C.k[t + 5] + w[(t + 5) & 15]);
        LHB_SHA_ROUND(c, d, e, f, g, h, a, b, C.k[t + 6] + w[(t + 6) & 15]);
        LHB_SHA_ROUND(b, c, d, e, f, g, h, a, C.k[t + 7] + w[(t + 7) & 15]);
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// Compression of the constant padding block that follows a 64-byte message.
LHB_SHA_FN void sha256_compress_pad64(uint32_t st[8]) {
    constexpr Sha256Consts C = make_sha256_consts();
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int t = 0; t < 64; t += 8) {
        LHB_SHA_ROUND(a, b, c, d, e, f, g, h, C.kw_pad[t + 0]);
        LHB_SHA_ROUND(h, a, b, c, d, e, f, g, C.kw_pad[t + 1]);
        LHB_SHA_ROUND(g, h, a, b, c, d, e, f, C.kw_pad[t + 2]);
        LHB_SHA_ROUND(f, g, h, a, b, c, d, e, C.kw_pad[t + 3]);
        LHB_SHA_ROUND(e, f, g, h, a, b, c, d, C.kw_pad[t + 4]);
        LHB_SHA_ROUND(d, e, f, g, h, a, b, c, C.kw_pad[t + 5]);
        LHB_SHA_ROUND(c, d, e, f, g, h, a, b, C.kw_pad[t + 6]);
        LHB_SHA_ROUND(b, c, d, e, f, g, h, a, C.kw_pad[t + 7]);
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

LHB_SHA_FN void sha256_init(uint32_t st[8]) {
    st[0] = 0x6a09e667; st[1] = 0xbb67ae85; st[2] = 0x3c6ef372; st[3] = 0xa54ff53a;
    st[4] = 0x510e527f; st[5] = 0x9b05688c; st[6] = 0x1f83d9ab; st[7] = 0x5be0cd19;
}

// hash32_concat on words: out = SHA256(l || r), all in big-endian-decoded word form.  out may alias l or r.
LHB_SHA_FN void hash_pair_inl(const uint32_t l[8], const uint32_t r[8], uint32_t out[8]) {
    uint32_t w[16], st[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { w[i] = l[i]; w[8 + i] = r[i]; }
    sha256_init(st);
    sha256_compress(st, w);
    sha256_compress_pad64(st);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = st[i];
}

// Out-of-line copy for cold code (tails, hash programs, expand_message_xmd) to bound code size.
LHB_SHA_NOINLINE void hash_pair(const uint32_t* l, const uint32_t* r, uint32_t* out) {
    uint32_t a[8], b[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = l[i]; b[i] = r[i]; }
    hash_pair_inl(a, b, o);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = o[i];
}

#ifndef LHB_HOSTSIM
// 32-byte chunk <-> 8 big-endian-decoded words.  p must be 16-byte aligned.
__device__ __forceinline__ void load_chunk(const uint8_t* p, uint32_t w[8]) {
    uint4 x = __ldg(reinterpret_cast<const uint4*>(p));
    uint4 y = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    w[0] = bswap32(x.x); w[1] = bswap32(x.y); w[2] = bswap32(x.z); w[3] = bswap32(x.w);
    w[4] = bswap32(y.x); w[5] = bswap32(y.y); w[6] = bswap32(y.z); w[7] = bswap32(y.w);
}
__device__ __forceinline__ void store_chunk(uint8_t* p, const uint32_t w[8]) {
    uint4 x = make_uint4(bswap32(w[0]), bswap32(w[1]), bswap32(w[2]), bswap32(w[3]));
    uint4 y = make_uint4(bswap32(w[4]), bswap32(w[5]), bswap32(w[6]), bswap32(w[7]));
    reinterpret_cast<uint4*>(p)[0] = x;
    reinterpret_cast<uint4*>(p)[1] = y;
}
#endif  // !LHB_HOSTSIM

// Generic SHA-256 of a short byte string (expand_message_xmd; not a throughput path).
LHB_SHA_NOINLINE void sha256_block_oob(uint32_t* st, const uint8_t* blk) {
    uint32_t w[16], s[8];
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) |
               blk[4 * i + 3];
    for (int i = 0; i < 8; i++) s[i] = st[i];
    sha256_compress(s, w);
    for (int i = 0; i < 8; i++) st[i] = s[i];
}
// msg length must be <= 183 bytes (3 blocks); out = 32 bytes
LHB_SHA_FN void sha256_short(const uint8_t* msg, int len, uint8_t* out) {
    uint8_t buf[192];
    const int nblk = (len + 9 + 63) / 64;
    for (int i = 0; i < nblk * 64; i++) buf[i] = i < len ? msg[i] : 0;
    buf[len] = 0x80;
    const uint32_t bits = (uint32_t)len * 8;
    buf[nblk * 64 - 1] = (uint8_t)bits;
    buf[nblk * 64 - 2] = (uint8_t)(bits >> 8);
    uint32_t st[8];
    sha256_init(st);
    for (int b = 0; b < nblk; b++) sha256_block_oob(st, buf + 64 * b);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = st[i] >> 24; out[4 * i + 1] = st[i] >> 16; out[4 * i + 2] = st[i] >> 8; out[4 * i + 3] = st[i];
    }
}

}  // namespace lhb200
