/*
 * oracle/ssz_sha256.c — TEST INFRASTRUCTURE ONLY (CPU oracle / CPU baseline).
 *
 * Plain-C restatement of the SHA-256 Merkle `tree_hash_root` path of sigp/lighthouse v5.3.0.
 * Nothing in lighthouse_b200/ (the product) may link, import or call this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * The arithmetic of this path lives in crates that are NOT vendored under /root/reference
 * (Cargo.lock pins): ethereum_hashing 0.6.0 (sha2 0.10.8 / ring 0.17.8), tree_hash 0.6.0,
 * tree_hash_derive 0.6.0, milhouse 0.1.0, ssz_types 0.6.0.  Their published algorithm (FIPS 180-4
 * SHA-256 + the consensus-spec SSZ merkleization rules) is restated here and anchored on the
 * reference's own call sites and vectors:
 *   hash32_concat / ZERO_HASHES ........ consensus/merkle_proof/src/lib.rs:1,91,166
 *   MerkleTree::create .................. consensus/merkle_proof/src/lib.rs:68-99
 *   merkle_root_from_branch ............. consensus/merkle_proof/src/lib.rs:372-389
 *   Validator (8 leaves) ................ consensus/types/src/validator.rs:25-35
 *   48/96-byte BLS blob roots ........... crypto/bls/src/macros.rs:4-27
 *   BeaconState (Deneb, 28 fields) ...... consensus/types/src/beacon_state.rs:339-490, :2031-2038
 *   ExecutionPayloadHeaderDeneb ......... consensus/types/src/execution_payload_header.rs:46-87
 *   mainnet sizes ....................... consensus/types/src/eth_spec.rs:389-430
 * Pinned by (tests/test_oracle_merkle.py): hashlib SHA-256, the genesis_validators_root embedded in the
 * reference's vendored mainnet/sepolia/gnosis genesis.ssz.zip, and deposit_message_root/deposit_data_root
 * of validator_manager/test_vectors.  Full Deneb BeaconState root: parity unpinned in-tree (only EF
 * ssz_static pins it; not on disk) — pinned transitively through the per-field rules above AND, field root by
 * field root, against an independent generic from-spec merkleization (tests/ssz_spec.py driven by the type
 * descriptors of lighthouse_b200/ssz_schema.py; tests/test_oracle_merkle.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#include <cpuid.h>
#endif

#define EXPORT __attribute__((visibility("default")))


/* ------------------------------------------------------------------ tiny pthread parallel-for
 * (the image has no libgomp).  Mirrors rayon::join fan-out in milhouse / blst's pool: static ranges. */
#include <pthread.h>
#include <unistd.h>
static int g_threads = 1;
typedef void (*range_fn)(uint64_t lo, uint64_t hi, void *ctx);
struct par_job { range_fn fn; void *ctx; uint64_t lo, hi; };
static void *par_tramp(void *p) { struct par_job *j = (struct par_job *)p; j->fn(j->lo, j->hi, j->ctx); return 0; }
void orc_par_for(uint64_t n, uint64_t min_grain, range_fn fn, void *ctx) {
    int t = g_threads;
    if (t < 1) t = 1;
    if (n < min_grain * 2 || t == 1) { fn(0, n, ctx); return; }
    if ((uint64_t)t > n / min_grain) t = (int)(n / min_grain);
    pthread_t th[256]; struct par_job jb[256];
    if (t > 256) t = 256;
    for (int i = 0; i < t; i++) {
        jb[i].fn = fn; jb[i].ctx = ctx; jb[i].lo = n * i / t; jb[i].hi = n * (i + 1) / t;
        if (i) pthread_create(&th[i], 0, par_tramp, &jb[i]);
    }
    par_tramp(&jb[0]);
    for (int i = 1; i < t; i++) pthread_join(th[i], 0);
}
EXPORT int orc_num_threads(void) { return g_threads; }
EXPORT int orc_hw_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }
EXPORT void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

/* ------------------------------------------------------------------ SHA-256 (FIPS 180-4) */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
static const uint32_t IV256[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                  0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};

static inline uint32_t ror32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void compress_plain(uint32_t st[8], const uint8_t blk[64]) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) |
               blk[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = ror32(w[i - 15], 7) ^ ror32(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = ror32(w[i - 2], 17) ^ ror32(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = ror32(e, 6) ^ ror32(e, 11) ^ ror32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = h + S1 + ch + K256[i] + w[i];
        uint32_t S0 = ror32(a, 2) ^ ror32(a, 13) ^ ror32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

#if defined(__x86_64__)
/* SHA-NI path: what ethereum_hashing selects on this class of CPU (lighthouse/src/main.rs:15,41 checks
 * have_sha_extensions()).  Used so the CPU baseline is not handicapped; verified against compress_plain. */
__attribute__((target("sha,sse4.1,ssse3"))) static void compress_shani(uint32_t st[8], const uint8_t blk[64]) {
    const __m128i MASK = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i TMP = _mm_loadu_si128((const __m128i *)&st[0]);
    __m128i STATE1 = _mm_loadu_si128((const __m128i *)&st[4]);
    TMP = _mm_shuffle_epi32(TMP, 0xB1);          /* CDAB */
    STATE1 = _mm_shuffle_epi32(STATE1, 0x1B);    /* EFGH */
    __m128i STATE0 = _mm_alignr_epi8(TMP, STATE1, 8); /* ABEF */
    STATE1 = _mm_blend_epi16(STATE1, TMP, 0xF0);      /* CDGH */
    __m128i ABEF_SAVE = STATE0, CDGH_SAVE = STATE1;
    __m128i M[4];
    for (int i = 0; i < 4; i++) M[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *)(blk + 16 * i)), MASK);
    for (int r = 0; r < 16; r++) {
        __m128i MSG = _mm_add_epi32(M[r & 3], _mm_loadu_si128((const __m128i *)&K256[4 * r]));
        STATE1 = _mm_sha256rnds2_epu32(STATE1, STATE0, MSG);
        if (r >= 3 && r < 15) { /* finish w[4(r+1)..] = msg2(msg1(M[r+1]) + alignr(M[r],M[r-1]), M[r]) */
            __m128i T = _mm_alignr_epi8(M[r & 3], M[(r - 1) & 3], 4);
            M[(r + 1) & 3] = _mm_sha256msg2_epu32(_mm_add_epi32(M[(r + 1) & 3], T), M[r & 3]);
        }
        MSG = _mm_shuffle_epi32(MSG, 0x0E);
        STATE0 = _mm_sha256rnds2_epu32(STATE0, STATE1, MSG);
        if (r >= 1 && r < 13) M[(r - 1) & 3] = _mm_sha256msg1_epu32(M[(r - 1) & 3], M[r & 3]);
    }
    STATE0 = _mm_add_epi32(STATE0, ABEF_SAVE);
    STATE1 = _mm_add_epi32(STATE1, CDGH_SAVE);
    TMP = _mm_shuffle_epi32(STATE0, 0x1B);       /* FEBA */
    STATE1 = _mm_shuffle_epi32(STATE1, 0xB1);    /* DCHG */
    STATE0 = _mm_blend_epi16(TMP, STATE1, 0xF0); /* DCBA */
    STATE1 = _mm_alignr_epi8(STATE1, TMP, 8);    /* ABEF */
    _mm_storeu_si128((__m128i *)&st[0], STATE0);
    _mm_storeu_si128((__m128i *)&st[4], STATE1);
}
#endif

static int g_use_shani = -1;
static void (*g_compress)(uint32_t *, const uint8_t *) = compress_plain;

EXPORT int orc_sha_backend(int force_plain) {
    /* returns 1 if SHA-NI is in use */
    g_compress = compress_plain;
    g_use_shani = 0;
#if defined(__x86_64__)
    if (!force_plain) {
        unsigned a, b, c, d;
        if (__get_cpuid_count(7, 0, &a, &b, &c, &d) && (b & (1u << 29))) {
            g_compress = compress_shani;
            g_use_shani = 1;
        }
    }
#endif
    return g_use_shani;
}

static inline void ensure_backend(void) {
    if (g_use_shani < 0) orc_sha_backend(0);
}

static void put_be32(uint8_t *p, uint32_t v) {
    p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v;
}

/* generic SHA-256 of a byte string (ethereum_hashing::hash) */
EXPORT void orc_sha256(const uint8_t *msg, uint64_t len, uint8_t out[32]) {
    ensure_backend();
    uint32_t st[8];
    memcpy(st, IV256, sizeof st);
    uint64_t i = 0;
    for (; i + 64 <= len; i += 64) g_compress(st, msg + i);
    uint8_t blk[128];
    memset(blk, 0, sizeof blk);
    uint64_t rem = len - i;
    memcpy(blk, msg + i, rem);
    blk[rem] = 0x80;
    int nb = (rem + 9 > 64) ? 2 : 1;
    uint64_t bits = len * 8;
    for (int k = 0; k < 8; k++) blk[64 * nb - 1 - k] = (uint8_t)(bits >> (8 * k));
    g_compress(st, blk);
    if (nb == 2) g_compress(st, blk + 64);
    for (int k = 0; k < 8; k++) put_be32(out + 4 * k, st[k]);
}

/* ethereum_hashing::hash32_concat(a, b) = SHA256(a || b): one data block + one constant padding block */
static const uint8_t PAD64[64] = {0x80, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                  0,    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                  0,    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x02, 0x00};

static inline void hash64(const uint8_t in[64], uint8_t out[32]) {
    uint32_t st[8];
    memcpy(st, IV256, sizeof st);
    g_compress(st, in);
    g_compress(st, PAD64);
    for (int k = 0; k < 8; k++) put_be32(out + 4 * k, st[k]);
}

EXPORT void orc_hash32_concat(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
    ensure_backend();
    uint8_t in[64];
    memcpy(in, a, 32);
    memcpy(in + 32, b, 32);
    hash64(in, out);
}

struct hp_ctx { const uint8_t *in; uint8_t *out; };
static void hp_range(uint64_t lo, uint64_t hi, void *c) {
    struct hp_ctx *x = (struct hp_ctx *)c;
    for (uint64_t i = lo; i < hi; i++) hash64(x->in + 64 * i, x->out + 32 * i);
}
EXPORT void orc_hash_pairs(const uint8_t *in, uint8_t *out, uint64_t n) {
    ensure_backend();
    struct hp_ctx c = {in, out};
    orc_par_for(n, 2048, hp_range, &c);
}

/* ZERO_HASHES[d] : root of an all-zero subtree of height d (merkle_proof/src/lib.rs:166) */
#define MAX_ZERO 65
static uint8_t ZH[MAX_ZERO][32];
static int zh_ready = 0;
static void ensure_zero(void) {
    ensure_backend();
    if (zh_ready) return;
    memset(ZH[0], 0, 32);
    for (int d = 0; d + 1 < MAX_ZERO; d++) orc_hash32_concat(ZH[d], ZH[d], ZH[d + 1]);
    zh_ready = 1;
}
EXPORT void orc_zero_hash(uint32_t depth, uint8_t out[32]) {
    ensure_zero();
    memcpy(out, ZH[depth], 32);
}

/* merkleize(chunks, limit=2^depth): pad every level's odd tail with zero[level]; empty => zero[depth].
 * `work` (n*32 bytes scratch) may alias nothing. */
EXPORT void orc_merkleize(const uint8_t *chunks, uint64_t n, uint32_t depth, uint8_t out[32]) {
    ensure_zero();
    if (n == 0) { memcpy(out, ZH[depth], 32); return; }
    uint8_t *cur = (uint8_t *)malloc((size_t)(n + 1) * 32);
    memcpy(cur, chunks, (size_t)n * 32);
    uint64_t cnt = n;
    for (uint32_t lvl = 0; lvl < depth; lvl++) {
        if (cnt == 1) { /* lone node: climb with zero siblings */
            orc_hash32_concat(cur, ZH[lvl], cur);
            continue;
        }
        if (cnt & 1) { memcpy(cur + cnt * 32, ZH[lvl], 32); cnt++; }
        uint64_t half = cnt / 2;
        /* in-place is safe when processed in increasing order serially; for the parallel case use a copy */
        if (half > 4096) {
            uint8_t *nxt = (uint8_t *)malloc((size_t)(half + 1) * 32);
            struct hp_ctx c = {cur, nxt};
            orc_par_for(half, 2048, hp_range, &c);
            free(cur);
            cur = nxt;
        } else {
            for (uint64_t i = 0; i < half; i++) {
                uint8_t tmp[32];
                hash64(cur + 64 * i, tmp);
                memcpy(cur + 32 * i, tmp, 32);
            }
        }
        cnt = half;
    }
    memcpy(out, cur, 32);
    free(cur);
}

EXPORT void orc_mix_in_length(const uint8_t root[32], uint64_t len, uint8_t out[32]) {
    uint8_t l[32];
    memset(l, 0, 32);
    for (int k = 0; k < 8; k++) l[k] = (uint8_t)(len >> (8 * k));
    orc_hash32_concat(root, l, out);
}

static uint32_t ceil_log2(uint64_t x) {
    uint32_t d = 0;
    while (((uint64_t)1 << d) < x) d++;
    return d;
}

/* packed basic list/vector: bytes are the little-endian serialisation; last chunk zero padded */
EXPORT void orc_merkleize_bytes(const uint8_t *bytes, uint64_t nbytes, uint32_t depth, uint8_t out[32]) {
    uint64_t n = (nbytes + 31) / 32;
    uint8_t *buf = (uint8_t *)calloc((size_t)(n ? n : 1), 32);
    memcpy(buf, bytes, (size_t)nbytes);
    orc_merkleize(buf, n, depth, out);
    free(buf);
}

/* 48-byte pubkey root: merkle_root of 2 chunks (crypto/bls/src/macros.rs:18-25) */
static void pubkey_root(const uint8_t pk[48], uint8_t out[32]) {
    uint8_t in[64];
    memset(in, 0, 64);
    memcpy(in, pk, 48);
    hash64(in, out);
}
EXPORT void orc_pubkey_root(const uint8_t pk[48], uint8_t out[32]) { ensure_backend(); pubkey_root(pk, out); }

/* Validator: 121-byte SSZ -> 8 leaves -> root (consensus/types/src/validator.rs:25-35) */
static void validator_root(const uint8_t *v, uint8_t out[32]) {
    uint8_t leaf[8][32];
    memset(leaf, 0, sizeof leaf);
    pubkey_root(v, leaf[0]);
    memcpy(leaf[1], v + 48, 32);
    memcpy(leaf[2], v + 80, 8);
    leaf[3][0] = v[88];
    memcpy(leaf[4], v + 89, 8);
    memcpy(leaf[5], v + 97, 8);
    memcpy(leaf[6], v + 105, 8);
    memcpy(leaf[7], v + 113, 8);
    uint8_t l1[4][32], l2[2][32];
    for (int i = 0; i < 4; i++) hash64((const uint8_t *)leaf + 64 * i, l1[i]);
    for (int i = 0; i < 2; i++) hash64((const uint8_t *)l1 + 64 * i, l2[i]);
    hash64((const uint8_t *)l2, out);
}
EXPORT void orc_validator_root(const uint8_t v[121], uint8_t out[32]) { ensure_backend(); validator_root(v, out); }

static void vr_range(uint64_t lo, uint64_t hi, void *c) {
    struct hp_ctx *x = (struct hp_ctx *)c;
    for (uint64_t i = lo; i < hi; i++) validator_root(x->in + 121 * i, x->out + 32 * i);
}
EXPORT void orc_validator_roots(const uint8_t *ssz, uint64_t n, uint8_t *out) {
    ensure_backend();
    struct hp_ctx c = {ssz, out};
    orc_par_for(n, 512, vr_range, &c);
}

/* List[Validator, 2^40] */
EXPORT void orc_validators_root(const uint8_t *ssz, uint64_t n, uint8_t out[32]) {
    uint8_t *roots = (uint8_t *)malloc((size_t)(n ? n : 1) * 32);
    orc_validator_roots(ssz, n, roots);
    uint8_t r[32];
    orc_merkleize(roots, n, 40, r);
    orc_mix_in_length(r, n, out);
    free(roots);
}

/* container of k field roots -> merkleize over next_pow2(k) */
static void container_root(uint8_t (*leaves)[32], uint32_t k, uint8_t out[32]) {
    orc_merkleize(&leaves[0][0], k, ceil_log2(k), out);
}

static void u64_chunk(const uint8_t *p, uint8_t out[32]) { memset(out, 0, 32); memcpy(out, p, 8); }
static uint32_t rd32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

static void checkpoint_root(const uint8_t *p, uint8_t out[32]) { /* {epoch u64, root H256} */
    uint8_t l[2][32];
    u64_chunk(p, l[0]);
    memcpy(l[1], p + 8, 32);
    container_root(l, 2, out);
}
static void eth1_data_root(const uint8_t *p, uint8_t out[32]) { /* {H256, u64, H256} (eth1_data.rs:27) */
    uint8_t l[3][32];
    memcpy(l[0], p, 32);
    u64_chunk(p + 32, l[1]);
    memcpy(l[2], p + 40, 32);
    container_root(l, 3, out);
}
static void sync_committee_root(const uint8_t *p, uint8_t out[32]) { /* {Vector[pubkey,512], pubkey} */
    uint8_t *pk = (uint8_t *)malloc(512 * 32);
    for (int i = 0; i < 512; i++) pubkey_root(p + 48 * i, pk + 32 * i);
    uint8_t l[2][32];
    orc_merkleize(pk, 512, 9, l[0]);
    pubkey_root(p + 48 * 512, l[1]);
    container_root(l, 2, out);
    free(pk);
}

/* ExecutionPayloadHeaderDeneb: 17 fields (execution_payload_header.rs:46-87) */
static int exec_header_root(const uint8_t *p, uint64_t len, uint8_t out[32]) {
    if (len < 584) return -1;
    uint8_t l[17][32];
    memset(l, 0, sizeof l);
    memcpy(l[0], p, 32);                       /* parent_hash */
    memcpy(l[1], p + 32, 20);                  /* fee_recipient (Address) */
    memcpy(l[2], p + 52, 32);                  /* state_root */
    memcpy(l[3], p + 84, 32);                  /* receipts_root */
    orc_merkleize(p + 116, 8, 3, l[4]);        /* logs_bloom: 256 B = 8 chunks */
    memcpy(l[5], p + 372, 32);                 /* prev_randao */
    memcpy(l[6], p + 404, 8);                  /* block_number */
    memcpy(l[7], p + 412, 8);                  /* gas_limit */
    memcpy(l[8], p + 420, 8);                  /* gas_used */
    memcpy(l[9], p + 428, 8);                  /* timestamp */
    uint32_t off = rd32(p + 436);              /* extra_data offset */
    memcpy(l[11], p + 440, 32);                /* base_fee_per_gas (u256 LE) */
    memcpy(l[12], p + 472, 32);                /* block_hash */
    memcpy(l[13], p + 504, 32);                /* transactions_root */
    memcpy(l[14], p + 536, 32);                /* withdrawals_root */
    memcpy(l[15], p + 568, 8);                 /* blob_gas_used */
    memcpy(l[16], p + 576, 8);                 /* excess_blob_gas */
    if (off != 584 || len - off > 32) return -2;
    uint8_t r[32];
    orc_merkleize_bytes(p + off, len - off, 0, r); /* List[u8,32]: one chunk */
    orc_mix_in_length(r, len - off, l[10]);
    container_root(l, 17, out);
    return 0;
}

#define DENEB_FIXED 2736653u

/* hash_tree_root(BeaconStateDeneb) from its SSZ bytes (mainnet preset). field_roots (28*32) optional. */
EXPORT int orc_beacon_state_root_deneb(const uint8_t *s, uint64_t len, uint8_t out[32], uint8_t *field_roots) {
    ensure_zero();
    if (len < DENEB_FIXED) return -1;
    uint8_t f[28][32];
    memset(f, 0, sizeof f);
    uint32_t o_hist = rd32(s + 524464), o_votes = rd32(s + 524540), o_val = rd32(s + 524552),
             o_bal = rd32(s + 524556), o_pp = rd32(s + 2687248), o_cp = rd32(s + 2687252),
             o_inact = rd32(s + 2687377), o_leph = rd32(s + 2736629), o_hs = rd32(s + 2736649);
    if (o_hist != DENEB_FIXED || !(o_hist <= o_votes && o_votes <= o_val && o_val <= o_bal && o_bal <= o_pp &&
                                   o_pp <= o_cp && o_cp <= o_inact && o_inact <= o_leph && o_leph <= o_hs &&
                                   o_hs <= len))
        return -2;
    u64_chunk(s + 0, f[0]);
    memcpy(f[1], s + 8, 32);
    u64_chunk(s + 40, f[2]);
    { /* Fork {[u8;4],[u8;4],u64} (fork.rs:26) */
        uint8_t l[3][32];
        memset(l, 0, sizeof l);
        memcpy(l[0], s + 48, 4);
        memcpy(l[1], s + 52, 4);
        memcpy(l[2], s + 56, 8);
        container_root(l, 3, f[3]);
    }
    { /* BeaconBlockHeader {slot, proposer_index, parent_root, state_root, body_root} */
        uint8_t l[5][32];
        u64_chunk(s + 64, l[0]);
        u64_chunk(s + 72, l[1]);
        memcpy(l[2], s + 80, 32);
        memcpy(l[3], s + 112, 32);
        memcpy(l[4], s + 144, 32);
        container_root(l, 5, f[4]);
    }
    orc_merkleize(s + 176, 8192, 13, f[5]);
    orc_merkleize(s + 262320, 8192, 13, f[6]);
    { /* historical_roots: List[H256, 2^24] */
        uint64_t n = (o_votes - o_hist) / 32;
        uint8_t r[32];
        orc_merkleize(s + o_hist, n, 24, r);
        orc_mix_in_length(r, n, f[7]);
    }
    eth1_data_root(s + 524468, f[8]);
    { /* eth1_data_votes: List[Eth1Data, 2048] */
        uint64_t n = (o_val - o_votes) / 72;
        uint8_t *r = (uint8_t *)malloc((size_t)(n ? n : 1) * 32), t[32];
        for (uint64_t i = 0; i < n; i++) eth1_data_root(s + o_votes + 72 * i, r + 32 * i);
        orc_merkleize(r, n, 11, t);
        orc_mix_in_length(t, n, f[9]);
        free(r);
    }
    u64_chunk(s + 524544, f[10]);
    orc_validators_root(s + o_val, (o_bal - o_val) / 121, f[11]);
    { /* balances: List[u64, 2^40] -> chunk depth 38 */
        uint64_t n = (o_pp - o_bal) / 8;
        uint8_t r[32];
        orc_merkleize_bytes(s + o_bal, n * 8, 38, r);
        orc_mix_in_length(r, n, f[12]);
    }
    orc_merkleize(s + 524560, 65536, 16, f[13]);
    orc_merkleize_bytes(s + 2621712, 65536, 11, f[14]); /* slashings: 8192 u64 = 2048 chunks */
    { /* participation: List[u8, 2^40] -> chunk depth 35 */
        uint64_t n = o_cp - o_pp;
        uint8_t r[32];
        orc_merkleize_bytes(s + o_pp, n, 35, r);
        orc_mix_in_length(r, n, f[15]);
        n = o_inact - o_cp;
        orc_merkleize_bytes(s + o_cp, n, 35, r);
        orc_mix_in_length(r, n, f[16]);
    }
    f[17][0] = s[2687256]; /* Bitvector[4] */
    checkpoint_root(s + 2687257, f[18]);
    checkpoint_root(s + 2687297, f[19]);
    checkpoint_root(s + 2687337, f[20]);
    {
        uint64_t n = (o_leph - o_inact) / 8;
        uint8_t r[32];
        orc_merkleize_bytes(s + o_inact, n * 8, 38, r);
        orc_mix_in_length(r, n, f[21]);
    }
    sync_committee_root(s + 2687381, f[22]);
    sync_committee_root(s + 2712005, f[23]);
    if (exec_header_root(s + o_leph, o_hs - o_leph, f[24])) return -3;
    u64_chunk(s + 2736633, f[25]);
    u64_chunk(s + 2736641, f[26]);
    { /* historical_summaries: List[{H256,H256}, 2^24] */
        uint64_t n = (len - o_hs) / 64;
        uint8_t *r = (uint8_t *)malloc((size_t)(n ? n : 1) * 32), t[32];
        orc_hash_pairs(s + o_hs, r, n);
        orc_merkleize(r, n, 24, t);
        orc_mix_in_length(t, n, f[27]);
        free(r);
    }
    if (field_roots) memcpy(field_roots, f, sizeof f);
    container_root(f, 28, out);
    return 0;
}

/* merkle_root_from_branch (consensus/merkle_proof/src/lib.rs:372-389) */
EXPORT void orc_merkle_root_from_branch(const uint8_t leaf[32], const uint8_t *branch, uint32_t depth, uint64_t index,
                                        uint8_t out[32]) {
    uint8_t cur[32];
    memcpy(cur, leaf, 32);
    for (uint32_t i = 0; i < depth; i++) {
        if ((index >> i) & 1) orc_hash32_concat(branch + 32 * i, cur, cur);
        else orc_hash32_concat(cur, branch + 32 * i, cur);
    }
    memcpy(out, cur, 32);
}

/* MerkleTree::create(leaves, depth).hash() + generate_proof(index, depth) (merkle_proof/src/lib.rs:68-99,290-324).
 * Returns the root and the bottom-up branch of `depth` siblings. */
EXPORT void orc_merkle_tree_proof(const uint8_t *leaves, uint64_t n, uint32_t depth, uint64_t index, uint8_t root[32],
                                  uint8_t *branch) {
    ensure_zero();
    uint8_t *cur = (uint8_t *)malloc((size_t)(n + 2) * 32);
    memcpy(cur, leaves, (size_t)n * 32);
    uint64_t cnt = n, idx = index;
    for (uint32_t lvl = 0; lvl < depth; lvl++) {
        uint64_t sib = idx ^ 1;
        if (sib < cnt) memcpy(branch + 32 * lvl, cur + 32 * sib, 32);
        else memcpy(branch + 32 * lvl, ZH[lvl], 32);
        if (cnt & 1) { memcpy(cur + cnt * 32, ZH[lvl], 32); cnt++; }
        uint64_t half = cnt / 2;
        for (uint64_t i = 0; i < half; i++) {
            uint8_t tmp[32];
            hash64(cur + 64 * i, tmp);
            memcpy(cur + 32 * i, tmp, 32);
        }
        cnt = half;
        idx >>= 1;
    }
    if (n == 0) memcpy(root, ZH[depth], 32);
    else memcpy(root, cur, 32);
    free(cur);
}


/* ---------------------------------------------------------------------------------------------------------
 * swap-or-not shuffle (consensus/swap_or_not_shuffle): restatement of shuffle_list (src/shuffle_list.rs:79-160,
 * the in-place pivot/mirror sweep) and compute_shuffled_index (src/compute_shuffled_index.rs:20-58).
 * TEST INFRASTRUCTURE (SURVEY §8f-4 widening). */
static void shuffle_hash(const uint8_t seed[32], uint8_t round, int with_pos, uint32_t pos, uint8_t out[32]) {
    uint8_t buf[37];
    memcpy(buf, seed, 32);
    buf[32] = round;
    buf[33] = (uint8_t)pos; buf[34] = (uint8_t)(pos >> 8); buf[35] = (uint8_t)(pos >> 16); buf[36] = (uint8_t)(pos >> 24);
    orc_sha256(buf, with_pos ? 37 : 33, out);
}
static uint64_t le64(const uint8_t *p) {
    uint64_t v = 0;
    for (int k = 7; k >= 0; k--) v = (v << 8) | p[k];
    return v;
}
/* returns 0 on success, -1 where the reference returns None */
EXPORT int orc_shuffle_list(uint64_t *list, uint64_t n, uint8_t rounds, const uint8_t seed[32], int forwards) {
    if (n == 0 || n > (1ull << 24) || rounds == 0) return -1;
    uint8_t r = forwards ? 0 : (uint8_t)(rounds - 1), d[32], source[32];
    for (;;) {
        shuffle_hash(seed, r, 0, 0, d);
        uint64_t pivot = le64(d) % n;
        uint64_t mirror = (pivot + 1) >> 1;
        shuffle_hash(seed, r, 1, (uint32_t)(pivot >> 8), source);
        uint8_t byte_v = source[(pivot & 0xff) >> 3];
        for (uint64_t i = 0; i < mirror; i++) {
            uint64_t j = pivot - i;
            if ((j & 0xff) == 0xff) shuffle_hash(seed, r, 1, (uint32_t)(j >> 8), source);
            if ((j & 0x07) == 0x07) byte_v = source[(j & 0xff) >> 3];
            if ((byte_v >> (j & 0x07)) & 1) { uint64_t t = list[i]; list[i] = list[j]; list[j] = t; }
        }
        mirror = (pivot + n + 1) >> 1;
        uint64_t end = n - 1;
        shuffle_hash(seed, r, 1, (uint32_t)(end >> 8), source);
        byte_v = source[(end & 0xff) >> 3];
        uint64_t it = 0;
        for (uint64_t i = pivot + 1; i < mirror; i++, it++) {
            uint64_t j = end - it;
            if ((j & 0xff) == 0xff) shuffle_hash(seed, r, 1, (uint32_t)(j >> 8), source);
            if ((j & 0x07) == 0x07) byte_v = source[(j & 0xff) >> 3];
            if ((byte_v >> (j & 0x07)) & 1) { uint64_t t = list[i]; list[i] = list[j]; list[j] = t; }
        }
        if (forwards) { r++; if (r == rounds) break; }
        else { if (r == 0) break; r--; }
    }
    return 0;
}
EXPORT int64_t orc_compute_shuffled_index(uint64_t index, uint64_t n, const uint8_t seed[32], uint8_t rounds) {
    if (n == 0 || index >= n || n > (1ull << 24)) return -1;
    uint8_t d[32];
    for (uint8_t r = 0; r < rounds; r++) {
        shuffle_hash(seed, r, 0, 0, d);
        uint64_t pivot = le64(d) % n;
        uint64_t flip = (pivot + (n - index)) % n;
        uint64_t pos = index > flip ? index : flip;
        shuffle_hash(seed, r, 1, (uint32_t)(pos >> 8), d);
        if ((d[(pos & 0xff) >> 3] >> (pos & 7)) & 1) index = flip;
    }
    return (int64_t)index;
}

/* ---------------------------------------------------------------------------------------------------------
 * hash_tree_root(BeaconBlockDeneb) from SSZ bytes, mainnet preset:
 *   BeaconBlock ........................ consensus/types/src/beacon_block.rs:56-78, canonical_root :158-160
 *   BeaconBlockBodyDeneb (12 fields) ... consensus/types/src/beacon_block_body.rs:70-121
 *   ExecutionPayloadDeneb (17 fields) .. consensus/types/src/execution_payload.rs:54-95
 *   operations ......................... proposer_slashing.rs:26, attester_slashing.rs:42, indexed_attestation.rs:53,
 *                                        attestation.rs:74, deposit.rs:27, deposit_data.rs:25, signed_voluntary_exit.rs:25,
 *                                        sync_aggregate.rs:38, signed_bls_to_execution_change.rs:22, withdrawal.rs:22
 *   limits ............................. consensus/types/src/eth_spec.rs:394-432
 * Pinned by tests/test_block_root.py against tests/ssz_spec.py (generic from-spec hashlib merkleization).  The
 * reference's own pin (EF ssz_static) is not on disk: parity unpinned in-tree, pinned transitively. */
static void bytes_root(const uint8_t *p, uint64_t n, uint32_t depth, uint8_t out[32]) { orc_merkleize_bytes(p, n, depth, out); }
static void list_bytes_root(const uint8_t *p, uint64_t nbytes, uint32_t depth, uint64_t count, uint8_t out[32]) {
    uint8_t r[32];
    orc_merkleize_bytes(p, nbytes, depth, r);
    orc_mix_in_length(r, count, out);
}
static void att_data_root(const uint8_t *p, uint8_t out[32]) { /* 128 B */
    uint8_t l[5][32];
    u64_chunk(p, l[0]);
    u64_chunk(p + 8, l[1]);
    memcpy(l[2], p + 16, 32);
    checkpoint_root(p + 48, l[3]);
    checkpoint_root(p + 88, l[4]);
    container_root(l, 5, out);
}
static void signed_header_root(const uint8_t *p, uint8_t out[32]) { /* 208 B */
    uint8_t h[5][32], l[2][32];
    u64_chunk(p, h[0]);
    u64_chunk(p + 8, h[1]);
    memcpy(h[2], p + 16, 32);
    memcpy(h[3], p + 48, 32);
    memcpy(h[4], p + 80, 32);
    container_root(h, 5, l[0]);
    bytes_root(p + 112, 96, 2, l[1]);
    container_root(l, 2, out);
}
static int indexed_attestation_root(const uint8_t *p, uint64_t len, uint8_t out[32]) {
    if (len < 228 || rd32(p) != 228 || (len - 228) % 8 || (len - 228) / 8 > 2048) return -1;
    uint8_t l[3][32];
    list_bytes_root(p + 228, len - 228, 9, (len - 228) / 8, l[0]);
    att_data_root(p + 4, l[1]);
    bytes_root(p + 132, 96, 2, l[2]);
    container_root(l, 3, out);
    return 0;
}
static int attestation_root(const uint8_t *p, uint64_t len, uint8_t out[32]) {
    if (len < 229 || rd32(p) != 228 || p[len - 1] == 0) return -1;
    uint64_t nb = len - 228;
    uint8_t last = p[len - 1];
    int top = 7;
    while (!((last >> top) & 1)) top--;
    uint64_t bitlen = 8 * (nb - 1) + (uint64_t)top;
    if (bitlen > 2048) return -1;
    uint8_t bits[260];
    memcpy(bits, p + 228, nb);
    bits[nb - 1] = (uint8_t)(last & ~(1u << top));
    uint8_t l[3][32];
    list_bytes_root(bits, (bitlen + 7) / 8, 3, bitlen, l[0]);
    att_data_root(p + 4, l[1]);
    bytes_root(p + 132, 96, 2, l[2]);
    container_root(l, 3, out);
    return 0;
}
static void deposit_root(const uint8_t *p, uint8_t out[32]) { /* 1240 B */
    uint8_t l[2][32], d[4][32];
    orc_merkleize(p, 33, 6, l[0]);
    pubkey_root(p + 1056, d[0]);
    memcpy(d[1], p + 1104, 32);
    u64_chunk(p + 1136, d[2]);
    bytes_root(p + 1144, 96, 2, d[3]);
    container_root(d, 4, l[1]);
    container_root(l, 2, out);
}
static void exit_root(const uint8_t *p, uint8_t out[32]) { /* 112 B */
    uint8_t m[2][32], l[2][32];
    u64_chunk(p, m[0]);
    u64_chunk(p + 8, m[1]);
    container_root(m, 2, l[0]);
    bytes_root(p + 16, 96, 2, l[1]);
    container_root(l, 2, out);
}
static void bls_change_root(const uint8_t *p, uint8_t out[32]) { /* 172 B */
    uint8_t m[3][32], l[2][32];
    u64_chunk(p, m[0]);
    pubkey_root(p + 8, m[1]);
    memset(m[2], 0, 32);
    memcpy(m[2], p + 56, 20);
    container_root(m, 3, l[0]);
    bytes_root(p + 76, 96, 2, l[1]);
    container_root(l, 2, out);
}
static void withdrawal_root(const uint8_t *p, uint8_t out[32]) { /* 44 B */
    uint8_t l[4][32];
    u64_chunk(p, l[0]);
    u64_chunk(p + 8, l[1]);
    memset(l[2], 0, 32);
    memcpy(l[2], p + 16, 20);
    u64_chunk(p + 36, l[3]);
    container_root(l, 4, out);
}
/* list of variable-size items: region [p, p+len) starts with n 4-byte offsets */
static int var_list_bounds(const uint8_t *p, uint64_t len, uint64_t max_n, uint64_t *n, uint64_t **bounds) {
    *n = 0;
    *bounds = NULL;
    if (len == 0) return 0;
    if (len < 4) return -1;
    uint32_t first = rd32(p);
    if (first % 4 || first == 0 || first > len || first / 4 > max_n) return -1;
    uint64_t k = first / 4;
    uint64_t *b = (uint64_t *)malloc((k + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < k; i++) b[i] = rd32(p + 4 * i);
    b[k] = len;
    for (uint64_t i = 0; i < k; i++)
        if (b[i] > b[i + 1]) { free(b); return -1; }
    *n = k;
    *bounds = b;
    return 0;
}
static int payload_root_deneb(const uint8_t *p, uint64_t len, uint8_t out[32]) {
    if (len < 528) return -1;
    uint32_t o_extra = rd32(p + 436), o_tx = rd32(p + 504), o_wd = rd32(p + 508);
    if (o_extra != 528 || o_tx < o_extra || o_tx - o_extra > 32 || o_wd < o_tx || o_wd > len || (len - o_wd) % 44 ||
        (len - o_wd) / 44 > 16)
        return -1;
    uint8_t l[17][32];
    memset(l, 0, sizeof l);
    memcpy(l[0], p, 32);
    memcpy(l[1], p + 32, 20);
    memcpy(l[2], p + 52, 32);
    memcpy(l[3], p + 84, 32);
    orc_merkleize(p + 116, 8, 3, l[4]);
    memcpy(l[5], p + 372, 32);
    memcpy(l[6], p + 404, 8);
    memcpy(l[7], p + 412, 8);
    memcpy(l[8], p + 420, 8);
    memcpy(l[9], p + 428, 8);
    list_bytes_root(p + o_extra, o_tx - o_extra, 0, o_tx - o_extra, l[10]);
    memcpy(l[11], p + 440, 32);
    memcpy(l[12], p + 472, 32);
    { /* transactions: List[ByteList[2^30], 2^20] */
        uint64_t n, *b;
        if (var_list_bounds(p + o_tx, o_wd - o_tx, 1u << 20, &n, &b)) return -1;
        uint8_t *roots = (uint8_t *)malloc((size_t)(n ? n : 1) * 32);
        for (uint64_t i = 0; i < n; i++)
            list_bytes_root(p + o_tx + b[i], b[i + 1] - b[i], 25, b[i + 1] - b[i], roots + 32 * i);
        uint8_t r[32];
        orc_merkleize(roots, n, 20, r);
        orc_mix_in_length(r, n, l[13]);
        free(roots);
        free(b);
    }
    { /* withdrawals: List[Withdrawal, 16] */
        uint64_t n = (len - o_wd) / 44;
        uint8_t roots[16][32], r[32];
        for (uint64_t i = 0; i < n; i++) withdrawal_root(p + o_wd + 44 * i, roots[i]);
        orc_merkleize(&roots[0][0], n, 4, r);
        orc_mix_in_length(r, n, l[14]);
    }
    memcpy(l[15], p + 512, 8);
    memcpy(l[16], p + 520, 8);
    container_root(l, 17, out);
    return 0;
}
/* fixed-size item list: n = len / item, depth = log2(limit) */
static int fixed_list_root(const uint8_t *p, uint64_t len, uint32_t item, uint32_t limit_log, void (*f)(const uint8_t *, uint8_t *),
                           uint8_t out[32]) {
    if (len % item || len / item > (1ull << limit_log)) return -1;
    uint64_t n = len / item;
    uint8_t *roots = (uint8_t *)malloc((size_t)(n ? n : 1) * 32), r[32];
    for (uint64_t i = 0; i < n; i++) f(p + (uint64_t)item * i, roots + 32 * i);
    orc_merkleize(roots, n, limit_log, r);
    orc_mix_in_length(r, n, out);
    free(roots);
    return 0;
}
static void proposer_slashing_root(const uint8_t *p, uint8_t out[32]) { /* 416 B */
    uint8_t l[2][32];
    signed_header_root(p, l[0]);
    signed_header_root(p + 208, l[1]);
    container_root(l, 2, out);
}
static void kzg_commitment_root(const uint8_t *p, uint8_t out[32]) { pubkey_root(p, out); } /* 48 B blob (kzg_commitment.rs:51) */

/* blinded != 0: BlindedBeaconBlockBodyDeneb — field 9 is an ExecutionPayloadHeaderDeneb (beacon_block.rs:80,
 * payload.rs BlindedPayload) instead of the full payload; the body layout is otherwise identical. */
static int body_root_deneb(const uint8_t *p, uint64_t len, int blinded, uint8_t out[32]);
EXPORT int orc_beacon_block_body_root_deneb(const uint8_t *p, uint64_t len, uint8_t out[32]) {
    return body_root_deneb(p, len, 0, out);
}
static int body_root_deneb(const uint8_t *p, uint64_t len, int blinded, uint8_t out[32]) {
    ensure_backend();
    if (len < 392) return -1;
    uint32_t o_ps = rd32(p + 200), o_as = rd32(p + 204), o_at = rd32(p + 208), o_dp = rd32(p + 212), o_ex = rd32(p + 216),
             o_ep = rd32(p + 380), o_bc = rd32(p + 384), o_kz = rd32(p + 388);
    if (o_ps != 392 || o_as < o_ps || o_at < o_as || o_dp < o_at || o_ex < o_dp || o_ep < o_ex || o_bc < o_ep || o_kz < o_bc ||
        o_kz > len)
        return -1;
    uint8_t l[12][32];
    bytes_root(p, 96, 2, l[0]);
    eth1_data_root(p + 96, l[1]);
    memcpy(l[2], p + 168, 32);
    if (fixed_list_root(p + o_ps, o_as - o_ps, 416, 4, proposer_slashing_root, l[3])) return -1;
    { /* attester_slashings: List[AttesterSlashing, 2] (variable-size items) */
        uint64_t n, *b;
        if (var_list_bounds(p + o_as, o_at - o_as, 2, &n, &b)) return -1;
        uint8_t roots[2][32], r[32];
        for (uint64_t i = 0; i < n; i++) {
            const uint8_t *q = p + o_as + b[i];
            uint64_t ql = b[i + 1] - b[i];
            if (ql < 8) { free(b); return -1; }
            uint32_t a1 = rd32(q), a2 = rd32(q + 4);
            uint8_t pr[2][32];
            if (a1 != 8 || a2 < a1 || a2 > ql || indexed_attestation_root(q + a1, a2 - a1, pr[0]) ||
                indexed_attestation_root(q + a2, ql - a2, pr[1])) { free(b); return -1; }
            container_root(pr, 2, roots[i]);
        }
        orc_merkleize(&roots[0][0], n, 1, r);
        orc_mix_in_length(r, n, l[4]);
        free(b);
    }
    { /* attestations: List[Attestation, 128] */
        uint64_t n, *b;
        if (var_list_bounds(p + o_at, o_dp - o_at, 128, &n, &b)) return -1;
        uint8_t roots[128][32], r[32];
        for (uint64_t i = 0; i < n; i++)
            if (attestation_root(p + o_at + b[i], b[i + 1] - b[i], roots[i])) { free(b); return -1; }
        orc_merkleize(&roots[0][0], n, 7, r);
        orc_mix_in_length(r, n, l[5]);
        free(b);
    }
    if (fixed_list_root(p + o_dp, o_ex - o_dp, 1240, 4, deposit_root, l[6])) return -1;
    if (fixed_list_root(p + o_ex, o_ep - o_ex, 112, 4, exit_root, l[7])) return -1;
    { /* sync_aggregate: {Bitvector[512], signature} */
        uint8_t s[2][32];
        orc_merkleize(p + 220, 2, 1, s[0]);
        bytes_root(p + 284, 96, 2, s[1]);
        container_root(s, 2, l[8]);
    }
    if (blinded ? exec_header_root(p + o_ep, o_bc - o_ep, l[9]) : payload_root_deneb(p + o_ep, o_bc - o_ep, l[9])) return -1;
    if (fixed_list_root(p + o_bc, o_kz - o_bc, 172, 4, bls_change_root, l[10])) return -1;
    if (fixed_list_root(p + o_kz, len - o_kz, 48, 12, kzg_commitment_root, l[11])) return -1;
    container_root(l, 12, out);
    return 0;
}

/* BeaconBlock::canonical_root (beacon_block.rs:158-160).  body_root (32 B) optional. */
static int block_root_deneb(const uint8_t *p, uint64_t len, int blinded, uint8_t out[32], uint8_t *body_root);
EXPORT int orc_beacon_block_root_deneb(const uint8_t *p, uint64_t len, uint8_t out[32], uint8_t *body_root) {
    return block_root_deneb(p, len, 0, out, body_root);
}
/* BlindedBeaconBlock::canonical_root: equals the root of the full block it was blinded from. */
EXPORT int orc_blinded_beacon_block_root_deneb(const uint8_t *p, uint64_t len, uint8_t out[32], uint8_t *body_root) {
    return block_root_deneb(p, len, 1, out, body_root);
}
static int block_root_deneb(const uint8_t *p, uint64_t len, int blinded, uint8_t out[32], uint8_t *body_root) {
    if (len < 84 || rd32(p + 80) != 84) return -1;
    uint8_t l[5][32];
    u64_chunk(p, l[0]);
    u64_chunk(p + 8, l[1]);
    memcpy(l[2], p + 16, 32);
    memcpy(l[3], p + 48, 32);
    if (body_root_deneb(p + 84, len - 84, blinded, l[4])) return -1;
    if (body_root) memcpy(body_root, l[4], 32);
    container_root(l, 5, out);
    return 0;
}
