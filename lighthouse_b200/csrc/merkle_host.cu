// merkle_host.cu — host driver + C ABI of the tree-hash path.
//
// Host code only plans (offset parsing, literal chunk packing, launch tables); every SHA-256 compression
// runs on the device.  Mirrors, at the batch level, tree_hash::{merkle_root, mix_in_length, MerkleHasher},
// merkle_proof::MerkleTree and BeaconState::update_tree_hash_cache (cold) — see include/lhb200.h.
#include <string.h>
#include <algorithm>
#include <array>
#include <memory>
#include <vector>
#include "ctx.h"
#include "merkle.cuh"

namespace lhb200 {

static inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline uint32_t ceil_log2(uint64_t x) {
    uint32_t d = 0;
    while ((1ull << d) < x) d++;
    return d;
}

int32_t merkle_init() {
    Ctx& c = ctx();
    k_init_zero_hashes<<<1, 32, 0, c.stream>>>();
    count_launch();
    LHB_CUDA(cudaGetLastError());
    uint32_t words[MAX_ZERO_DEPTH + 1][8];
    LHB_CUDA(cudaStreamSynchronize(c.stream));
    LHB_CUDA(cudaMemcpyFromSymbol(words, g_zero_words, sizeof words));
    for (int d = 0; d <= MAX_ZERO_DEPTH; d++)
        for (int i = 0; i < 8; i++) {
            c.zero_hashes[d][4 * i + 0] = words[d][i] >> 24;
            c.zero_hashes[d][4 * i + 1] = words[d][i] >> 16;
            c.zero_hashes[d][4 * i + 2] = words[d][i] >> 8;
            c.zero_hashes[d][4 * i + 3] = words[d][i];
        }
    return LHB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Plan: everything needed to hash one object — leaf kernels, reduce passes, the tail hash program.
// All device memory a plan touches lives in one arena: [staged inputs | scratch nodes | literals | slots].
struct LeafLaunch {
    int kind;  // 0 validator roots, 1 pubkey roots, 2 eth1data roots, 3 hash pairs
    const uint8_t* in;
    uint8_t* out;
    uint64_t n;
};

struct Plan {
    uint8_t* arena = nullptr;   // device base
    size_t arena_bytes = 0;
    size_t bump = 0;            // allocation cursor (bytes)
    std::vector<uint8_t> lit;   // host literal block (32-byte chunks), copied to arena+lit_off
    size_t lit_off = 0, lit_cap = 0;
    std::vector<LeafLaunch> leaves;
    std::vector<std::vector<MerkleSeg>> passes;
    std::vector<HashOp> ops;
    std::vector<int> op_wave;
    // op outputs come from a dense node pool right after the literals, so "which wave produces this operand" is an
    // array lookup (a slot table over the whole 72 MB state arena cost ~2 ms per plan build)
    size_t node_off = 0, node_cap = 0, node_used = 0;
    bool node_overflow = false;
    std::vector<int16_t> slot_wave;                    // pool slot -> wave producing it
    uint64_t forced_base = 0;                          // forced destinations (block batch outputs): one dense range
    std::vector<int16_t> forced_wave;
    std::vector<ByteItem> items;                       // packed byte strings hashed straight from the staged blob
    ByteItem* d_items = nullptr;
    std::vector<int32_t> h_waves;                      // host copy of the wave table (wide waves get their own launch)
    uint64_t forced_dst = 0;                           // destination of the next op_hash (0 = allocate)
    // lists big enough for reduce passes: where their leaf chunks live and where the tail program reads their data
    // root (for the warm path, lhb200_state_enable_incremental)
    struct TreeSpec { const uint8_t* chunks; uint64_t n_chunks; uint64_t top_addr; const uint8_t* src; int kind;
                      uint64_t src_off, src_bytes; uint32_t item_bytes; };
    std::vector<TreeSpec> trees;
    uint64_t hash_units = 0;
    // SSZ provenance of literal chunks (for lhb200_state_patch): chunk index <- n bytes at SSZ offset src_off
    struct LitSrc { uint32_t lit_index; uint32_t n; uint64_t src_off; };
    std::vector<LitSrc> lit_src;
    const uint8_t* ssz_base = nullptr;
    uint64_t ssz_len = 0;
    // tail program (device copies)
    HashOp* d_ops = nullptr;
    int32_t* d_waves = nullptr;
    int n_waves = 0;
    uint64_t root_addr = 0;

    uint8_t* alloc(size_t nbytes, size_t align = 256) {  // device sub-allocation
        size_t off = align_up(bump, align);
        bump = off + nbytes;
        return arena ? arena + off : reinterpret_cast<uint8_t*>(off);
    }
    static uint64_t zero_op(uint32_t level) { return OP_ZERO_FLAG | level; }
    uint64_t literal_raw(const uint8_t chunk[32]) {
        size_t i = lit.size();
        lit.resize(i + 32);
        memcpy(&lit[i], chunk, 32);
        return reinterpret_cast<uint64_t>(arena + lit_off + i);
    }
    uint64_t literal_bytes(const uint8_t* p, size_t n) {  // zero-padded chunk from <=32 bytes
        uint8_t c[32] = {0};
        memcpy(c, p, n);
        if (ssz_base && p >= ssz_base && p + n <= ssz_base + ssz_len)
            lit_src.push_back({(uint32_t)(lit.size() / 32), (uint32_t)n, (uint64_t)(p - ssz_base)});
        return literal_raw(c);
    }
    uint64_t literal(const uint8_t chunk[32]) { return literal_bytes(chunk, 32); }
    uint64_t literal_u64(uint64_t v) {
        uint8_t c[32] = {0};
        for (int k = 0; k < 8; k++) c[k] = (uint8_t)(v >> (8 * k));
        return literal_raw(c);
    }
    int wave_of(uint64_t operand) const {
        if (operand & OP_ZERO_FLAG) return -1;
        const size_t slot = (operand - reinterpret_cast<uint64_t>(arena) - node_off) / 32;  // outside the pool: huge
        if (slot < slot_wave.size()) return slot_wave[slot];
        const size_t fs = (operand - forced_base) / 32;
        if (fs < forced_wave.size()) return forced_wave[fs];
        return -1;  // staged data / literals / leaf outputs: ready at start
    }
    uint64_t op_hash(uint64_t a, uint64_t b) {
        const int w = std::max(wave_of(a), wave_of(b)) + 1;
        uint64_t dst;
        if (forced_dst) {
            dst = forced_dst;
            forced_dst = 0;
            const size_t fs = (dst - forced_base) / 32;
            if (fs < forced_wave.size()) forced_wave[fs] = (int16_t)w;
        } else {
            if (node_used >= node_cap) { node_overflow = true; node_used = 0; }  // reported by build_plan
            dst = reinterpret_cast<uint64_t>(arena) + node_off + 32 * node_used;
            if (node_used >= slot_wave.size()) slot_wave.resize(std::max<size_t>(node_used + 1, 2 * slot_wave.size()), -1);
            slot_wave[node_used++] = (int16_t)w;
        }
        ops.push_back({dst, a, b});
        op_wave.push_back(w);
        hash_units++;
        return dst;
    }
    uint64_t mix_in_length(uint64_t root, uint64_t len) { return op_hash(root, literal_u64(len)); }
    // merkleize k operands over next_pow2(k) leaves (container / small vectors), padding with zero hashes
    // `final_dst` (optional): device address that receives the root (depth >= 1)
    uint64_t small_tree(std::vector<uint64_t> nodes, uint32_t depth, uint64_t final_dst = 0) {
        if (nodes.empty()) return zero_op(depth);
        for (uint32_t l = 0; l < depth; l++) {
            std::vector<uint64_t> nx;
            for (size_t i = 0; i < nodes.size(); i += 2) {
                if (l + 1 == depth) forced_dst = final_dst;
                nx.push_back(op_hash(nodes[i], i + 1 < nodes.size() ? nodes[i + 1] : zero_op(l)));
            }
            nodes.swap(nx);
        }
        return nodes[0];
    }
    uint64_t container(const std::vector<uint64_t>& fields, uint64_t final_dst = 0) {
        return small_tree(fields, ceil_log2(fields.size()), final_dst);
    }
    // hash_tree_root of a packed byte string resident at d_src (any alignment): merkleize(pack(bytes), 2^depth),
    // optionally mixed in with `length`; `last_mask` is ANDed onto the last byte (bitlist delimiter removal)
    uint64_t bytes_item(const uint8_t* d_src, uint64_t nbytes, uint32_t depth, bool mix, uint64_t length = 0,
                        uint32_t last_mask = 0xff) {
        uint8_t* out = alloc(32, 32);
        ByteItem it;
        it.src = d_src; it.nbytes = nbytes; it.out = out; it.length = length; it.depth = depth;
        it.flags = (mix ? 1u : 0u) | (last_mask << 8);
        items.push_back(it);
        uint64_t n = (nbytes + 31) / 32;
        for (uint32_t l = 1; l <= depth && n > 1; l++) { n = (n + 1) / 2; hash_units += n; }
        hash_units += mix ? 1 : 0;
        return reinterpret_cast<uint64_t>(out);
    }
    // merkleize n chunks resident at d_in (16-B aligned) with limit 2^depth
    uint64_t merkle_list(const uint8_t* d_in, uint64_t n, uint32_t depth) {
        if (n == 0) return zero_op(depth);
        if (n <= 8) {
            std::vector<uint64_t> nodes;
            for (uint64_t i = 0; i < n; i++) nodes.push_back(reinterpret_cast<uint64_t>(d_in + 32 * i));
            uint32_t d0 = std::min(depth, ceil_log2(n));
            uint64_t r = small_tree(nodes, d0);
            for (uint32_t l = d0; l < depth; l++) r = op_hash(r, zero_op(l));
            return r;
        }
        uint32_t level = 0;
        const uint8_t* in = d_in;
        size_t p = 0;
        const uint64_t hash_units_n0 = n;  // leaf count of this list
        while (n > 1 && level < depth) {
            uint32_t tl = std::min<uint32_t>(MAX_TILE_LOG, depth - level);
            tl = std::min<uint32_t>(tl, ceil_log2(n));  // do not fold past the single-root level here
            uint64_t n_out = ceil_div(n, 1ull << tl);
            uint8_t* out = alloc(n_out * 32);
            if (passes.size() <= p) passes.resize(p + 1);
            MerkleSeg sg;
            sg.in = in; sg.out = out; sg.n_in = n; sg.level_in = level; sg.tile_log = tl;
            sg.cta_begin = 0; sg.n_tiles = (uint32_t)n_out;
            passes[p++].push_back(sg);
            // hashes actually performed: nodes with a valid left child at each level
            for (uint32_t l = 1; l <= tl; l++) hash_units += ceil_div(n, 1ull << l);
            level += tl;
            n = n_out;
            in = out;
        }
        uint64_t r = reinterpret_cast<uint64_t>(in);
        trees.push_back({d_in, hash_units_n0, r, nullptr, 1, ~0ull, 0, 32});
        for (uint32_t l = level; l < depth; l++) r = op_hash(r, zero_op(l));
        return r;
    }
    uint8_t* leaf_kernel(int kind, const uint8_t* in, uint64_t n) {
        uint8_t* out = alloc(std::max<uint64_t>(n, 1) * 32);
        if (n) leaves.push_back({kind, in, out, n});
        static const int units[6] = {8, 1, 3, 1, 1, 3};
        hash_units += (uint64_t)units[kind] * n;
        return out;
    }
};

static int32_t plan_enqueue_tail(Plan& pl, cudaStream_t s);
// Enqueue a finished plan on `s`.  Literals must already be in the arena.
static int32_t plan_enqueue(Plan& pl, cudaStream_t s, cudaEvent_t e0 = nullptr, cudaEvent_t e1 = nullptr) {
    for (const LeafLaunch& L : pl.leaves) {
        switch (L.kind) {
            case 0:
                if (e0) cudaEventRecord(e0, s);
                k_validator_roots<<<(unsigned)ceil_div(L.n, VAL_PER_CTA), VAL_PER_CTA, 0, s>>>(L.in, L.n, L.out);
                if (e1) cudaEventRecord(e1, s);
                break;
            case 1:
                k_record_roots<<<(unsigned)ceil_div(L.n, 128), 128, 0, s>>>(L.in, L.n, 0, L.out);
                break;
            case 2:
                k_record_roots<<<(unsigned)ceil_div(L.n, 128), 128, 0, s>>>(L.in, L.n, 1, L.out);
                break;
            case 4:
                k_record_roots<<<(unsigned)ceil_div(L.n, 128), 128, 0, s>>>(L.in, L.n, 2, L.out);
                break;
            case 5:
                k_record_roots<<<(unsigned)ceil_div(L.n, 128), 128, 0, s>>>(L.in, L.n, 3, L.out);
                break;
            default:
                k_hash_pairs<<<(unsigned)ceil_div(L.n, 256), 256, 0, s>>>(L.in, L.out, L.n);
        }
        count_launch();
    }
    for (auto& pass : pl.passes) {
        for (size_t i = 0; i < pass.size(); i += MAX_SEGS) {
            MerkleSegTable tab;
            tab.n = (int)std::min<size_t>(MAX_SEGS, pass.size() - i);
            uint32_t ctas = 0;
            for (int k = 0; k < tab.n; k++) {
                tab.s[k] = pass[i + k];
                tab.s[k].cta_begin = ctas;
                ctas += tab.s[k].n_tiles;
            }
            k_merkle_reduce<<<ctas, REDUCE_THREADS, 0, s>>>(tab);
            count_launch();
        }
    }
    if (!pl.items.empty()) {
        k_byte_items<<<(unsigned)pl.items.size(), ITEM_THREADS, 0, s>>>(pl.d_items);
        count_launch();
    }
    return plan_enqueue_tail(pl, s);
}
// The tail hash program only (zero ladders, length mix-ins, containers): all a warm root needs after the trees.
static int32_t plan_enqueue_tail(Plan& pl, cudaStream_t s) {
    // narrow waves run back to back inside one CTA; a wide wave (block batches) gets the whole grid
    constexpr int WIDE_WAVE = 1024;
    for (int w = 0; w < pl.n_waves;) {
        const int cnt = pl.h_waves[w + 1] - pl.h_waves[w];
        if (cnt >= WIDE_WAVE) {
            k_hash_ops<<<(unsigned)ceil_div(cnt, PROG_THREADS), PROG_THREADS, 0, s>>>(pl.d_ops + pl.h_waves[w], cnt);
            w++;
        } else {
            int e = w;
            while (e < pl.n_waves && pl.h_waves[e + 1] - pl.h_waves[e] < WIDE_WAVE) e++;
            k_hash_program<<<1, PROG_THREADS, 0, s>>>(pl.d_ops, pl.d_waves + w, e - w);
            w = e;
        }
        count_launch();
    }
    LHB_CUDA(cudaGetLastError());
    return LHB200_OK;
}

// Sort ops by wave, build wave table; returns host blobs to upload.
static void plan_finalize_program(Plan& pl, std::vector<HashOp>& ops_sorted, std::vector<int32_t>& waves) {
    int nw = 0;
    for (int w : pl.op_wave) nw = std::max(nw, w + 1);
    std::vector<std::vector<HashOp>> by(nw);
    for (size_t i = 0; i < pl.ops.size(); i++) by[pl.op_wave[i]].push_back(pl.ops[i]);
    waves.assign(1, 0);
    for (int w = 0; w < nw; w++) {
        for (auto& o : by[w]) ops_sorted.push_back(o);
        waves.push_back((int32_t)ops_sorted.size());
    }
    pl.n_waves = nw;
    pl.h_waves = waves;
}

// ---------------------------------------------------------------------------------------------------------
// A "session": dry-run the planner to size the arena, allocate, then plan for real.  `build` must be
// deterministic in its allocation sequence.
template <class F>
static int32_t build_plan(Plan& pl, uint8_t* arena, size_t arena_bytes, size_t lit_cap, F&& build,
                          size_t node_cap = 8192) {
    pl = Plan();
    pl.arena = arena;
    pl.arena_bytes = arena_bytes;
    pl.lit_cap = lit_cap;
    pl.lit_off = 0;
    pl.node_off = align_up(lit_cap, 256);   // literals first, then the node pool
    pl.node_cap = node_cap;
    pl.bump = pl.node_off + node_cap * 32;
    build(pl);
    if (pl.lit.size() > lit_cap) {
        set_error("internal: literal block overflow (%zu > %zu)", pl.lit.size(), lit_cap);
        return LHB200_EINVAL;
    }
    if (pl.node_overflow) {
        set_error("internal: hash program larger than its node pool (%zu nodes)", node_cap);
        return LHB200_EINVAL;
    }
    return LHB200_OK;
}

// upload literals + program; program blobs live at the end of the arena
static int32_t plan_upload(Plan& pl, cudaStream_t s, std::vector<HashOp>& ops_sorted, std::vector<int32_t>& waves,
                           void* h_stage) {
    // h_stage: pinned host staging of at least lit + ops + waves bytes
    uint8_t* h = static_cast<uint8_t*>(h_stage);
    size_t o = 0;
    if (!pl.lit.empty()) {
        memcpy(h + o, pl.lit.data(), pl.lit.size());
        LHB_CUDA(cudaMemcpyAsync(pl.arena + pl.lit_off, h + o, pl.lit.size(), cudaMemcpyHostToDevice, s));
        o += align_up(pl.lit.size(), 256);
    }
    if (!ops_sorted.empty()) {
        size_t nb = ops_sorted.size() * sizeof(HashOp);
        pl.d_ops = reinterpret_cast<HashOp*>(pl.alloc(nb));
        memcpy(h + o, ops_sorted.data(), nb);
        LHB_CUDA(cudaMemcpyAsync(pl.d_ops, h + o, nb, cudaMemcpyHostToDevice, s));
        o += align_up(nb, 256);
        size_t wb = waves.size() * sizeof(int32_t);
        pl.d_waves = reinterpret_cast<int32_t*>(pl.alloc(wb));
        memcpy(h + o, waves.data(), wb);
        LHB_CUDA(cudaMemcpyAsync(pl.d_waves, h + o, wb, cudaMemcpyHostToDevice, s));
        o += align_up(wb, 256);
    }
    if (!pl.items.empty()) {
        size_t ib = pl.items.size() * sizeof(ByteItem);
        pl.d_items = reinterpret_cast<ByteItem*>(pl.alloc(ib));
        memcpy(h + o, pl.items.data(), ib);
        LHB_CUDA(cudaMemcpyAsync(pl.d_items, h + o, ib, cudaMemcpyHostToDevice, s));
    }
    return LHB200_OK;
}

// Generic runner for "one input blob -> one root" host entry points.
//   stage(plan, d_in) describes the work given the device copy of the input.
template <class F>
static int32_t run_simple(const uint8_t* h_in, size_t in_bytes, uint8_t out[32], F&& describe, bool in_on_device = false) {
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    // pass 1: dry run to size
    Plan dry;
    const size_t lit_cap = 4096;
    size_t in_pad = align_up(in_bytes + 32, 256);
    uint64_t root_dry = 0;
    build_plan(dry, nullptr, 0, lit_cap, [&](Plan& p) {
        uint8_t* d_in = p.alloc(in_pad);
        root_dry = describe(p, d_in);
    });
    size_t prog_bytes = align_up(dry.ops.size() * sizeof(HashOp), 256) + align_up((dry.ops.size() + 2) * 4, 256) + 512;
    size_t need = align_up(dry.bump, 256) + prog_bytes + 256;
    uint8_t* arena = static_cast<uint8_t*>(dev_scratch(need));
    if (!arena) return LHB200_ENOMEM;
    size_t stage_bytes = in_bytes + lit_cap + prog_bytes + 1024;
    uint8_t* hst = static_cast<uint8_t*>(pinned_scratch(stage_bytes));
    if (!hst) return LHB200_ENOMEM;
    Plan pl;
    uint64_t root = 0;
    uint8_t* d_in = nullptr;
    int32_t rc = build_plan(pl, arena, need, lit_cap, [&](Plan& p) {
        d_in = p.alloc(in_pad);
        root = describe(p, d_in);
    });
    if (rc) return rc;
    // H2D input (zero the padding tail so packed lists see zero-filled last chunks)
    LHB_CUDA(cudaMemsetAsync(d_in + in_bytes / 256 * 256, 0, in_pad - in_bytes / 256 * 256, c.stream));
    if (in_bytes && in_on_device) {
        LHB_CUDA(cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyDeviceToDevice, c.stream));
    } else if (in_bytes) {
        cudaPointerAttributes at;
        bool pinned = cudaPointerGetAttributes(&at, h_in) == cudaSuccess && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
        const void* src = h_in;
        if (!pinned) {
            memcpy(hst, h_in, in_bytes);
            src = hst;
        }
        LHB_CUDA(cudaMemcpyAsync(d_in, src, in_bytes, cudaMemcpyHostToDevice, c.stream));
    }
    std::vector<HashOp> ops_sorted;
    std::vector<int32_t> waves;
    plan_finalize_program(pl, ops_sorted, waves);
    rc = plan_upload(pl, c.stream, ops_sorted, waves, hst + align_up(in_bytes, 256));
    if (rc) return rc;
    rc = plan_enqueue(pl, c.stream);
    if (rc) return rc;
    if (root & OP_ZERO_FLAG) {
        LHB_CUDA(cudaStreamSynchronize(c.stream));
        memcpy(out, c.zero_hashes[root & 0xff], 32);
        return LHB200_OK;
    }
    uint8_t* h_out = hst + stage_bytes - 64;
    LHB_CUDA(cudaMemcpyAsync(h_out, reinterpret_cast<void*>(root), 32, cudaMemcpyDeviceToHost, c.stream));
    LHB_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(out, h_out, 32);
    return LHB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Deneb BeaconState (mainnet preset).  Fixed-part offsets: see DESIGN.md §3 / oracle for the derivation.
namespace deneb {
constexpr uint32_t O_GENESIS_TIME = 0, O_GVR = 8, O_SLOT = 40, O_FORK = 48, O_LBH = 64, O_BLOCK_ROOTS = 176,
                   O_STATE_ROOTS = 262320, O_HIST_OFF = 524464, O_ETH1_DATA = 524468, O_VOTES_OFF = 524540,
                   O_DEPOSIT_INDEX = 524544, O_VAL_OFF = 524552, O_BAL_OFF = 524556, O_RANDAO = 524560,
                   O_SLASHINGS = 2621712, O_PP_OFF = 2687248, O_CP_OFF = 2687252, O_JUST = 2687256,
                   O_PJC = 2687257, O_CJC = 2687297, O_FC = 2687337, O_INACT_OFF = 2687377, O_CSC = 2687381,
                   O_NSC = 2712005, O_LEPH_OFF = 2736629, O_NWI = 2736633, O_NWVI = 2736641, O_HS_OFF = 2736649,
                   O_ELECTRA_U64 = 2736653, O_PBD_OFF = 2736701, O_PPW_OFF = 2736705, O_PC_OFF = 2736709;
constexpr uint32_t SYNC_COMMITTEE_BYTES = 513 * 48;
}  // namespace deneb

static inline uint32_t rd32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

}  // namespace lhb200

namespace lhb200 {
// Multi-GPU sharding of one state: rank r of `world` (a power of two) owns the leaf range
// [r * 2^s, (r+1) * 2^s) of every big list (s = ceil_log2(#chunks) - log2(world)) and produces the 32-byte root of
// that height-s subtree; small fields are computed by every rank.  After one all-gather of the subtree roots each
// rank folds them (log2(world) levels + zero ladder + length mix-in + the 32-leaf container) in lhb200_state_combine.
struct StageCopy {
    size_t src_off, nbytes;
    uint8_t* dst;
    size_t pad_to;  // zero-fill up to this many bytes at dst
};

struct ShardCfg {
    uint32_t rank = 0, world = 1;
    int32_t fork = LHB200_FORK_DENEB;   // which BeaconState variant the SSZ is (fork_spec)
};
struct ShardedList {
    int field;            // index in the state container
    uint32_t s;           // height of the per-rank subtree
    uint32_t limit_depth; // chunk-tree depth of the list limit
    uint64_t mix_len;     // length to mix in (lists); UINT64_MAX = vector (no mix-in)
    uint64_t local_op;    // operand holding this rank's subtree root
};

}  // namespace lhb200

static uint8_t* g_spare_arena = nullptr;  // guarded by ctx().mu
static size_t g_spare_bytes = 0;
// lhb200_shutdown: the recycled arena belongs to the context that is going away
namespace lhb200 {
void merkle_shutdown();
}
void lhb200::merkle_shutdown() {
    if (g_spare_arena) cudaFree(g_spare_arena);
    g_spare_arena = nullptr;
    g_spare_bytes = 0;
}

struct lhb200_state {
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0;
    lhb200::Plan plan;
    uint64_t field_ops[40];   // MAX_FIELDS (declared below) rounded up
    uint64_t root_op = 0;
    uint8_t* d_result = nullptr;  // (1 + MAX_FIELDS) * 32 bytes: root + field roots gathered
    cudaEvent_t e_k0 = nullptr, e_k1 = nullptr;  // around k_validator_roots
    lhb200::ShardCfg shard;
    std::vector<lhb200::ShardedList> sharded;
    uint8_t* d_coll = nullptr;              // lhb200_state_root_sharded: gather ops | own subtree roots | all ranks' roots
    std::vector<lhb200::StageCopy> copies;  // SSZ ranges resident in the arena (for lhb200_state_patch)
    std::vector<uint32_t> copy_order, lit_order;   // offset-sorted indices into copies / plan.lit_src (patch lookups)
    std::vector<int32_t> copy_tree;         // copies[k] -> index of its resident tree (warm path), -1 if none
    size_t copy_tree_trees = 0;
    // warm path (lhb200_state_enable_incremental): full level arrays per big list + dirty leaves since the last root
    // dirty leaves since the last root: a host BITMAP per tree (marking is O(1) per edit and dedups for free; the next
    // root extracts the sorted index list with a ctz scan — no sort) plus the number of marks made
    struct Tree {
        lhb200::TreeDev dev; uint64_t src_off, src_bytes; uint32_t item_bytes;
        std::vector<uint64_t> dirty_bits; std::vector<uint32_t> dirty; uint64_t n_marks = 0;
        void mark(uint64_t leaf) { dirty_bits[leaf >> 6] |= 1ull << (leaf & 63); n_marks++; }
    };
    bool incremental = false, need_full = false;
    std::vector<Tree> trees;
    uint8_t* d_levels = nullptr;
    lhb200::TreeDev* d_trees = nullptr;
    uint32_t* d_dirty = nullptr;
    uint64_t last_root_hashes = 0;       // hash32_concat units of the last root (full or incremental)
    static constexpr uint32_t DIRTY_CAP = 1u << 16;
};

namespace lhb200 {

// Describe the whole Deneb state.  `s` = host SSZ (read for offsets and small literal fields only).
// Big fields are placed in the arena by `place(src_off, nbytes)` which records an H2D copy.
// The post-Altair BeaconState variants (consensus/types/src/beacon_state.rs:224-571) share the first 24 fields and
// their fixed-part offsets; later forks APPEND fields and widen the execution payload header:
//   Altair     24 fields                                                        fixed part 2 736 629 B
//   Bellatrix  + latest_execution_payload_header (14 fields, 536 B fixed)       2 736 633 B
//   Capella    header + withdrawals_root (15 fields, 568 B); + next_withdrawal_index, next_withdrawal_validator_index,
//              historical_summaries                                             2 736 653 B
//   Deneb      header + blob_gas_used, excess_blob_gas (17 fields, 584 B)       2 736 653 B
// so one describer covers them all; the kernels are fork-agnostic.
constexpr int MAX_FIELDS = 37;   // BeaconStateElectra; the field-root block of a handle is root + MAX_FIELDS chunks
struct ForkSpec {
    int n_fields;        // 24 / 25 / 28 / 28 / 37
    uint32_t fixed;      // bytes of the fixed part
    uint32_t hdr_fixed;  // fixed part of the execution payload header (0: no header)
    int hdr_fields;
};
static bool fork_spec(int32_t fork, ForkSpec* f) {
    switch (fork) {
        case LHB200_FORK_ALTAIR: *f = {24, 2736629, 0, 0}; return true;
        case LHB200_FORK_BELLATRIX: *f = {25, 2736633, 536, 14}; return true;
        case LHB200_FORK_CAPELLA: *f = {28, 2736653, 568, 15}; return true;
        case LHB200_FORK_DENEB: *f = {28, 2736653, 584, 17}; return true;
        case LHB200_FORK_ELECTRA: *f = {37, 2736713, 648, 19}; return true;
    }
    return false;
}

static int32_t describe_deneb(Plan& p, const uint8_t* s, uint64_t len, std::vector<StageCopy>* copies,
                              uint64_t* field_ops, uint64_t* root_op, ShardCfg sh = ShardCfg(),
                              std::vector<ShardedList>* sharded = nullptr) {
    using namespace deneb;
    ForkSpec fk;
    if (!fork_spec(sh.fork, &fk)) { set_error("unknown fork id %d", sh.fork); return LHB200_EINVAL; }
    const uint32_t FIXED = fk.fixed;
    if (len < FIXED) { set_error("BeaconState SSZ shorter than its fixed part"); return LHB200_EINVAL; }
    const uint32_t o_hist = rd32(s + O_HIST_OFF), o_votes = rd32(s + O_VOTES_OFF), o_val = rd32(s + O_VAL_OFF),
                   o_bal = rd32(s + O_BAL_OFF), o_pp = rd32(s + O_PP_OFF), o_cp = rd32(s + O_CP_OFF),
                   o_inact = rd32(s + O_INACT_OFF);
    const uint32_t o_leph = fk.hdr_fields ? rd32(s + O_LEPH_OFF) : (uint32_t)len;
    const uint32_t o_hs = fk.n_fields >= 28 ? rd32(s + O_HS_OFF) : (uint32_t)len;
    const bool electra = fk.n_fields == 37;
    const uint32_t o_pbd = electra ? rd32(s + O_PBD_OFF) : (uint32_t)len, o_ppw = electra ? rd32(s + O_PPW_OFF) : (uint32_t)len,
                   o_pc = electra ? rd32(s + O_PC_OFF) : (uint32_t)len;
    if (o_hist != FIXED || !(o_hist <= o_votes && o_votes <= o_val && o_val <= o_bal && o_bal <= o_pp &&
                             o_pp <= o_cp && o_cp <= o_inact && o_inact <= o_leph && o_leph <= o_hs && o_hs <= o_pbd &&
                             o_pbd <= o_ppw && o_ppw <= o_pc && o_pc <= len)) {
        set_error("BeaconState SSZ: inconsistent variable-part offsets");
        return LHB200_EINVAL;
    }
    const uint64_t n_hist = (o_votes - o_hist) / 32, n_votes = (o_val - o_votes) / 72, n_val = (o_bal - o_val) / 121,
                   n_bal = (o_pp - o_bal) / 8, n_pp = o_cp - o_pp, n_cp = o_inact - o_cp,
                   n_inact = (o_leph - o_inact) / 8, leph_len = o_hs - o_leph, n_hs = (o_pbd - o_hs) / 64,
                   n_pbd = (o_ppw - o_pbd) / 16, n_ppw = (o_pc - o_ppw) / 24, n_pc = (len - o_pc) / 16;
    if ((o_votes - o_hist) % 32 || (o_val - o_votes) % 72 || (o_bal - o_val) % 121 || (o_pp - o_bal) % 8 ||
        (o_leph - o_inact) % 8 || (o_pbd - o_hs) % 64 || (o_ppw - o_pbd) % 16 || (o_pc - o_ppw) % 24 || (len - o_pc) % 16 ||
        n_pbd > (1u << 27) || n_ppw > (1u << 27) || n_pc > (1u << 18) ||
        (fk.hdr_fields && (leph_len < fk.hdr_fixed || leph_len > fk.hdr_fixed + 32 || rd32(s + o_leph + 436) != fk.hdr_fixed)) ||
        n_votes > 2048 || n_hist > (1u << 24) || n_hs > (1u << 24)) {
        set_error("BeaconState SSZ: malformed variable part");
        return LHB200_EINVAL;
    }
    p.ssz_base = s;
    p.ssz_len = len;
    auto place = [&](size_t src_off, size_t nbytes) -> uint8_t* {
        size_t padded = align_up(nbytes + 32, 256);
        uint8_t* d = p.alloc(padded);
        if (copies) copies->push_back({src_off, nbytes, d, padded});
        return d;
    };
    auto chunk = [&](uint32_t off) { return p.literal(s + off); };
    uint64_t* f = field_ops;
    const uint32_t lg_world = ceil_log2(sh.world);
    // Big list with `n_chunks` leaf chunks produced from `n_items` source items of `item_bytes` at `src_off`
    // (leaf_kind < 0: the bytes already are the chunks).  Unsharded: the full field root.  Sharded: this rank's subtree.
    auto big_list = [&](int field, size_t src_off, uint64_t n_items, uint32_t item_bytes, int leaf_kind,
                        uint64_t n_chunks, uint32_t limit_depth, uint64_t mix_len) -> uint64_t {
        const uint32_t d0 = ceil_log2(std::max<uint64_t>(n_chunks, 1));
        const bool shard = sh.world > 1 && sharded && d0 >= lg_world + 6;
        if (!shard) {
            const uint8_t* src = place(src_off, n_items * item_bytes);
            const uint8_t* chunks = leaf_kind >= 0 ? p.leaf_kernel(leaf_kind, src, n_items) : src;
            const size_t nt = p.trees.size();
            uint64_t r = p.merkle_list(chunks, n_chunks, limit_depth);
            if (p.trees.size() > nt && leaf_kind <= 0) {   // warm-path provenance: validators (kind 0) or packed bytes
                Plan::TreeSpec& t = p.trees.back();
                t.src = src; t.kind = leaf_kind == 0 ? 0 : 1; t.src_off = src_off; t.src_bytes = n_items * item_bytes;
                t.item_bytes = leaf_kind == 0 ? item_bytes : 32;
            }
            return mix_len == UINT64_MAX ? r : p.mix_in_length(r, mix_len);
        }
        const uint32_t sub = d0 - lg_world;
        const uint64_t c_lo = std::min<uint64_t>(n_chunks, (uint64_t)sh.rank << sub);
        const uint64_t c_hi = std::min<uint64_t>(n_chunks, ((uint64_t)sh.rank + 1) << sub);
        const uint64_t cnt = c_hi - c_lo;
        uint64_t op;
        if (leaf_kind >= 0) {          // one chunk per source item
            const uint8_t* src = place(src_off + c_lo * item_bytes, cnt * item_bytes);
            op = p.merkle_list(p.leaf_kernel(leaf_kind, src, cnt), cnt, sub);
        } else {                       // packed bytes: chunk c covers bytes [32c, 32c+32) of the field
            const uint64_t total_bytes = n_items * item_bytes;
            const uint64_t b_lo = c_lo * 32, b_hi = std::min<uint64_t>(total_bytes, c_hi * 32);
            const uint8_t* src = place(src_off + b_lo, b_hi > b_lo ? b_hi - b_lo : 0);
            op = p.merkle_list(src, cnt, sub);
        }
        sharded->push_back({field, sub, limit_depth, mix_len, op});
        return Plan::zero_op(0);       // placeholder; the field root is formed in lhb200_state_combine
    };
    f[0] = p.literal_bytes(s + O_GENESIS_TIME, 8);
    f[1] = chunk(O_GVR);
    f[2] = p.literal_bytes(s + O_SLOT, 8);
    f[3] = p.container({p.literal_bytes(s + O_FORK, 4), p.literal_bytes(s + O_FORK + 4, 4),
                        p.literal_bytes(s + O_FORK + 8, 8)});
    f[4] = p.container({p.literal_bytes(s + O_LBH, 8), p.literal_bytes(s + O_LBH + 8, 8), chunk(O_LBH + 16),
                        chunk(O_LBH + 48), chunk(O_LBH + 80)});
    auto plain_vector = [&](uint32_t off, uint64_t nbytes, uint32_t depth) {   // fixed vectors of chunks / packed u64
        const uint8_t* src = place(off, nbytes);
        const size_t nt = p.trees.size();
        uint64_t r = p.merkle_list(src, nbytes / 32, depth);
        if (p.trees.size() > nt) {
            Plan::TreeSpec& t = p.trees.back();
            t.src = src; t.kind = 1; t.src_off = off; t.src_bytes = nbytes; t.item_bytes = 32;
        }
        return r;
    };
    f[5] = plain_vector(O_BLOCK_ROOTS, 8192 * 32, 13);
    f[6] = plain_vector(O_STATE_ROOTS, 8192 * 32, 13);
    f[7] = p.mix_in_length(p.merkle_list(place(o_hist, n_hist * 32), n_hist, 24), n_hist);
    f[8] = p.container({chunk(O_ETH1_DATA), p.literal_bytes(s + O_ETH1_DATA + 32, 8), chunk(O_ETH1_DATA + 40)});
    f[9] = p.mix_in_length(p.merkle_list(p.leaf_kernel(2, place(o_votes, n_votes * 72), n_votes), n_votes, 11), n_votes);
    f[10] = p.literal_bytes(s + O_DEPOSIT_INDEX, 8);
    f[11] = big_list(11, o_val, n_val, 121, 0, n_val, 40, n_val);
    f[12] = big_list(12, o_bal, n_bal, 8, -1, ceil_div(n_bal * 8, 32), 38, n_bal);
    f[13] = big_list(13, O_RANDAO, 65536, 32, -1, 65536, 16, UINT64_MAX);
    f[14] = plain_vector(O_SLASHINGS, 8192 * 8, 11);
    f[15] = big_list(15, o_pp, n_pp, 1, -1, ceil_div(n_pp, 32), 35, n_pp);
    f[16] = big_list(16, o_cp, n_cp, 1, -1, ceil_div(n_cp, 32), 35, n_cp);
    f[17] = p.literal_bytes(s + O_JUST, 1);
    f[18] = p.container({p.literal_bytes(s + O_PJC, 8), chunk(O_PJC + 8)});
    f[19] = p.container({p.literal_bytes(s + O_CJC, 8), chunk(O_CJC + 8)});
    f[20] = p.container({p.literal_bytes(s + O_FC, 8), chunk(O_FC + 8)});
    f[21] = big_list(21, o_inact, n_inact, 8, -1, ceil_div(n_inact * 8, 32), 38, n_inact);
    for (int k = 0; k < 2; k++) {
        uint8_t* roots = p.leaf_kernel(1, place(k ? O_NSC : O_CSC, SYNC_COMMITTEE_BYTES), 513);
        f[22 + k] = p.container({p.merkle_list(roots, 512, 9), reinterpret_cast<uint64_t>(roots + 512 * 32)});
    }
    if (fk.hdr_fields) {
        const uint8_t* h = s + o_leph;
        std::vector<uint64_t> bloom;
        for (int i = 0; i < 8; i++) bloom.push_back(p.literal(h + 116 + 32 * i));
        const uint64_t extra_len = leph_len - fk.hdr_fixed;
        std::vector<uint64_t> hf = {p.literal(h), p.literal_bytes(h + 32, 20), p.literal(h + 52), p.literal(h + 84),
                                    p.small_tree(bloom, 3), p.literal(h + 372), p.literal_bytes(h + 404, 8),
                                    p.literal_bytes(h + 412, 8), p.literal_bytes(h + 420, 8), p.literal_bytes(h + 428, 8),
                                    p.mix_in_length(p.literal_bytes(h + fk.hdr_fixed, extra_len), extra_len),
                                    p.literal(h + 440), p.literal(h + 472), p.literal(h + 504)};
        if (fk.hdr_fields >= 15) hf.push_back(p.literal(h + 536));                       // withdrawals_root (Capella)
        if (fk.hdr_fields >= 17) { hf.push_back(p.literal_bytes(h + 568, 8)); hf.push_back(p.literal_bytes(h + 576, 8)); }
        if (fk.hdr_fields >= 19) { hf.push_back(p.literal(h + 584)); hf.push_back(p.literal(h + 616)); }   // request roots (Electra)
        f[24] = p.container(hf);
    }
    if (fk.n_fields >= 28) {
        f[25] = p.literal_bytes(s + O_NWI, 8);
        f[26] = p.literal_bytes(s + O_NWVI, 8);
        f[27] = p.mix_in_length(p.merkle_list(p.leaf_kernel(3, place(o_hs, n_hs * 64), n_hs), n_hs, 24), n_hs);
    }
    if (electra) {   // beacon_state.rs:487-525
        for (int k = 0; k < 6; k++) f[28 + k] = p.literal_bytes(s + O_ELECTRA_U64 + 8 * k, 8);
        f[34] = p.mix_in_length(p.merkle_list(p.leaf_kernel(4, place(o_pbd, n_pbd * 16), n_pbd), n_pbd, 27), n_pbd);
        f[35] = p.mix_in_length(p.merkle_list(p.leaf_kernel(5, place(o_ppw, n_ppw * 24), n_ppw), n_ppw, 27), n_ppw);
        f[36] = p.mix_in_length(p.merkle_list(p.leaf_kernel(4, place(o_pc, n_pc * 16), n_pc), n_pc, 18), n_pc);
    }
    for (int k = fk.n_fields; k < MAX_FIELDS; k++) f[k] = Plan::zero_op(0);   // absent in this fork (not part of its container)
    *root_op = p.container(std::vector<uint64_t>(f, f + fk.n_fields));
    return LHB200_OK;
}

__global__ void k_gather_nodes(const HashOp* __restrict__ srcs, int n, uint8_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[8];
    load_operand(srcs[i].a, w);
    store_chunk(out + 32 * i, w);
}

}  // namespace lhb200

using namespace lhb200;

extern "C" {

int32_t lhb200_hash_pairs(const uint8_t* in, uint8_t* out, uint64_t n) {
    LHB_REQUIRE_READY();
    if (n == 0) return LHB200_OK;
    if (!in || !out) { set_error("null buffer"); return LHB200_EINVAL; }
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    uint8_t* d = static_cast<uint8_t*>(dev_scratch(n * 96 + 512));
    uint8_t* h = static_cast<uint8_t*>(pinned_scratch(n * 96));
    if (!d || !h) return LHB200_ENOMEM;
    memcpy(h, in, n * 64);
    LHB_CUDA(cudaMemcpyAsync(d, h, n * 64, cudaMemcpyHostToDevice, c.stream));
    uint8_t* d_out = d + align_up(n * 64, 256);
    k_hash_pairs<<<(unsigned)ceil_div(n, 256), 256, 0, c.stream>>>(d, d_out, n);
    count_launch();
    LHB_CUDA(cudaGetLastError());
    LHB_CUDA(cudaMemcpyAsync(h + n * 64, d_out, n * 32, cudaMemcpyDeviceToHost, c.stream));
    LHB_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(out, h + n * 64, n * 32);
    return LHB200_OK;
}

int32_t lhb200_dev_hash_pairs(const void* d_in, void* d_out, uint64_t n, void* stream) {
    LHB_REQUIRE_READY();
    if (n == 0) return LHB200_OK;
    if (!d_in || !d_out || (reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15)) {
        set_error("device buffers must be non-null and 16-byte aligned");
        return LHB200_EINVAL;
    }
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx().stream;
    k_hash_pairs<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(static_cast<const uint8_t*>(d_in),
                                                           static_cast<uint8_t*>(d_out), n);
    count_launch();
    LHB_CUDA(cudaGetLastError());
    return LHB200_OK;
}

int32_t lhb200_merkleize(const uint8_t* chunks, uint64_t n_chunks, uint32_t depth, uint8_t out[32]) {
    LHB_REQUIRE_READY();
    if (!out || (n_chunks && !chunks) || depth > 64 || (depth < 64 && n_chunks > (1ull << depth))) {
        set_error("merkleize: bad arguments (n_chunks must be <= 2^depth, depth <= 64)");
        return LHB200_EINVAL;
    }
    return run_simple(chunks, n_chunks * 32, out,
                      [&](Plan& p, uint8_t* d_in) { return p.merkle_list(d_in, n_chunks, depth); });
}

int32_t lhb200_dev_merkleize(const void* d_chunks, uint64_t n_chunks, uint32_t depth, void* d_out32, void* stream) {
    LHB_REQUIRE_READY();
    if (!d_out32 || (n_chunks && !d_chunks) || (reinterpret_cast<uintptr_t>(d_chunks) & 15) ||
        (reinterpret_cast<uintptr_t>(d_out32) & 15) || depth > 64 || (depth < 64 && n_chunks > (1ull << depth))) {
        set_error("dev_merkleize: bad arguments");
        return LHB200_EINVAL;
    }
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c.stream;
    Plan dry;
    uint64_t root = 0;
    auto describe = [&](Plan& p) { root = p.merkle_list(static_cast<const uint8_t*>(d_chunks), n_chunks, depth); };
    build_plan(dry, nullptr, 0, 256, describe);
    size_t prog = align_up(dry.ops.size() * sizeof(HashOp), 256) + align_up((dry.ops.size() + 2) * 4, 256) + 512;
    size_t need = align_up(dry.bump, 256) + prog + 256;
    uint8_t* arena = static_cast<uint8_t*>(dev_scratch(need));
    uint8_t* hst = static_cast<uint8_t*>(pinned_scratch(prog + 1024));
    if (!arena || !hst) return LHB200_ENOMEM;
    Plan pl;
    int32_t rc = build_plan(pl, arena, need, 256, describe);
    if (rc) return rc;
    std::vector<HashOp> ops_sorted;
    std::vector<int32_t> waves;
    plan_finalize_program(pl, ops_sorted, waves);
    rc = plan_upload(pl, s, ops_sorted, waves, hst);
    if (rc) return rc;
    rc = plan_enqueue(pl, s);
    if (rc) return rc;
    if (root & OP_ZERO_FLAG) {
        memcpy(hst + prog, c.zero_hashes[root & 0xff], 32);
        LHB_CUDA(cudaMemcpyAsync(d_out32, hst + prog, 32, cudaMemcpyHostToDevice, s));
    } else {
        LHB_CUDA(cudaMemcpyAsync(d_out32, reinterpret_cast<void*>(root), 32, cudaMemcpyDeviceToDevice, s));
    }
    // the scratch arena and pinned staging are reused by the next call: finish before returning
    LHB_CUDA(cudaStreamSynchronize(s));
    return LHB200_OK;
}

int32_t lhb200_mix_in_length(const uint8_t root[32], uint64_t len, uint8_t out[32]) {
    LHB_REQUIRE_READY();
    if (!root || !out) return LHB200_EINVAL;
    return run_simple(root, 32, out, [&](Plan& p, uint8_t* d_in) {
        return p.mix_in_length(reinterpret_cast<uint64_t>(d_in), len);
    });
}

int32_t lhb200_zero_hash(uint32_t depth, uint8_t out[32]) {
    LHB_REQUIRE_READY();
    if (depth > 64 || !out) return LHB200_EINVAL;
    memcpy(out, ctx().zero_hashes[depth], 32);
    return LHB200_OK;
}

int32_t lhb200_validators_root(const uint8_t* ssz, uint64_t n, uint8_t out[32]) {
    LHB_REQUIRE_READY();
    if (!out || (n && !ssz) || n > (1ull << 40)) return LHB200_EINVAL;
    return run_simple(ssz, n * 121, out, [&](Plan& p, uint8_t* d_in) {
        return p.mix_in_length(p.merkle_list(p.leaf_kernel(0, d_in, n), n, 40), n);
    });
}

int32_t lhb200_validator_roots(const uint8_t* ssz, uint64_t n, uint8_t* out_roots) {
    LHB_REQUIRE_READY();
    if (n == 0) return LHB200_OK;
    if (!ssz || !out_roots) return LHB200_EINVAL;
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    size_t in_pad = align_up(n * 121, 256);
    uint8_t* d = static_cast<uint8_t*>(dev_scratch(in_pad + n * 32 + 256));
    uint8_t* h = static_cast<uint8_t*>(pinned_scratch(n * 121 + n * 32));
    if (!d || !h) return LHB200_ENOMEM;
    memcpy(h, ssz, n * 121);
    LHB_CUDA(cudaMemcpyAsync(d, h, n * 121, cudaMemcpyHostToDevice, c.stream));
    k_validator_roots<<<(unsigned)ceil_div(n, VAL_PER_CTA), VAL_PER_CTA, 0, c.stream>>>(d, n, d + in_pad);
    count_launch();
    LHB_CUDA(cudaGetLastError());
    LHB_CUDA(cudaMemcpyAsync(h + n * 121, d + in_pad, n * 32, cudaMemcpyDeviceToHost, c.stream));
    LHB_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(out_roots, h + n * 121, n * 32);
    return LHB200_OK;
}

static int32_t stage_deneb(const uint8_t* ssz, uint64_t len, ShardCfg sh, lhb200_state** out) {
    LHB_REQUIRE_READY();
    if (!ssz || !out) return LHB200_EINVAL;
    if (sh.world == 0 || (sh.world & (sh.world - 1)) || sh.rank >= sh.world) {
        set_error("state shard: world must be a power of two and rank < world");
        return LHB200_EINVAL;
    }
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    const size_t lit_cap = 16384;
    Plan dry;
    uint64_t fops[MAX_FIELDS], rop;
    int32_t rc = LHB200_OK;
    std::vector<ShardedList> dry_sh;
    build_plan(dry, nullptr, 0, lit_cap, [&](Plan& p) { rc = describe_deneb(p, ssz, len, nullptr, fops, &rop, sh, &dry_sh); });
    if (rc) return rc;
    size_t prog = align_up(dry.ops.size() * sizeof(HashOp), 256) + align_up((dry.ops.size() + 2) * 4, 256) + 512;
    size_t need = align_up(dry.bump, 256) + prog + (1 + MAX_FIELDS) * 32 + (1 + MAX_FIELDS) * sizeof(HashOp) + 1024;
    std::unique_ptr<lhb200_state> st(new lhb200_state());
    if (g_spare_arena && g_spare_bytes >= need) {       // recycled from the last released handle (no cudaMalloc)
        st->arena = g_spare_arena;
        st->arena_bytes = g_spare_bytes;
        g_spare_arena = nullptr;
        g_spare_bytes = 0;
    } else {
        LHB_CUDA(cudaMalloc(reinterpret_cast<void**>(&st->arena), need + (need >> 3)));
        st->arena_bytes = need + (need >> 3);
    }
    std::vector<StageCopy> copies;
    rc = build_plan(st->plan, st->arena, st->arena_bytes, lit_cap, [&](Plan& p) {
        st->sharded.clear();
        rc = describe_deneb(p, ssz, len, &copies, st->field_ops, &st->root_op, sh, &st->sharded);
    });
    st->shard = sh;
    if (rc) { cudaFree(st->arena); return rc; }
    st->copies = copies;
    st->plan.ssz_base = nullptr;  // the caller's buffer is not retained
    // H2D: per-field copies into the aligned layout.  Pinned caller memory goes straight to the copy engine;
    // pageable memory is bounced through the pinned staging slab.
    cudaPointerAttributes at;
    bool pinned = cudaPointerGetAttributes(&at, ssz) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const uint8_t* src = ssz;
    std::vector<HashOp> ops_sorted;
    std::vector<int32_t> waves;
    plan_finalize_program(st->plan, ops_sorted, waves);
    size_t stage_bytes = (pinned ? 0 : len) + lit_cap + prog + (1 + MAX_FIELDS) * sizeof(HashOp) + 2048;
    uint8_t* hst = static_cast<uint8_t*>(pinned_scratch(stage_bytes));
    if (!hst) { cudaFree(st->arena); return LHB200_ENOMEM; }
    size_t ho = 0;
    if (!pinned) {
        memcpy(hst, ssz, len);
        src = hst;
        ho = align_up(len, 256);
    }
    for (const StageCopy& cp : copies) {
        size_t z0 = cp.nbytes / 256 * 256;
        LHB_CUDA(cudaMemsetAsync(cp.dst + z0, 0, cp.pad_to - z0, c.stream));
        if (cp.nbytes)
            LHB_CUDA(cudaMemcpyAsync(cp.dst, src + cp.src_off, cp.nbytes, cudaMemcpyHostToDevice, c.stream));
    }
    rc = plan_upload(st->plan, c.stream, ops_sorted, waves, hst + ho);
    if (rc) { cudaFree(st->arena); return rc; }
    // gather table: root + field roots -> contiguous result block
    HashOp gath[1 + MAX_FIELDS];
    gath[0] = {0, st->root_op, 0};
    for (int i = 0; i < MAX_FIELDS; i++) gath[i + 1] = {0, st->field_ops[i], 0};
    uint8_t* h_g = hst + stage_bytes - (1 + MAX_FIELDS) * sizeof(HashOp) - 64;
    memcpy(h_g, gath, sizeof gath);
    uint8_t* d_g = st->plan.alloc(sizeof gath);
    st->d_result = st->plan.alloc((1 + MAX_FIELDS) * 32);
    LHB_CUDA(cudaMemcpyAsync(d_g, h_g, sizeof gath, cudaMemcpyHostToDevice, c.stream));
    st->plan.root_addr = reinterpret_cast<uint64_t>(d_g);
    LHB_CUDA(cudaStreamSynchronize(c.stream));
    *out = st.release();
    return LHB200_OK;
}

int32_t lhb200_state_stage_deneb(const uint8_t* ssz, uint64_t len, lhb200_state** out) {
    return stage_deneb(ssz, len, ShardCfg(), out);
}
// Same for any post-Altair fork (LHB200_FORK_*): the describer is table-driven, the kernels are shared.
int32_t lhb200_state_stage(const uint8_t* ssz, uint64_t len, int32_t fork, lhb200_state** out) {
    ShardCfg sh;
    sh.fork = fork;
    return stage_deneb(ssz, len, sh, out);
}

int32_t lhb200_state_stage_deneb_shard(const uint8_t* ssz, uint64_t len, uint32_t rank, uint32_t world,
                                       lhb200_state** out) {
    ShardCfg sh;
    sh.rank = rank;
    sh.world = world;
    return stage_deneb(ssz, len, sh, out);
}

// Run this rank's part and return the subtree roots of the sharded lists (n_lists x 32 bytes, fixed list order).
int32_t lhb200_state_shard_roots(lhb200_state* st, uint8_t* out, uint32_t* n_lists) {
    LHB_REQUIRE_READY();
    if (!st || !out || !n_lists) return LHB200_EINVAL;
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    int32_t rc = lhb200_state_root_enqueue(st, c.stream, nullptr);
    if (rc) return rc;
    const uint32_t n = (uint32_t)st->sharded.size();
    *n_lists = n;
    if (n == 0) { LHB_CUDA(cudaStreamSynchronize(c.stream)); return LHB200_OK; }
    std::vector<HashOp> gath(n);
    for (uint32_t i = 0; i < n; i++) gath[i] = {0, st->sharded[i].local_op, 0};
    uint8_t* h = static_cast<uint8_t*>(pinned_scratch(n * sizeof(HashOp) + n * 32 + 512));
    uint8_t* d = static_cast<uint8_t*>(dev_scratch(n * sizeof(HashOp) + n * 32 + 512));
    if (!h || !d) return LHB200_ENOMEM;
    memcpy(h, gath.data(), n * sizeof(HashOp));
    uint8_t* d_res = d + ((n * sizeof(HashOp) + 255) / 256) * 256;
    LHB_CUDA(cudaMemcpyAsync(d, h, n * sizeof(HashOp), cudaMemcpyHostToDevice, c.stream));
    k_gather_nodes<<<1, 32, 0, c.stream>>>(reinterpret_cast<const HashOp*>(d), (int)n, d_res);
    count_launch();
    uint8_t* h_res = h + ((n * sizeof(HashOp) + 255) / 256) * 256;
    LHB_CUDA(cudaMemcpyAsync(h_res, d_res, n * 32, cudaMemcpyDeviceToHost, c.stream));
    LHB_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(out, h_res, n * 32);
    return LHB200_OK;
}

// Fold the all-gathered subtree roots (rank-major: gathered[(g * n_lists + l) * 32]) into the state root.
// Must be called after lhb200_state_shard_roots on the same handle (the unsharded field roots live in its arena).
static int32_t state_combine_impl(lhb200_state* st, const uint8_t* gathered, uint8_t out[32], bool on_device) {
    if (!st || !gathered || !out) return LHB200_EINVAL;
    const uint32_t n = (uint32_t)st->sharded.size(), world = st->shard.world;
    uint32_t lg = 0;
    while ((1u << lg) < world) lg++;
    return run_simple(gathered, (size_t)world * n * 32, out, [&](Plan& p, uint8_t* d_in) {
        std::vector<uint64_t> f(st->field_ops, st->field_ops + 28);   // (sharded handles are Deneb: 28 fields)
        for (uint32_t l = 0; l < n; l++) {
            const ShardedList& L = st->sharded[l];
            std::vector<uint64_t> nodes;
            for (uint32_t gidx = 0; gidx < world; gidx++)
                nodes.push_back(reinterpret_cast<uint64_t>(d_in + ((size_t)gidx * n + l) * 32));
            uint64_t r = p.small_tree(nodes, lg);
            for (uint32_t d = L.s + lg; d < L.limit_depth; d++) r = p.op_hash(r, Plan::zero_op(d));
            if (L.mix_len != UINT64_MAX) r = p.mix_in_length(r, L.mix_len);
            f[L.field] = r;
        }
        return p.container(f);
    }, on_device);
}
int32_t lhb200_state_combine(lhb200_state* st, const uint8_t* gathered, uint8_t out[32]) {
    LHB_REQUIRE_READY();
    return state_combine_impl(st, gathered, out, false);
}

// One BeaconState root over the ranks of the library's communicator (lhb200_comm_init), on a handle staged with
// lhb200_state_stage_deneb_shard(rank, world): the shard's subtree roots stay on the device, one ncclAllGather of
// n_lists x 32 B per rank, then every rank folds the top (log2(world) levels, zero ladders, length mix-ins, container)
// — all on the library's stream; the only host transfer is the 32-byte root (SURVEY.md §8e "Tree hash").
int32_t lhb200_state_root_sharded(lhb200_state* st, uint8_t out[32]) {
    LHB_REQUIRE_READY();
    if (!st || !out) return LHB200_EINVAL;
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    const uint32_t n = (uint32_t)st->sharded.size(), world = st->shard.world;
    if ((int)world != comm_world() && world != 1) { set_error("state sharded %u ways but the communicator has %d ranks", world, comm_world()); return LHB200_EINVAL; }
    int32_t rc = lhb200_state_root_enqueue(st, c.stream, nullptr);
    if (rc) return rc;
    if (n == 0) { set_error("handle is not sharded"); return LHB200_EINVAL; }
    std::vector<HashOp> gath(n);
    for (uint32_t i = 0; i < n; i++) gath[i] = {0, st->sharded[i].local_op, 0};
    if (!st->d_coll) {   // [gather ops | my roots | all roots], owned by the handle (dev_scratch is reused by the combine)
        LHB_CUDA(cudaMalloc(reinterpret_cast<void**>(&st->d_coll), 4096 + (size_t)(world + 1) * n * 32 + 512));
    }
    uint8_t* h = static_cast<uint8_t*>(pinned_scratch(n * sizeof(HashOp) + 512));
    if (!h) return LHB200_ENOMEM;
    memcpy(h, gath.data(), n * sizeof(HashOp));
    uint8_t* d_ops = st->d_coll;
    uint8_t* d_mine = st->d_coll + 4096;
    uint8_t* d_all = d_mine + ((n * 32 + 255) / 256) * 256;
    if (n * sizeof(HashOp) > 4096) return LHB200_EINVAL;
    LHB_CUDA(cudaMemcpyAsync(d_ops, h, n * sizeof(HashOp), cudaMemcpyHostToDevice, c.stream));
    k_gather_nodes<<<1, 32, 0, c.stream>>>(reinterpret_cast<const HashOp*>(d_ops), (int)n, d_mine);
    count_launch();
    rc = comm_allgather_bytes(d_mine, d_all, (size_t)n * 32, c.stream);
    if (rc) return rc;
    return state_combine_impl(st, d_all, out, true);
}

// Apply same-length mutations to a staged state (the resident analogue of BeaconState::apply_pending_mutations,
// consensus/types/src/beacon_state.rs:2459-2481): `data` replaces SSZ bytes [ssz_offset, ssz_offset + len) of the
// encoding the handle was staged from.  Big-field bytes are patched in place in HBM, small-field bytes re-pack their
// literal chunks.  List lengths / variable-part offsets must not change (re-stage for that).  The next
// lhb200_state_root re-hashes the whole state (0.85 ms at 500 k validators - cheaper than tracking dirty paths).
// n same-length mutations in one call: offsets[i], lens[i], bytes concatenated in `data`.  Ranges must not overlap.
// One H2D copy of the blob + one scatter kernel; dirty leaves are recorded for the warm path.
int32_t lhb200_state_patch_batch(lhb200_state* st, const uint64_t* offsets, const uint32_t* lens, const uint8_t* data,
                                 uint32_t n) {
    LHB_REQUIRE_READY();
    if (!st || (n && (!offsets || !lens || !data))) return LHB200_EINVAL;
    if (n == 0) return LHB200_OK;
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++) total += lens[i];
    std::vector<ScatterOp> ops;
    ops.reserve(n);
    Plan& pl = st->plan;
    uint64_t blob_off = 0;
    uint32_t lit_lo = ~0u, lit_hi = 0;
    // Range lookups by binary search over offset-sorted indices (built once per handle): O(log fields) per edit, so a
    // slot's worth of mutations (tens of thousands of 8-byte edits) costs the host well under a millisecond.
    if (st->copy_order.empty() && !st->copies.empty()) {
        st->copy_order.resize(st->copies.size());
        for (size_t k = 0; k < st->copies.size(); k++) st->copy_order[k] = (uint32_t)k;
        std::sort(st->copy_order.begin(), st->copy_order.end(),
                  [&](uint32_t a, uint32_t b) { return st->copies[a].src_off < st->copies[b].src_off; });
        st->copy_tree.assign(st->copies.size(), -1);
        for (size_t k = 0; k < st->copies.size(); k++)
            for (size_t t = 0; t < st->trees.size(); t++)
                if (st->trees[t].src_off == st->copies[k].src_off) st->copy_tree[k] = (int32_t)t;
        st->lit_order.resize(pl.lit_src.size());
        for (size_t k = 0; k < pl.lit_src.size(); k++) st->lit_order[k] = (uint32_t)k;
        std::sort(st->lit_order.begin(), st->lit_order.end(),
                  [&](uint32_t a, uint32_t b) { return pl.lit_src[a].src_off < pl.lit_src[b].src_off; });
    }
    if (st->incremental && st->copy_tree_trees != st->trees.size()) {   // trees appear with enable_incremental
        for (size_t k = 0; k < st->copies.size(); k++) {
            st->copy_tree[k] = -1;
            for (size_t t = 0; t < st->trees.size(); t++)
                if (st->trees[t].src_off == st->copies[k].src_off) st->copy_tree[k] = (int32_t)t;
        }
        st->copy_tree_trees = st->trees.size();
    }
    // pass 1: every edit must hit resident bytes — validated BEFORE anything is modified, so a rejected batch leaves the
    // handle (host literals, dirty lists, device copy) exactly as it was
    size_t hit = ~(size_t)0;
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t lo = offsets[i], hi = lo + lens[i];
        if (lens[i] == 0) continue;
        bool touched = false;
        if (hit < st->copy_order.size() && st->copies[st->copy_order[hit]].src_off <= lo &&
            hi <= st->copies[st->copy_order[hit]].src_off + st->copies[st->copy_order[hit]].nbytes) continue;   // same field as the last edit
        size_t k = std::partition_point(st->copy_order.begin(), st->copy_order.end(), [&](uint32_t ci) {
                       const StageCopy& cp = st->copies[ci];
                       return cp.src_off + cp.nbytes <= lo;
                   }) - st->copy_order.begin();
        if (k < st->copy_order.size() && st->copies[st->copy_order[k]].src_off < hi) { touched = true; hit = k; }
        if (!touched) {
            size_t m = std::partition_point(st->lit_order.begin(), st->lit_order.end(), [&](uint32_t li) {
                           const Plan::LitSrc& ls = pl.lit_src[li];
                           return ls.src_off + ls.n <= lo;
                       }) - st->lit_order.begin();
            if (m < st->lit_order.size() && pl.lit_src[st->lit_order[m]].src_off < hi) touched = true;
        }
        if (!touched) {
            set_error("state_patch: range [%llu, %llu) is not resident on this handle (offset table or another rank's shard)",
                      (unsigned long long)lo, (unsigned long long)hi);
            return LHB200_EINVAL;
        }
    }
    size_t last_k = ~(size_t)0;
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t lo = offsets[i], hi = lo + lens[i];
        const uint8_t* src = data + blob_off;
        bool touched = lens[i] == 0;
        // first resident range whose end lies beyond lo (consecutive edits usually hit the same field: try it first)
        size_t k = last_k;
        if (!(k < st->copy_order.size() && st->copies[st->copy_order[k]].src_off <= lo &&
              lo < st->copies[st->copy_order[k]].src_off + st->copies[st->copy_order[k]].nbytes))
            k = std::partition_point(st->copy_order.begin(), st->copy_order.end(), [&](uint32_t ci) {
                    const StageCopy& cp = st->copies[ci];
                    return cp.src_off + cp.nbytes <= lo;
                }) - st->copy_order.begin();
        last_k = k;
        for (; k < st->copy_order.size(); k++) {
            const uint32_t ci = st->copy_order[k];
            const StageCopy& cp = st->copies[ci];
            if (cp.src_off >= hi) break;
            const uint64_t a = std::max<uint64_t>(lo, cp.src_off), b = std::min<uint64_t>(hi, cp.src_off + cp.nbytes);
            if (a >= b) continue;
            ops.push_back({cp.dst + (a - cp.src_off), (uint32_t)(b - a), (uint32_t)(blob_off + (a - lo))});
            touched = true;
            if (st->incremental && !st->need_full) {   // warm path: which leaves of which tree does this touch?
                const int32_t ti = st->copy_tree[ci];
                if (ti < 0) { st->need_full = true; continue; }   // a list without a resident tree (votes, summaries, ...)
                lhb200_state::Tree& t = st->trees[ti];
                const uint64_t i0 = (a - t.src_off) / t.item_bytes, i1 = (b - 1 - t.src_off) / t.item_bytes;
                if (t.n_marks + (i1 - i0 + 1) > 4ull * lhb200_state::DIRTY_CAP) { st->need_full = true; continue; }
                for (uint64_t q = i0; q <= i1; q++) t.mark(q);
            }
        }
        size_t m = std::partition_point(st->lit_order.begin(), st->lit_order.end(), [&](uint32_t li) {
                       const Plan::LitSrc& ls = pl.lit_src[li];
                       return ls.src_off + ls.n <= lo;
                   }) - st->lit_order.begin();
        for (; m < st->lit_order.size(); m++) {          // small fixed fields live in host-packed literal chunks
            const Plan::LitSrc& ls = pl.lit_src[st->lit_order[m]];
            if (ls.src_off >= hi) break;
            const uint64_t a = std::max<uint64_t>(lo, ls.src_off), b = std::min<uint64_t>(hi, ls.src_off + ls.n);
            if (a >= b) continue;
            memcpy(&pl.lit[(size_t)ls.lit_index * 32] + (a - ls.src_off), src + (a - lo), b - a);
            lit_lo = std::min(lit_lo, ls.lit_index);
            lit_hi = std::max(lit_hi, ls.lit_index + 1);
            touched = true;
        }
        if (!touched) {
            set_error("state_patch: range [%llu, %llu) is not resident on this handle (offset table or another rank's shard)",
                      (unsigned long long)lo, (unsigned long long)hi);
            return LHB200_EINVAL;
        }
        blob_off += lens[i];
    }
    const size_t ob = align_up(ops.size() * sizeof(ScatterOp), 256), bb = align_up(total + 16, 256);
    const size_t lb = lit_lo < lit_hi ? (size_t)(lit_hi - lit_lo) * 32 : 0;
    uint8_t* h = static_cast<uint8_t*>(pinned_scratch(ob + bb + lb + 256));
    if (!h) return LHB200_ENOMEM;
    if (!ops.empty()) {
        uint8_t* d = static_cast<uint8_t*>(dev_scratch(ob + bb));
        if (!d) return LHB200_ENOMEM;
        memcpy(h, ops.data(), ops.size() * sizeof(ScatterOp));
        memcpy(h + ob, data, total);
        LHB_CUDA(cudaMemcpyAsync(d, h, ob + total, cudaMemcpyHostToDevice, c.stream));
        k_scatter_bytes<<<(unsigned)ceil_div(ops.size(), 8), 256, 0, c.stream>>>(reinterpret_cast<const ScatterOp*>(d),
                                                                                 (uint32_t)ops.size(), d + ob);
        count_launch();
        LHB_CUDA(cudaGetLastError());
    }
    if (lb) {
        memcpy(h + ob + bb, &pl.lit[(size_t)lit_lo * 32], lb);
        LHB_CUDA(cudaMemcpyAsync(pl.arena + pl.lit_off + (size_t)lit_lo * 32, h + ob + bb, lb, cudaMemcpyHostToDevice, c.stream));
    }
    LHB_CUDA(cudaStreamSynchronize(c.stream));   // the staging slabs are reused by the next call
    return LHB200_OK;
}
int32_t lhb200_state_patch(lhb200_state* st, uint64_t ssz_offset, const uint8_t* data, uint64_t len) {
    if (len > 0xffffffffull) { set_error("state_patch: a single patch is limited to 4 GiB"); return LHB200_EINVAL; }
    const uint32_t l32 = (uint32_t)len;
    return lhb200_state_patch_batch(st, &ssz_offset, &l32, data, len ? 1 : 0);
}

// (Re)build every level of every resident tree from its leaf chunks (which a cold root has just refreshed).
static int32_t state_build_levels(lhb200_state* st, cudaStream_t s) {
    for (lhb200_state::Tree& t : st->trees) {
        uint64_t n = t.dev.n_leaves;
        for (uint32_t l = 0; l < t.dev.top; l++) {
            k_tree_level<<<(unsigned)ceil_div(ceil_div(n, 2), 256), 256, 0, s>>>(t.dev.lvl[l], n, t.dev.lvl[l + 1], l);
            count_launch();
            n = ceil_div(n, 2);
        }
        t.dirty.clear();
        std::fill(t.dirty_bits.begin(), t.dirty_bits.end(), 0ull);
        t.n_marks = 0;
    }
    st->need_full = false;
    LHB_CUDA(cudaGetLastError());
    return LHB200_OK;
}
// Warm root: re-hash the paths above the dirty leaves (one CTA per tree), then the tail program.
static int32_t state_incremental_enqueue(lhb200_state* st, cudaStream_t s) {
    uint32_t total = 0;
    for (lhb200_state::Tree& t : st->trees) {
        t.dirty.clear();
        if (t.n_marks) {
            for (size_t w = 0; w < t.dirty_bits.size(); w++) {
                uint64_t bits = t.dirty_bits[w];
                while (bits) {
                    t.dirty.push_back((uint32_t)(w * 64 + (uint32_t)__builtin_ctzll(bits)));
                    bits &= bits - 1;
                }
                t.dirty_bits[w] = 0;
            }
            t.n_marks = 0;
        }
        if (t.dirty.size() > lhb200_state::DIRTY_CAP) {   // too many distinct leaves for the warm buffers: cold root instead
            st->need_full = true;
            return 1;   // (positive: not an error) the caller falls back to the cold path, which also rebuilds the levels
        }
        total += (uint32_t)t.dirty.size();
    }
    uint64_t hashes = st->plan.ops.size();
    if (total) {
        const size_t tb = st->trees.size() * sizeof(TreeDev);
        uint8_t* h = static_cast<uint8_t*>(pinned_scratch(tb + (size_t)total * 4 + 256));
        if (!h) return LHB200_ENOMEM;
        uint32_t* hd = reinterpret_cast<uint32_t*>(h + align_up(tb, 256));
        uint32_t off = 0;
        for (size_t k = 0; k < st->trees.size(); k++) {
            lhb200_state::Tree& t = st->trees[k];
            t.dev.dirty = st->d_dirty + off;
            t.dev.n_dirty = (uint32_t)t.dirty.size();
            if (!t.dirty.empty()) memcpy(hd + off, t.dirty.data(), t.dirty.size() * 4);
            off += t.dev.n_dirty;
            memcpy(h + k * sizeof(TreeDev), &t.dev, sizeof(TreeDev));
            // hashes = distinct parents per level (+ 8 per dirty validator): neighbours d[j-1] < d[j] have distinct
            // ancestors exactly at the levels up to the highest bit in which they differ
            if (!t.dirty.empty()) {
                uint64_t cnt = t.dev.top;  // the path of the first dirty leaf
                for (size_t j = 1; j < t.dirty.size(); j++) {
                    const uint32_t hb = 32 - (uint32_t)__builtin_clz(t.dirty[j] ^ t.dirty[j - 1]);  // ancestors equal from level hb up
                    cnt += std::min<uint32_t>(hb - 1, t.dev.top);
                }
                hashes += cnt + (t.dev.kind == 0 ? 8ull * t.dirty.size() : 0);
            }
            t.dirty.clear();
        }
        LHB_CUDA(cudaMemcpyAsync(st->d_trees, h, tb, cudaMemcpyHostToDevice, s));
        LHB_CUDA(cudaMemcpyAsync(st->d_dirty, hd, (size_t)total * 4, cudaMemcpyHostToDevice, s));
        // one launch per level covers every tree (blockIdx.y); levels are ordered by the stream
        uint32_t max_nd = 0, max_top = 0;
        bool any_validators = false;
        for (const lhb200_state::Tree& t : st->trees) {
            max_nd = std::max(max_nd, t.dev.n_dirty);
            if (t.dev.n_dirty) { max_top = std::max(max_top, t.dev.top); any_validators |= t.dev.kind == 0; }
        }
        const dim3 grid((unsigned)ceil_div(max_nd, 256), (unsigned)st->trees.size());
        for (int l = any_validators ? -1 : 0; l < (int)max_top; l++) {
            k_tree_update_level<<<grid, 256, 0, s>>>(st->d_trees, l);
            count_launch();
        }
    }
    st->last_root_hashes = hashes;
    return plan_enqueue_tail(st->plan, s);
}

// Switch a resident (unsharded) state to the warm path: allocate and build the level arrays of its big lists
// (validators, balances, inactivity scores, participation x2, randao mixes, block/state roots, slashings).
// Afterwards lhb200_state_patch marks dirty leaves and lhb200_state_root re-hashes only the paths above them
// (plus the tail program); patches outside those lists, or more than 65 536 dirty leaves, fall back to a cold root.
int32_t lhb200_state_enable_incremental(lhb200_state* st) {
    LHB_REQUIRE_READY();
    if (!st) return LHB200_EINVAL;
    if (st->shard.world != 1) { set_error("incremental roots need the whole state on this handle (world == 1)"); return LHB200_EINVAL; }
    if (st->incremental) return LHB200_OK;
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    size_t bytes = 0;
    for (const Plan::TreeSpec& ts : st->plan.trees) {
        if (ts.src_off == ~0ull) continue;
        for (uint64_t n = ceil_div(ts.n_chunks, 2);; n = ceil_div(n, 2)) { bytes += align_up(n * 32, 256); if (n == 1) break; }
    }
    LHB_CUDA(cudaMalloc(reinterpret_cast<void**>(&st->d_levels), bytes + 256));
    size_t off = 0;
    for (const Plan::TreeSpec& ts : st->plan.trees) {
        if (ts.src_off == ~0ull) continue;
        lhb200_state::Tree t;
        memset(&t.dev, 0, sizeof t.dev);
        t.dev.src = ts.src;
        t.dev.kind = (uint32_t)ts.kind;
        t.dev.n_leaves = ts.n_chunks;
        t.dev.top = ceil_log2(ts.n_chunks);
        t.dev.top_dst = reinterpret_cast<uint8_t*>(ts.top_addr);
        t.dev.lvl[0] = const_cast<uint8_t*>(ts.chunks);
        uint64_t n = ts.n_chunks;
        for (uint32_t l = 1; l <= t.dev.top; l++) {
            n = ceil_div(n, 2);
            t.dev.lvl[l] = st->d_levels + off;
            off += align_up(n * 32, 256);
        }
        t.src_off = ts.src_off; t.src_bytes = ts.src_bytes; t.item_bytes = ts.item_bytes;
        t.dirty_bits.assign((ts.n_chunks + 63) / 64, 0ull);
        st->trees.push_back(t);
    }
    LHB_CUDA(cudaMalloc(reinterpret_cast<void**>(&st->d_trees), st->trees.size() * sizeof(TreeDev) + 256));
    LHB_CUDA(cudaMalloc(reinterpret_cast<void**>(&st->d_dirty), (size_t)lhb200_state::DIRTY_CAP * 4 * st->trees.size()));
    st->incremental = true;
    st->need_full = true;   // the first root after enabling is cold and builds the levels
    return LHB200_OK;
}
uint64_t lhb200_state_last_root_hashes(const lhb200_state* st) { return st ? st->last_root_hashes : 0; }

int32_t lhb200_state_root_enqueue(lhb200_state* st, void* stream, const void** d_root) {
    LHB_REQUIRE_READY();
    if (!st) return LHB200_EINVAL;
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx().stream;
    if (!st->e_k0) { cudaEventCreate(&st->e_k0); cudaEventCreate(&st->e_k1); }
    int32_t rc;
    rc = 1;
    if (st->incremental && !st->need_full) rc = state_incremental_enqueue(st, s);
    if (rc > 0) {
        rc = plan_enqueue(st->plan, s, st->e_k0, st->e_k1);
        st->last_root_hashes = st->plan.hash_units;
        if (!rc && st->incremental) rc = state_build_levels(st, s);   // a cold root leaves the level arrays stale
    }
    if (rc) return rc;
    k_gather_nodes<<<1, 64, 0, s>>>(reinterpret_cast<const HashOp*>(st->plan.root_addr), 1 + MAX_FIELDS, st->d_result);
    count_launch();
    LHB_CUDA(cudaGetLastError());
    if (d_root) *d_root = st->d_result;
    return LHB200_OK;
}

int32_t lhb200_state_root(lhb200_state* st, uint8_t out[32], uint8_t* field_roots) {
    LHB_REQUIRE_READY();
    if (!st || !out) return LHB200_EINVAL;
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    int32_t rc = lhb200_state_root_enqueue(st, c.stream, nullptr);
    if (rc) return rc;
    uint8_t* h = static_cast<uint8_t*>(pinned_scratch((1 + MAX_FIELDS) * 32));
    if (!h) return LHB200_ENOMEM;
    LHB_CUDA(cudaMemcpyAsync(h, st->d_result, (1 + MAX_FIELDS) * 32, cudaMemcpyDeviceToHost, c.stream));
    LHB_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(out, h, 32);
    // 28 field roots for every fork up to Deneb (the unused tail is zero chunks), 37 for an Electra handle
    if (field_roots) memcpy(field_roots, h + 32, (st->shard.fork == LHB200_FORK_ELECTRA ? MAX_FIELDS : 28) * 32);
    return LHB200_OK;
}

int32_t lhb200_state_release(lhb200_state* st) {
    if (!st) return LHB200_OK;
    std::lock_guard<std::recursive_mutex> g(ctx().mu);
    if (ctx().ready) cudaStreamSynchronize(ctx().stream);
    if (st->arena) {  // keep one arena around for the next stage call (host-buffer entry point re-stages every call)
        if (g_spare_arena) cudaFree(g_spare_arena);
        g_spare_arena = st->arena;
        g_spare_bytes = st->arena_bytes;
    }
    if (st->e_k0) cudaEventDestroy(st->e_k0);
    if (st->e_k1) cudaEventDestroy(st->e_k1);
    if (st->d_levels) cudaFree(st->d_levels);
    if (st->d_trees) cudaFree(st->d_trees);
    if (st->d_dirty) cudaFree(st->d_dirty);
    if (st->d_coll) cudaFree(st->d_coll);
    delete st;
    return LHB200_OK;
}

uint64_t lhb200_state_hash_units(const lhb200_state* st) { return st ? st->plan.hash_units : 0; }

float lhb200_state_dominant_kernel_ms(const lhb200_state* st) {
    float ms = -1.f;
    if (!st || !st->e_k0 || cudaEventElapsedTime(&ms, st->e_k0, st->e_k1) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    return ms;
}

int32_t lhb200_beacon_state_root_deneb(const uint8_t* ssz, uint64_t len, uint8_t out[32], uint8_t* field_roots) {
    return lhb200_beacon_state_root(ssz, len, LHB200_FORK_DENEB, out, field_roots);
}
// BeaconState::update_tree_hash_cache for any post-Altair variant of the superstruct (beacon_state.rs:224-571).
// field_roots (optional): 28 x 32 bytes (entries beyond the fork's field count are the zero chunk); 37 x 32 for Electra.
int32_t lhb200_beacon_state_root(const uint8_t* ssz, uint64_t len, int32_t fork, uint8_t out[32], uint8_t* field_roots) {
    LHB_REQUIRE_READY();
    lhb200_state* st = nullptr;
    int32_t rc = lhb200_state_stage(ssz, len, fork, &st);
    if (rc) return rc;
    rc = lhb200_state_root(st, out, field_roots);
    lhb200_state_release(st);
    return rc;
}

int32_t lhb200_merkle_tree_proof(const uint8_t* leaves, uint64_t n, uint32_t depth, uint64_t index, uint8_t root[32],
                                 uint8_t* branch) {
    LHB_REQUIRE_READY();
    if (!root || (depth && !branch) || (n && !leaves) || depth > 32 || n > (1ull << depth) ||
        (depth < 64 && index >= (1ull << depth))) {
        set_error("merkle_tree_proof: bad arguments");
        return LHB200_EINVAL;
    }
    // Level-by-level device build (every level materialised so siblings can be read back), then gather.
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    std::vector<uint64_t> cnt(depth + 1);
    cnt[0] = n;
    size_t total = align_up(std::max<uint64_t>(n, 1) * 32, 256);
    for (uint32_t l = 1; l <= depth; l++) {
        cnt[l] = ceil_div(cnt[l - 1], 2);
        total += align_up(std::max<uint64_t>(cnt[l], 1) * 32, 256);
    }
    size_t gath_bytes = (depth + 1) * sizeof(HashOp);
    uint8_t* d = static_cast<uint8_t*>(dev_scratch(total + gath_bytes + (depth + 1) * 32 + 1024));
    uint8_t* h = static_cast<uint8_t*>(pinned_scratch(n * 32 + gath_bytes + (depth + 1) * 32 + 512));
    if (!d || !h) return LHB200_ENOMEM;
    if (n) {
        memcpy(h, leaves, n * 32);
        LHB_CUDA(cudaMemcpyAsync(d, h, n * 32, cudaMemcpyHostToDevice, c.stream));
    }
    std::vector<uint8_t*> lvl(depth + 1);
    size_t off = 0;
    for (uint32_t l = 0; l <= depth; l++) {
        lvl[l] = d + off;
        off += align_up(std::max<uint64_t>(cnt[l], 1) * 32, 256);
    }
    for (uint32_t l = 0; l < depth && cnt[l] > 0; l++) {
        if (cnt[l] == 1 && l > 0 && cnt[l + 1] == 1) {
            // lone node climbing a zero ladder: still one hash per level
        }
        MerkleSegTable tab;
        tab.n = 1;
        tab.s[0].in = lvl[l]; tab.s[0].out = lvl[l + 1]; tab.s[0].n_in = cnt[l]; tab.s[0].level_in = l;
        tab.s[0].tile_log = 1; tab.s[0].cta_begin = 0; tab.s[0].n_tiles = (uint32_t)cnt[l + 1];
        k_merkle_reduce<<<(unsigned)cnt[l + 1], REDUCE_THREADS, 0, c.stream>>>(tab);
        count_launch();
    }
    LHB_CUDA(cudaGetLastError());
    std::vector<HashOp> gath(depth + 1);
    gath[0] = {0, n ? reinterpret_cast<uint64_t>(lvl[depth]) : (OP_ZERO_FLAG | depth), 0};
    uint64_t idx = index;
    for (uint32_t l = 0; l < depth; l++) {
        uint64_t sib = idx ^ 1;
        gath[l + 1] = {0, sib < cnt[l] ? reinterpret_cast<uint64_t>(lvl[l] + 32 * sib) : (OP_ZERO_FLAG | l), 0};
        idx >>= 1;
    }
    uint8_t* h_g = h + align_up(n * 32, 256);
    memcpy(h_g, gath.data(), gath_bytes);
    uint8_t* d_g = d + total;
    uint8_t* d_res = d_g + align_up(gath_bytes, 256);
    LHB_CUDA(cudaMemcpyAsync(d_g, h_g, gath_bytes, cudaMemcpyHostToDevice, c.stream));
    k_gather_nodes<<<(depth + 32) / 32, 32, 0, c.stream>>>(reinterpret_cast<const HashOp*>(d_g), depth + 1, d_res);
    count_launch();
    uint8_t* h_res = h_g + align_up(gath_bytes, 256);
    LHB_CUDA(cudaMemcpyAsync(h_res, d_res, (depth + 1) * 32, cudaMemcpyDeviceToHost, c.stream));
    LHB_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(root, h_res, 32);
    if (depth) memcpy(branch, h_res + 32, depth * 32);
    return LHB200_OK;
}

int32_t lhb200_verify_merkle_proofs(const uint8_t* leaves, const uint8_t* branches, uint32_t depth,
                                    const uint64_t* indices, const uint8_t* roots, uint64_t n, uint8_t* ok) {
    LHB_REQUIRE_READY();
    if (n == 0) return LHB200_OK;
    if (!leaves || !indices || !roots || !ok || (depth && !branches)) return LHB200_EINVAL;
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    size_t b_leaves = align_up(n * 32, 256), b_br = align_up(n * depth * 32 + 32, 256), b_idx = align_up(n * 8, 256),
           b_roots = align_up(n * 32, 256), b_ok = align_up(n, 256);
    size_t tot = b_leaves + b_br + b_idx + b_roots + b_ok;
    uint8_t* d = static_cast<uint8_t*>(dev_scratch(tot));
    uint8_t* h = static_cast<uint8_t*>(pinned_scratch(tot));
    if (!d || !h) return LHB200_ENOMEM;
    memcpy(h, leaves, n * 32);
    if (depth) memcpy(h + b_leaves, branches, n * depth * 32);
    memcpy(h + b_leaves + b_br, indices, n * 8);
    memcpy(h + b_leaves + b_br + b_idx, roots, n * 32);
    LHB_CUDA(cudaMemcpyAsync(d, h, tot - b_ok, cudaMemcpyHostToDevice, c.stream));
    k_verify_branches<<<(unsigned)ceil_div(n, 128), 128, 0, c.stream>>>(
        d, d + b_leaves, depth, reinterpret_cast<const uint64_t*>(d + b_leaves + b_br), d + b_leaves + b_br + b_idx, n,
        d + tot - b_ok);
    count_launch();
    LHB_CUDA(cudaGetLastError());
    LHB_CUDA(cudaMemcpyAsync(h + tot - b_ok, d + tot - b_ok, n, cudaMemcpyDeviceToHost, c.stream));
    LHB_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(ok, h + tot - b_ok, n);
    return LHB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// BeaconBlockDeneb (mainnet preset): BeaconBlock::canonical_root (consensus/types/src/beacon_block.rs:158-160).
// Layouts: beacon_block.rs:56-78 (84-byte fixed part), beacon_block_body.rs:70-121 (392), execution_payload.rs:54-95
// (528); operation containers as cited in include/lhb200.h.  The host only walks SSZ offsets: every packed byte string
// (transactions, signatures, pubkeys, bit lists, index lists, proofs, blooms) is hashed by k_byte_items straight
// from the staged blob, fixed 8/20/32-byte fields become literal chunks, and the container structure above them is
// a hash program.  `s` = host bytes, `d` = the same bytes on the device.
}  // extern "C"
namespace {
struct BlockDescriber {
    Plan& p;
    const uint8_t* s;
    const uint8_t* d;
    bool bad = false;
    bool blinded = false;   // BlindedBeaconBlock: field 9 of the body is an ExecutionPayloadHeader
    int32_t fork = LHB200_FORK_DENEB;   // beacon_block_body.rs superstruct variant: Altair 9 body fields (no payload),
                                        // Bellatrix 10 (14-field payload), Capella 11 (+ withdrawals, BLS changes), Deneb 12

    uint64_t u64(uint64_t off) { return p.literal_bytes(s + off, 8); }
    uint64_t h256(uint64_t off) { return p.literal_bytes(s + off, 32); }
    uint64_t addr20(uint64_t off) { return p.literal_bytes(s + off, 20); }
    uint64_t blob(uint64_t off, uint64_t n, uint32_t depth) { return p.bytes_item(d + off, n, depth, false); }
    uint64_t sig(uint64_t off) { return blob(off, 96, 2); }
    uint64_t pubkey(uint64_t off) { return blob(off, 48, 1); }
    uint64_t checkpoint(uint64_t off) { return p.op_hash(u64(off), h256(off + 8)); }
    uint64_t att_data(uint64_t off) {  // 128 B (attestation_data.rs:28)
        return p.container({u64(off), u64(off + 8), h256(off + 16), checkpoint(off + 48), checkpoint(off + 88)});
    }
    uint64_t signed_header(uint64_t off) {  // 208 B
        uint64_t h = p.container({u64(off), u64(off + 8), h256(off + 16), h256(off + 48), h256(off + 80)});
        return p.op_hash(h, sig(off + 112));
    }
    uint64_t proposer_slashing(uint64_t off) { return p.op_hash(signed_header(off), signed_header(off + 208)); }
    uint64_t indexed_attestation(uint64_t off, uint64_t len) {
        if (len < 228 || rd32(s + off) != 228 || (len - 228) % 8 || (len - 228) / 8 > 2048) { bad = true; return 0; }
        uint64_t idx = p.bytes_item(d + off + 228, len - 228, 9, true, (len - 228) / 8);
        return p.container({idx, att_data(off + 4), sig(off + 132)});
    }
    uint64_t attestation(uint64_t off, uint64_t len) {
        if (len < 229 || rd32(s + off) != 228 || s[off + len - 1] == 0) { bad = true; return 0; }
        const uint64_t nb = len - 228;
        const uint8_t last = s[off + len - 1];
        int top = 7;
        while (!((last >> top) & 1)) top--;
        const uint64_t bitlen = 8 * (nb - 1) + (uint64_t)top;
        if (bitlen > 2048) { bad = true; return 0; }
        // drop the delimiter: either the whole last byte (top == 0) or its top bit
        uint64_t bits = p.bytes_item(d + off + 228, (bitlen + 7) / 8, 3, true, bitlen, top ? (uint32_t)((1u << top) - 1) : 0xff);
        return p.container({bits, att_data(off + 4), sig(off + 132)});
    }
    uint64_t deposit(uint64_t off) {  // 1240 B
        uint64_t data = p.container({pubkey(off + 1056), h256(off + 1104), u64(off + 1136), sig(off + 1144)});
        return p.op_hash(blob(off, 33 * 32, 6), data);
    }
    uint64_t voluntary_exit(uint64_t off) { return p.op_hash(p.op_hash(u64(off), u64(off + 8)), sig(off + 16)); }
    uint64_t bls_change(uint64_t off) {
        return p.op_hash(p.container({u64(off), pubkey(off + 8), addr20(off + 56)}), sig(off + 76));
    }
    uint64_t withdrawal(uint64_t off) { return p.container({u64(off), u64(off + 8), addr20(off + 16), u64(off + 36)}); }
    uint64_t list_of(std::vector<uint64_t>& roots, uint32_t limit_log) {
        uint64_t n = roots.size();
        uint64_t r = p.small_tree(roots, std::min<uint32_t>(limit_log, ceil_log2(std::max<uint64_t>(n, 1))));
        if (n == 0) r = Plan::zero_op(limit_log);
        else for (uint32_t l = ceil_log2(n); l < limit_log; l++) r = p.op_hash(r, Plan::zero_op(l));
        return p.mix_in_length(r, n);
    }
    template <class F>
    uint64_t fixed_list(uint64_t off, uint64_t len, uint32_t item, uint32_t limit_log, F&& f) {
        if (len % item || len / item > (1ull << limit_log)) { bad = true; return 0; }
        std::vector<uint64_t> roots;
        for (uint64_t i = 0; i < len / item; i++) roots.push_back(f(off + item * i));
        return list_of(roots, limit_log);
    }
    // offsets table of a list of variable-size items occupying [off, off+len)
    bool var_bounds(uint64_t off, uint64_t len, uint64_t max_n, std::vector<uint64_t>& b) {
        b.clear();
        if (len == 0) { b.push_back(0); return true; }
        if (len < 4) return false;
        const uint32_t first = rd32(s + off);
        if (first % 4 || first == 0 || first > len || first / 4 > max_n) return false;
        for (uint32_t i = 0; i < first / 4; i++) b.push_back(rd32(s + off + 4 * i));
        b.push_back(len);
        for (size_t i = 0; i + 1 < b.size(); i++)
            if (b[i] > b[i + 1]) return false;
        return true;
    }
    uint64_t payload(uint64_t off, uint64_t len) {
        // fixed part: 508 B (Bellatrix: ... transactions offset), 512 (Capella: + withdrawals offset), 528 (Deneb: + blob gas)
        const bool has_wd = fork >= LHB200_FORK_CAPELLA, has_blob = fork >= LHB200_FORK_DENEB;
        const uint32_t fixed = has_blob ? 528 : has_wd ? 512 : 508;
        if (len < fixed) { bad = true; return 0; }
        const uint32_t o_extra = rd32(s + off + 436), o_tx = rd32(s + off + 504), o_wd = has_wd ? rd32(s + off + 508) : (uint32_t)len;
        if (o_extra != fixed || o_tx < o_extra || o_tx - o_extra > 32 || o_wd < o_tx || o_wd > len || (len - o_wd) % 44 ||
            (len - o_wd) / 44 > 16) { bad = true; return 0; }
        std::vector<uint64_t> f(14);
        f[0] = h256(off); f[1] = addr20(off + 32); f[2] = h256(off + 52); f[3] = h256(off + 84);
        f[4] = blob(off + 116, 256, 3);
        f[5] = h256(off + 372); f[6] = u64(off + 404); f[7] = u64(off + 412); f[8] = u64(off + 420); f[9] = u64(off + 428);
        f[10] = p.bytes_item(d + off + o_extra, o_tx - o_extra, 0, true, o_tx - o_extra);
        f[11] = h256(off + 440); f[12] = h256(off + 472);
        std::vector<uint64_t> b, roots;
        if (!var_bounds(off + o_tx, o_wd - o_tx, 1u << 20, b)) { bad = true; return 0; }
        for (size_t i = 0; i + 1 < b.size(); i++)  // ByteList[2^30]: 2^25 chunks
            roots.push_back(p.bytes_item(d + off + o_tx + b[i], b[i + 1] - b[i], 25, true, b[i + 1] - b[i]));
        f[13] = list_of(roots, 20);
        if (has_wd) f.push_back(fixed_list(off + o_wd, len - o_wd, 44, 4, [&](uint64_t o) { return withdrawal(o); }));
        if (has_blob) { f.push_back(u64(off + 512)); f.push_back(u64(off + 520)); }
        return p.container(f);
    }
    // ExecutionPayloadHeaderDeneb (execution_payload_header.rs:46-87): 584-byte fixed part + extra_data
    uint64_t payload_header(uint64_t off, uint64_t len) {
        // 536-byte fixed part (Bellatrix, 14 fields), 568 (Capella: + withdrawals_root), 584 (Deneb: + blob gas)
        const bool has_wd = fork >= LHB200_FORK_CAPELLA, has_blob = fork >= LHB200_FORK_DENEB;
        const uint32_t fixed = has_blob ? 584 : has_wd ? 568 : 536;
        if (len < fixed || len > fixed + 32 || rd32(s + off + 436) != fixed) { bad = true; return 0; }
        std::vector<uint64_t> f(14);
        f[0] = h256(off); f[1] = addr20(off + 32); f[2] = h256(off + 52); f[3] = h256(off + 84);
        f[4] = blob(off + 116, 256, 3);
        f[5] = h256(off + 372); f[6] = u64(off + 404); f[7] = u64(off + 412); f[8] = u64(off + 420); f[9] = u64(off + 428);
        f[10] = p.bytes_item(d + off + fixed, len - fixed, 0, true, len - fixed);
        f[11] = h256(off + 440); f[12] = h256(off + 472);
        f[13] = h256(off + 504);            // transactions_root
        if (has_wd) f.push_back(h256(off + 536));            // withdrawals_root
        if (has_blob) { f.push_back(u64(off + 568)); f.push_back(u64(off + 576)); }
        return p.container(f);
    }
    uint64_t body(uint64_t off, uint64_t len, uint64_t dst) {
        const bool has_ep = fork >= LHB200_FORK_BELLATRIX, has_bc = fork >= LHB200_FORK_CAPELLA, has_kz = fork >= LHB200_FORK_DENEB;
        const uint32_t fixed = 380 + (has_ep ? 4 : 0) + (has_bc ? 4 : 0) + (has_kz ? 4 : 0);
        if (len < fixed) { bad = true; return 0; }
        const uint32_t o_ps = rd32(s + off + 200), o_as = rd32(s + off + 204), o_at = rd32(s + off + 208),
                       o_dp = rd32(s + off + 212), o_ex = rd32(s + off + 216),
                       o_ep = has_ep ? rd32(s + off + 380) : (uint32_t)len, o_bc = has_bc ? rd32(s + off + 384) : (uint32_t)len,
                       o_kz = has_kz ? rd32(s + off + 388) : (uint32_t)len;
        if (o_ps != fixed || o_as < o_ps || o_at < o_as || o_dp < o_at || o_ex < o_dp || o_ep < o_ex || o_bc < o_ep ||
            o_kz < o_bc || o_kz > len) { bad = true; return 0; }
        std::vector<uint64_t> f(9), b, roots;
        f[0] = sig(off);
        f[1] = p.container({h256(off + 96), u64(off + 128), h256(off + 136)});  // eth1_data.rs:27
        f[2] = h256(off + 168);
        f[3] = fixed_list(off + o_ps, o_as - o_ps, 416, 4, [&](uint64_t o) { return proposer_slashing(o); });
        if (!var_bounds(off + o_as, o_at - o_as, 2, b)) { bad = true; return 0; }
        for (size_t i = 0; i + 1 < b.size(); i++) {
            const uint64_t q = off + o_as + b[i], ql = b[i + 1] - b[i];
            if (ql < 8) { bad = true; return 0; }
            const uint32_t a1 = rd32(s + q), a2 = rd32(s + q + 4);
            if (a1 != 8 || a2 < a1 || a2 > ql) { bad = true; return 0; }
            uint64_t r1 = indexed_attestation(q + a1, a2 - a1), r2 = indexed_attestation(q + a2, ql - a2);
            if (bad) return 0;
            roots.push_back(p.op_hash(r1, r2));
        }
        f[4] = list_of(roots, 1);
        roots.clear();
        if (!var_bounds(off + o_at, o_dp - o_at, 128, b)) { bad = true; return 0; }
        for (size_t i = 0; i + 1 < b.size(); i++) {
            roots.push_back(attestation(off + o_at + b[i], b[i + 1] - b[i]));
            if (bad) return 0;
        }
        f[5] = list_of(roots, 7);
        f[6] = fixed_list(off + o_dp, o_ex - o_dp, 1240, 4, [&](uint64_t o) { return deposit(o); });
        f[7] = fixed_list(off + o_ex, o_ep - o_ex, 112, 4, [&](uint64_t o) { return voluntary_exit(o); });
        f[8] = p.op_hash(blob(off + 220, 64, 1), sig(off + 284));  // sync_aggregate.rs:38
        if (has_ep) f.push_back(blinded ? payload_header(off + o_ep, o_bc - o_ep) : payload(off + o_ep, o_bc - o_ep));
        if (has_bc) f.push_back(fixed_list(off + o_bc, o_kz - o_bc, 172, 4, [&](uint64_t o) { return bls_change(o); }));
        if (has_kz) f.push_back(fixed_list(off + o_kz, len - o_kz, 48, 12, [&](uint64_t o) { return pubkey(o); }));  // kzg_commitment.rs:51
        if (bad) return 0;
        return p.container(f, dst);
    }
    uint64_t block(uint64_t off, uint64_t len, uint64_t dst_root, uint64_t dst_body) {
        if (len < 84 || rd32(s + off + 80) != 84) { bad = true; return 0; }
        uint64_t b = body(off + 84, len - 84, dst_body);
        if (bad) return 0;
        return p.container({u64(off), u64(off + 8), h256(off + 16), h256(off + 48), b}, dst_root);
    }
};
}  // namespace
extern "C" {

constexpr int32_t LHB200_ERETRY = -1000;   // internal: the plan did not fit the arena bound of this attempt

// transactions in one BeaconBlockDeneb blob (0 when the offsets are not plausible — the describer reports that)
static uint64_t prescan_transactions(const uint8_t* blk, uint64_t len, int32_t fork) {
    auto rd = [&](uint64_t o) { uint32_t v; memcpy(&v, blk + o, 4); return (uint64_t)v; };
    if (fork < LHB200_FORK_BELLATRIX || len < 84 + 392) return 0;
    const uint64_t body = 84, o_ep = rd(body + 380), o_bc = fork >= LHB200_FORK_CAPELLA ? rd(body + 384) : len - body;
    if (o_ep > o_bc || body + o_bc > len || o_bc - o_ep < 528) return 0;
    const uint64_t pay = body + o_ep, plen = o_bc - o_ep, o_tx = rd(pay + 504),
                   o_wd = fork >= LHB200_FORK_CAPELLA ? rd(pay + 508) : plen;
    if (o_tx > o_wd || o_wd > plen || o_wd - o_tx < 4) return 0;
    const uint64_t first = rd(pay + o_tx);
    return first <= o_wd - o_tx ? first / 4 : 0;
}

static int32_t block_roots_attempt(Ctx& c, const uint8_t* ssz, const uint64_t* offsets, uint32_t n, uint8_t* roots,
                                   uint8_t* body_roots, bool blinded, int32_t fork, uint64_t base, uint64_t total, size_t in_pad,
                                   size_t max_nodes, size_t lit_cap) {
    uint8_t *d_in = nullptr, *d_roots = nullptr, *d_body = nullptr;
    bool bad = false;
    auto build = [&](Plan& p) {
        d_in = p.alloc(in_pad);
        d_roots = p.alloc(64ull * n);
        d_body = d_roots + 32ull * n;
        p.forced_base = reinterpret_cast<uint64_t>(d_roots);
        p.forced_wave.assign(2ull * n, -1);
        BlockDescriber bd{p, ssz + base, d_in};
        bd.blinded = blinded;
        bd.fork = fork;
        for (uint32_t i = 0; i < n && !bd.bad; i++)
            bd.block(offsets[i] - base, offsets[i + 1] - offsets[i], reinterpret_cast<uint64_t>(d_roots + 32ull * i),
                     reinterpret_cast<uint64_t>(d_body + 32ull * i));
        bad = bd.bad;
    };
    // One planning pass over the bounded arena (host time matters here: a block is only ~10^4 hashes).
    const size_t prog_bytes = align_up(max_nodes * sizeof(HashOp), 256) + align_up((max_nodes + 2) * 4, 256) +
                              align_up(max_nodes * sizeof(ByteItem), 256) + 1024;
    // literals | node pool (op outputs) | staged blob | roots | item outputs | program blobs
    const size_t need = lit_cap + 32 * max_nodes + in_pad + 64ull * n + 32 * max_nodes + prog_bytes + 8192;
    uint8_t* arena = static_cast<uint8_t*>(dev_scratch(need));
    const size_t stage_bytes = align_up(total, 256) + lit_cap + prog_bytes + 64ull * n + 1024;
    uint8_t* hst = static_cast<uint8_t*>(pinned_scratch(stage_bytes));
    if (!arena || !hst) return LHB200_ENOMEM;
    Plan pl;
    int32_t rc = build_plan(pl, arena, need, lit_cap, build, max_nodes);
    if (bad) { set_error("BeaconBlock SSZ: malformed offsets or lengths for this fork"); return LHB200_EINVAL; }
    if (pl.node_overflow || pl.lit.size() > lit_cap || pl.ops.size() + pl.items.size() > max_nodes ||
        pl.bump + prog_bytes > need) {
        set_error("internal: block plan exceeds its arena bound (%zu ops + %zu items of %zu nodes, %zu of %zu literal bytes, "
                  "%zu of %zu arena bytes)", pl.ops.size(), pl.items.size(), max_nodes, pl.lit.size(), lit_cap,
                  pl.bump + prog_bytes, need);
        return LHB200_ERETRY;
    }
    if (rc) return rc;
    memcpy(hst, ssz + base, total);
    memset(hst + total, 0, align_up(total, 256) - total);
    LHB_CUDA(cudaMemcpyAsync(d_in, hst, align_up(total, 256), cudaMemcpyHostToDevice, c.stream));
    std::vector<HashOp> ops_sorted;
    std::vector<int32_t> waves;
    plan_finalize_program(pl, ops_sorted, waves);
    rc = plan_upload(pl, c.stream, ops_sorted, waves, hst + align_up(total, 256));
    if (rc) return rc;
    rc = plan_enqueue(pl, c.stream);
    if (rc) return rc;
    uint8_t* h_out = hst + stage_bytes - 64ull * n - 64;
    LHB_CUDA(cudaMemcpyAsync(h_out, d_roots, 32ull * n, cudaMemcpyDeviceToHost, c.stream));
    if (body_roots) LHB_CUDA(cudaMemcpyAsync(h_out + 32ull * n, d_body, 32ull * n, cudaMemcpyDeviceToHost, c.stream));
    LHB_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(roots, h_out, 32ull * n);
    if (body_roots) memcpy(body_roots, h_out + 32ull * n, 32ull * n);
    return LHB200_OK;
}

// n BeaconBlockDeneb SSZ blobs, concatenated; offsets[n+1]; roots n*32; body_roots n*32 or NULL.
static int32_t block_roots_deneb(const uint8_t* ssz, const uint64_t* offsets, uint32_t n, uint8_t* roots,
                                 uint8_t* body_roots, bool blinded, int32_t fork = LHB200_FORK_DENEB) {
    LHB_REQUIRE_READY();
    if (fork < LHB200_FORK_ALTAIR || fork > LHB200_FORK_DENEB || (blinded && fork < LHB200_FORK_BELLATRIX)) {
        set_error("beacon_block_roots: fork id %d not supported (Altair .. Deneb; blinded blocks from Bellatrix)", fork);
        return LHB200_EINVAL;
    }
    if (!ssz || !offsets || !roots || n == 0) { set_error("beacon_block_roots: null argument or zero blocks"); return LHB200_EINVAL; }
    for (uint32_t i = 0; i < n; i++)
        if (offsets[i] > offsets[i + 1]) { set_error("beacon_block_roots: offsets not monotone"); return LHB200_EINVAL; }
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    const uint64_t base = offsets[0], total = offsets[n] - offsets[0];
    const size_t in_pad = align_up(total + 64, 256);
    // Offset pre-scan: the only SSZ shape with more than one tree node per ~14 input bytes is a run of (near-)empty
    // transactions — a 4-byte offset each, one byte item + one list node + one length literal.  Count them per block
    // (three offset reads) so the arena bound is tight for real blocks and still holds for that shape.
    uint64_t n_tx = 0;
    if (!blinded)
        for (uint32_t i = 0; i < n; i++) n_tx += prescan_transactions(ssz + offsets[i], offsets[i + 1] - offsets[i], fork);
    int32_t rc = LHB200_OK;
    for (int attempt = 0; attempt < 2; attempt++) {
        // attempt 0: transactions counted, everything else <= one node per 12 bytes and one literal per 8 bytes;
        // attempt 1 (only if a plan ever exceeds that): the unconditional bound of one node and literal per 2 bytes.
        const size_t max_nodes = attempt == 0 ? 2 * n_tx + total / 12 + 512ull * n : total / 2 + 512ull * n;
        const size_t lit_cap = align_up(32 * (attempt == 0 ? n_tx + total / 8 + 128ull * n : total / 2 + 128ull * n), 256);
        rc = block_roots_attempt(c, ssz, offsets, n, roots, body_roots, blinded, fork, base, total, in_pad, max_nodes, lit_cap);
        if (rc != LHB200_ERETRY) break;
    }
    return rc == LHB200_ERETRY ? LHB200_EINVAL : rc;
}
int32_t lhb200_beacon_block_roots_deneb(const uint8_t* ssz, const uint64_t* offsets, uint32_t n, uint8_t* roots,
                                        uint8_t* body_roots) {
    return block_roots_deneb(ssz, offsets, n, roots, body_roots, false);
}
int32_t lhb200_beacon_block_root_deneb(const uint8_t* ssz, uint64_t len, uint8_t out[32], uint8_t* body_root) {
    const uint64_t offs[2] = {0, len};
    return block_roots_deneb(ssz, offs, 1, out, body_root, false);
}
// The earlier variants of the BeaconBlock superstruct (beacon_block.rs:41-90, beacon_block_body.rs:43-110): fork is
// LHB200_FORK_ALTAIR .. LHB200_FORK_DENEB; blinded != 0 selects the BlindedBeaconBlock form (Bellatrix and later).
int32_t lhb200_beacon_block_roots(const uint8_t* ssz, const uint64_t* offsets, uint32_t n, int32_t fork, int32_t blinded,
                                  uint8_t* roots, uint8_t* body_roots) {
    return block_roots_deneb(ssz, offsets, n, roots, body_roots, blinded != 0, fork);
}
// BlindedBeaconBlock (beacon_block.rs:80): the body carries the ExecutionPayloadHeader; the root equals the full block's.
int32_t lhb200_blinded_beacon_block_roots_deneb(const uint8_t* ssz, const uint64_t* offsets, uint32_t n, uint8_t* roots,
                                                uint8_t* body_roots) {
    return block_roots_deneb(ssz, offsets, n, roots, body_roots, true);
}

// swap_or_not_shuffle::shuffle_list(input, rounds, seed, forwards) (consensus/swap_or_not_shuffle/src/shuffle_list.rs:79).
// The reference returns None for an empty list, more than 2^24 elements or zero rounds: LHB200_EINVAL here.
int32_t lhb200_shuffle_list(const uint64_t* input, uint64_t n, uint8_t rounds, const uint8_t seed[32], int32_t forwards,
                            uint64_t* out) {
    LHB_REQUIRE_READY();
    if (!input || !out || !seed || n == 0 || n > (1ull << 24) || rounds == 0) {
        set_error("shuffle_list: empty list, more than 2^24 elements or zero rounds (reference returns None)");
        return LHB200_EINVAL;
    }
    Ctx& c = ctx();
    std::lock_guard<std::recursive_mutex> g(c.mu);
    const uint32_t n_blocks = (uint32_t)ceil_div(n, 256);
    const size_t b_in = align_up(n * 8, 256), b_src = align_up((size_t)rounds * n_blocks * 32, 256),
                 b_piv = 2048 + 256;   // 255 rounds x 8-byte pivots (2040 B), then the seed in its own slot
    uint8_t* d = static_cast<uint8_t*>(dev_scratch(2 * b_in + b_src + b_piv + 256));
    uint8_t* h = static_cast<uint8_t*>(pinned_scratch(2 * b_in + 64));
    if (!d || !h) return LHB200_ENOMEM;
    uint64_t* d_in = reinterpret_cast<uint64_t*>(d);
    uint64_t* d_out = reinterpret_cast<uint64_t*>(d + b_in);
    uint8_t* d_src = d + 2 * b_in;
    uint64_t* d_piv = reinterpret_cast<uint64_t*>(d_src + b_src);
    uint8_t* d_seed = reinterpret_cast<uint8_t*>(d_piv) + 2048;
    memcpy(h, input, n * 8);
    memcpy(h + b_in, seed, 32);
    LHB_CUDA(cudaMemcpyAsync(d_in, h, n * 8, cudaMemcpyHostToDevice, c.stream));
    LHB_CUDA(cudaMemcpyAsync(d_seed, h + b_in, 32, cudaMemcpyHostToDevice, c.stream));
    const uint64_t total = std::max<uint64_t>((uint64_t)rounds * n_blocks, rounds);
    k_shuffle_hashes<<<(unsigned)ceil_div(total, 128), 128, 0, c.stream>>>(d_seed, rounds, n, n_blocks, d_piv, d_src);
    k_shuffle_permute<<<(unsigned)ceil_div(n, 256), 256, 0, c.stream>>>(d_in, d_out, n, rounds, n_blocks, d_piv, d_src,
                                                                        forwards ? 1 : 0);
    count_launch(2);
    LHB_CUDA(cudaGetLastError());
    LHB_CUDA(cudaMemcpyAsync(h, d_out, n * 8, cudaMemcpyDeviceToHost, c.stream));
    LHB_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(out, h, n * 8);
    return LHB200_OK;
}

}  // extern "C"
