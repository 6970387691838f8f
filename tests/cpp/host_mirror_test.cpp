// tests/cpp/host_mirror_test.cpp — exercises the C++ host layer (include/lhb200.hpp) the way Lighthouse code
// uses crypto/bls and tree_hash: deserialize keys/signatures, build SignatureSets, verify_signature_sets, tamper,
// MerkleTree create/prove/verify.  Vectors come from a file written by the pytest (pk48 | msg32 | sig96 records).
// Exit codes: 0 all good, 2 no usable GPU (the library has no CPU fallback), 1 a check failed.
#include <cstdio>
#include <vector>
#include "lhb200.hpp"

using namespace lhb200;

#define CHECK(c)                                                         \
    do {                                                                 \
        if (!(c)) { std::fprintf(stderr, "CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    if (lhb200_init(0) != LHB200_OK) {
        std::fprintf(stderr, "no device: %s\n", lhb200_last_error());
        return 2;
    }
    if (argc < 2) return 1;
    std::FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 1;
    std::vector<bls::PublicKey> pks;
    std::vector<bls::Signature> sigs;
    std::vector<Hash256> msgs;
    uint8_t rec[176];
    while (std::fread(rec, 1, sizeof rec, f) == sizeof rec) {
        pks.push_back(bls::PublicKey::deserialize(rec, 48));
        Hash256 m;
        std::memcpy(m.data(), rec + 48, 32);
        msgs.push_back(m);
        sigs.push_back(bls::Signature::deserialize(rec + 80, 96));
    }
    std::fclose(f);
    CHECK(!pks.empty());
    std::vector<bls::SignatureSet> sets;
    for (size_t i = 0; i < pks.size(); i++) sets.push_back(bls::SignatureSet::single_pubkey(sigs[i], pks[i], msgs[i]));
    CHECK(sets[0].verify());
    CHECK(bls::verify_signature_sets(sets.begin(), sets.end()));
    CHECK(!bls::verify_signature_sets(sets.begin(), sets.begin()));                 // empty iterator -> false
    {   // corrupt one message
        std::vector<bls::SignatureSet> bad = sets;
        bad[bad.size() / 2].message[0] ^= 1;
        CHECK(!bls::verify_signature_sets(bad.begin(), bad.end()));
    }
    {   // empty and infinity signatures
        bls::Signature e = bls::Signature::empty(), inf = bls::Signature::infinity();
        std::vector<bls::SignatureSet> bad = sets;
        bad[0].signature = &e;
        CHECK(!bls::verify_signature_sets(bad.begin(), bad.end()));
        bad[0].signature = &inf;
        CHECK(!bls::verify_signature_sets(bad.begin(), bad.end()));
        CHECK(bls::eth_fast_aggregate_verify(inf, msgs[0], {}));
        CHECK(!bls::fast_aggregate_verify(sigs[0], msgs[0], {}));
    }
    {   // aggregation surface: one message signed by every key -> aggregate -> fast_aggregate_verify
        // (the vectors are single-key triples on distinct messages: aggregate_verify is their natural check)
        std::vector<const bls::Signature*> sp;
        std::vector<const bls::PublicKey*> kp;
        for (size_t i = 0; i < pks.size(); i++) { sp.push_back(&sigs[i]); kp.push_back(&pks[i]); }
        bls::AggregateSignature agg = bls::AggregateSignature::aggregate(sp);
        CHECK(agg.aggregate_verify(msgs, kp));
        bls::AggregateSignature step;                              // add_assign one by one == aggregate at once
        for (const bls::Signature* s : sp) step.add_assign(*s);
        CHECK(step.serialize() == agg.serialize());
        std::vector<Hash256> bad = msgs;
        bad[0][5] ^= 1;
        CHECK(!agg.aggregate_verify(bad, kp));
        CHECK(!agg.aggregate_verify({}, {}));
        bls::AggregatePublicKey apk = bls::AggregatePublicKey::aggregate(kp);
        bls::PublicKey round = bls::PublicKey::deserialize_uncompressed(apk.to_public_key().serialize_uncompressed().data(), 96);
        CHECK(round == apk.to_public_key());
    }
    bool threw = false;
    try {
        uint8_t infpk[48] = {0xc0};
        bls::PublicKey::deserialize(infpk, 48);
    } catch (const Error&) { threw = true; }
    CHECK(threw);
    // merkle: create + prove + verify (consensus/merkle_proof tests :412-430)
    std::vector<Hash256> leaves(5);
    for (size_t i = 0; i < leaves.size(); i++) leaves[i].fill(static_cast<uint8_t>(i + 1));
    auto tree = merkle_proof::MerkleTree::create(leaves, 3);
    Hash256 root = tree.hash();
    for (uint64_t idx = 0; idx < 8; idx++) {
        auto pr = tree.generate_proof(idx);
        CHECK(merkle_proof::verify_merkle_proof(pr.first, pr.second, 3, idx, root));
        CHECK(!merkle_proof::verify_merkle_proof(pr.first, pr.second, 3, idx ^ 1, root) || pr.second[0] == pr.first);
    }
    Hash256 z{};
    CHECK(tree_hash::hash32_concat(z, z) == tree_hash::merkle_root(z.data(), 64 > 32 ? 32 : 32, 2));  // zero[1]
    CHECK(tree_hash::mix_in_length(root, 5) != root);
    std::printf("OK %zu sets\n", sets.size());
    return 0;
}
