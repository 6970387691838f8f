#!/usr/bin/env python3
"""Extract the reference's own in-tree golden vectors into tests/golden/ (run in the build container only;
/root/reference does not exist on the GPU box).  Sources (all under /root/reference):
  common/eth2_network_config/built_in_network_configs/{mainnet,sepolia,gnosis}/genesis.ssz.zip
      phase0 genesis BeaconState; bytes 8..40 = genesis_validators_root = hash_tree_root(validators)
      (checked by common/eth2_network_config/src/lib.rs:227-233)
  validator_manager/test_vectors/vectors/*/validator_keys/deposit_data-*.json
      22 (pubkey, signature, deposit_message_root, deposit_data_root) entries, asserted valid by
      validator_manager/src/create_validators.rs:749-768
  common/eth2_interop_keypairs/specs/keygen_10_validators.yaml  (sk -> pk, tests/generation.rs:6-64)
"""
import glob, io, json, lzma, os, struct, sys, zipfile
import yaml

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)

# phase0 BeaconState fixed part: offset of the `validators` u32 offset
VAL_OFF_POS = 8 + 32 + 8 + 16 + 112 + 2 * 8192 * 32 + 4 + 72 + 4 + 8
meta = {}
for net in ("mainnet", "sepolia", "gnosis"):
    z = zipfile.ZipFile(f"{REF}/common/eth2_network_config/built_in_network_configs/{net}/genesis.ssz.zip")
    state = z.read(z.namelist()[0])
    o_val, o_bal = struct.unpack_from("<II", state, VAL_OFF_POS)
    vals = state[o_val:o_bal]
    assert len(vals) % 121 == 0
    with lzma.open(os.path.join(OUT, f"genesis_validators_{net}.bin.xz"), "wb", preset=9) as f:
        f.write(vals)
    meta[net] = {"n_validators": len(vals) // 121, "genesis_validators_root": state[8:40].hex()}
json.dump(meta, open(os.path.join(OUT, "genesis_validators.json"), "w"), indent=1)

deps = []
for p in sorted(glob.glob(f"{REF}/validator_manager/test_vectors/vectors/*/validator_keys/deposit_data-*.json")):
    for d in json.load(open(p)):
        deps.append({k: d[k] for k in ("pubkey", "withdrawal_credentials", "amount", "signature",
                                        "deposit_message_root", "deposit_data_root", "fork_version")}
                    | {"source": os.path.relpath(p, REF)})
json.dump(deps, open(os.path.join(OUT, "deposit_data.json"), "w"), indent=1)

kp = yaml.safe_load(open(f"{REF}/common/eth2_interop_keypairs/specs/keygen_10_validators.yaml"))
json.dump(kp, open(os.path.join(OUT, "interop_keypairs.json"), "w"), indent=1)
print({k: v["n_validators"] for k, v in meta.items()}, len(deps), "deposits", len(kp), "keypairs")
