#!/usr/bin/env python3
"""bench.py — headline benchmark of the two hot paths (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" is one pass of the hot path over one batch of synthetic input:
  * BLS  (headline): bls::verify_signature_sets over 100 000 aggregate attestations x 128 pubkeys per GPU
    (BASELINE configs[2]); weak scaling — every rank verifies its own shard, one NCCL all-reduce(min) of the verdicts.
  * tree hash (second object in the same JSON line): cold tree_hash_root of a 500 000-validator Deneb
    BeaconState (BASELINE configs[1]); N > 1 = N independent states (SURVEY §8e "round-robin whole states").
`value` is device-resident throughput (inputs already in HBM, CUDA events on the launching stream, max over
ranks); `e2e` is the same metric through the C-ABI call with pinned HOST buffers, H2D and D2H inside the timed
region.  `--impl reference` times the CPU oracle (oracle/, kind "port": the reference's Rust/blst path cannot be
built here) on the host cores on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SETS = int(os.environ.get("LHB_BENCH_SETS", "100000"))
KEYS_PER_SET = int(os.environ.get("LHB_BENCH_KEYS", "128"))
N_VALIDATORS_BLS = 16384
N_VALIDATORS_STATE = int(os.environ.get("LHB_BENCH_VALIDATORS", "500000"))
BLS_BYTES_PER_SET = 96 * KEYS_PER_SET + 96 + 32 + 8      # SURVEY §8d algorithmic bytes per unit (12 424 B at k=128)
STATE_UNIT_BYTES = 96                                    # one hash32_concat: 64 B in + 32 B out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture summary (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        return json.load(open(p)).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active")
                                                          for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def dist_env():
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), ws


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import lighthouse_b200
    from lighthouse_b200 import bls, tree_hash as T, _ffi
    from lighthouse_b200.synthetic import attestation_batch, beacon_state_deneb_ssz
    lighthouse_b200.init(local_rank)
    stream = torch.cuda.Stream(device=dev)
    sp = stream.cuda_stream
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------------ BLS workload (per rank)
    ab = attestation_batch(N_SETS, keys_per_set=KEYS_PER_SET, n_validators=N_VALIDATORS_BLS, seed=0x11570002 + rank)
    n_keys = N_SETS * KEYS_PER_SET
    rng = np.random.default_rng(99 + rank)
    rands = rng.integers(1, 2 ** 63, size=N_SETS, dtype=np.uint64) * 2 + 1          # nonzero 64-bit scalars
    batch = bls.Batch(N_SETS, n_keys)
    batch.upload(ab.sigs, ab.msgs, ab.pks, ab.offsets, rands)
    verdict = torch.zeros(1, dtype=torch.int32, device=dev)

    def bls_step():
        with torch.cuda.stream(stream):
            batch.enqueue(sp)
            ok = batch.result(sp)
            verdict.fill_(1 if ok else 0)
            if world > 1:
                dist.all_reduce(verdict, op=dist.ReduceOp.MIN)       # the one collective of the path
        return ok

    for _ in range(W):
        assert bls_step(), "synthetic batch must verify"
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _ffi.lib.lhb200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dom_ms = []
    with torch.cuda.stream(stream):
        e0.record(stream)
    for _ in range(K):
        ok = bls_step()
        dom_ms.append(batch.dominant_kernel_ms)
    with torch.cuda.stream(stream):
        e1.record(stream)
    barrier()
    bls_ms = max_over_ranks(e0.elapsed_time(e1)) / K
    bls_launches = _ffi.lib.lhb200_launch_count() - l0
    clocks = sampler.stop()
    assert ok and int(verdict.item()) == 1
    bls_value = N_SETS * world / (bls_ms / 1e3)

    # e2e: pinned host buffers -> H2D -> kernels -> D2H verdict, through the staged C-ABI calls
    pin = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).pin_memory()
    h_sigs, h_msgs, h_pks = pin(ab.sigs), pin(ab.msgs), pin(ab.pks)
    h_offs = torch.from_numpy(ab.offsets.copy()).pin_memory()
    h_rands = torch.from_numpy(rands.copy()).pin_memory()
    h2d = h_sigs.numel() + h_msgs.numel() + h_pks.numel() + h_offs.numel() * 4 + h_rands.numel() * 8
    vp = lambda t: C.c_void_p(t.data_ptr())

    def bls_e2e_step():
        # streamed upload: key chunks cross the host link while k_sig_prepare / k_hash_to_g2 already run
        _ffi.check(_ffi.lib.lhb200_bls_batch_upload_async(batch._h, vp(h_sigs), vp(h_msgs), vp(h_pks), vp(h_offs),
                                                          vp(h_rands), N_SETS, sp), "upload_async")
        batch.enqueue(sp)
        return batch.result(sp)

    assert bls_e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        assert bls_e2e_step()
    torch.cuda.synchronize(dev)
    bls_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / K
    barrier()

    # e2e with the device-resident pubkey table: sets carry u32 validator indices
    table = bls.PubkeyTable(N_VALIDATORS_BLS)
    table.append(ab.pk_table.tobytes())
    h_idx = torch.from_numpy(ab.committees.reshape(-1).astype(np.uint32)).pin_memory()
    h2d_idx = h_sigs.numel() + h_msgs.numel() + h_idx.numel() * 4 + h_offs.numel() * 4 + h_rands.numel() * 8

    def bls_e2e_idx_step():
        _ffi.check(_ffi.lib.lhb200_bls_batch_upload_indexed(batch._h, table._h, vp(h_sigs), vp(h_msgs), vp(h_idx),
                                                            vp(h_offs), vp(h_rands), N_SETS), "upload_indexed")
        batch.enqueue(sp)
        return batch.result(sp)

    assert bls_e2e_idx_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        assert bls_e2e_idx_step()
    torch.cuda.synchronize(dev)
    bls_e2e_idx_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / K
    barrier()

    # ------------------------------------------------------------------ tree-hash workload (per rank)
    ssz = beacon_state_deneb_ssz(N_VALIDATORS_STATE, seed=42 + rank)
    st = T.ResidentState(ssz)
    root0 = st.root()
    roots_all = torch.zeros(world, 32, dtype=torch.uint8, device=dev)

    def state_step():
        with torch.cuda.stream(stream):
            d_root = st.enqueue(sp)
            if world > 1:   # one all-gather of the 32-byte roots
                mine = torch.frombuffer((C.c_uint8 * 32).from_buffer_copy(root0), dtype=torch.uint8).to(dev)
                dist.all_gather_into_tensor(roots_all.view(-1), mine)

    for _ in range(W):
        state_step()
    barrier()
    l0 = _ffi.lib.lhb200_launch_count()
    st_dom = []
    with torch.cuda.stream(stream):
        e0.record(stream)
    for _ in range(K):
        state_step()
    with torch.cuda.stream(stream):
        e1.record(stream)
    barrier()
    st_dom.append(st.dominant_kernel_ms)
    st_ms = max_over_ranks(e0.elapsed_time(e1)) / K
    st_launches = _ffi.lib.lhb200_launch_count() - l0
    assert st.root() == root0

    # one state sharded over all ranks: per-rank leaf ranges, one all-gather of subtree roots, combine
    sharded = None
    if world > 1 and (world & (world - 1)) == 0:
        common = beacon_state_deneb_ssz(N_VALIDATORS_STATE, seed=4242)          # the same state on every rank
        sh = T.ShardedState(common, rank, world)

        def sharded_step():
            mine = sh.shard_roots()
            t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
            out = torch.empty(world * len(mine), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(out, t)
            return sh.combine(bytes(out.cpu().tolist()))

        r_sh = sharded_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            r_sh = sharded_step()
        torch.cuda.synchronize(dev)
        sh_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / K
        full = T.beacon_state_root_deneb(common) if rank == 0 else None
        sharded = {"ms_per_root": sh_ms, "roots_per_s": 1e3 / sh_ms, "matches_single_gpu_root": (r_sh == full) if rank == 0 else None,
                   "note": "one 500k-validator state split by leaf range over all ranks; wall clock incl. the all-gather and D2H/H2D of 6x32 B"}
        sh.release()
        barrier()

    h_ssz = pin(ssz)
    out32 = C.create_string_buffer(32)

    def state_e2e_step():
        _ffi.check(_ffi.lib.lhb200_beacon_state_root_deneb(vp(h_ssz), len(ssz), out32, None), "state root")

    state_e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        state_e2e_step()
    st_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / K
    assert out32.raw == root0
    units = st.hash_units

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu_bls = cpu_state = None
    if rank == 0 and world == 1:
        from tests import oracle_lib as O
        cores = O.hw_threads()
        O.set_threads(cores)
        sample = min(N_SETS, max(1024, 256 * cores))   # ~1-3 s wall on all cores, amortises the serial final exponentiation
        t0 = time.perf_counter()
        ok_cpu = O.bls_verify_signature_sets(ab.sigs[:96 * sample], ab.msgs[:32 * sample],
                                             ab.pks[:96 * KEYS_PER_SET * sample], ab.offsets[:sample + 1], rands[:sample])
        dt = time.perf_counter() - t0
        assert ok_cpu, "CPU oracle disagrees with the GPU verdict"
        cpu_bls = {"value": sample / dt, "unit": "sets/s", "cores": cores, "kind": "port",
                   "sample": f"first {sample} of the {N_SETS} sets, C oracle (oracle/bls12_381.c), {cores} threads, {dt:.1f} s"}
        t0 = time.perf_counter()
        want, _ = O.beacon_state_root_deneb(ssz)
        dt = time.perf_counter() - t0
        assert want == root0, "CPU oracle root differs from the GPU root"
        cpu_state = {"value": 1.0 / dt, "unit": "roots/s", "cores": cores, "kind": "port",
                     "sample": f"1 full {N_VALIDATORS_STATE}-validator state, C oracle with SHA-NI, {cores} threads, {dt*1e3:.1f} ms"}
        O.set_threads(1)

    if rank == 0:
        peak, peak_src = measured_peaks()
        dom = sorted(x for x in dom_ms if x > 0)
        dom_bls = dom[len(dom) // 2] if dom else None
        dom_st = st_dom[0] if st_dom and st_dom[0] > 0 else None
        bls_ach = BLS_BYTES_PER_SET * N_SETS / (dom_bls * 1e-3) / 1e9 if dom_bls else None
        st_units_dom = 8 * N_VALIDATORS_STATE                       # k_validator_roots: 8 hash32_concat per validator
        st_ach = STATE_UNIT_BYTES * st_units_dom / (dom_st * 1e-3) / 1e9 if dom_st else None
        line = {
            "metric": "bls_sig_sets_verified_per_sec", "value": bls_value, "unit": "sets/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": bls_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (381-bit Montgomery integers)", "data": "synthetic",
            "config": {"workload": f"verify_signature_sets: {N_SETS} aggregate attestations x {KEYS_PER_SET} pubkeys per GPU "
                                   f"(BASELINE configs[2]), mainnet spec, interop keys over {N_VALIDATORS_BLS} validators, "
                                   "every message distinct", "sets_per_gpu": N_SETS, "keys_per_set": KEYS_PER_SET,
                       "l2": f"inputs per step {ab.input_bytes/1e6:.0f} MB > 126 MB L2 (no flush needed)",
                       "collective": "ncclAllReduce(min) of 1 x int32 verdict per step" if world > 1 else "none (1 GPU)"},
            "clocks": clocks,
            "e2e": {"value": N_SETS * world / (bls_e2e_ms / 1e3), "unit": "sets/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": 1, "ms_per_step": bls_e2e_ms, "timer": "perf_counter around synchronised steps"},
            "e2e_indexed": {"value": N_SETS * world / (bls_e2e_idx_ms / 1e3), "unit": "sets/s",
                            "h2d_bytes_per_step": int(h2d_idx), "d2h_bytes_per_step": 1, "ms_per_step": bls_e2e_idx_ms,
                            "note": "keys referenced by u32 index into the device-resident pubkey table (ValidatorPubkeyCache mirror)"},
            "gpu_launches": int(bls_launches),
            "roofline": {"kernel": "k_miller_multi", "bound": "hbm", "achieved": bls_ach, "peak": peak, "unit": "GB/s",
                         "frac": (bls_ach / peak) if bls_ach else None, "traffic": profile_traffic("k_miller_multi"),
                         "peak_source": peak_src, "kernel_ms": dom_bls,
                         "note": "integer-ALU bound by construction (SURVEY §8d): see alu_frac in DESIGN.md / profiles/"},
            "alu": {"pipe": "fmaheavy (IMAD.WIDE.U32: 1 warp-instruction / cycle / SM, measured, DESIGN.md §2.2)",
                    "fp_mul_per_set": 20400, "imad_per_fp_mul": 302,
                    "achieved_gmul_s": 20400 * bls_value / world / 1e9,
                    "peak_gmul_s": 148 * 32 * (clocks.get("sm_mhz") or 1965) * 1e6 / 302 / 1e9,
                    "frac": (20400 * bls_value / world) / (148 * 32 * (clocks.get("sm_mhz") or 1965) * 1e6 / 302)},
            "cpu_baseline": cpu_bls,
            "tree_hash": {
                "metric": "beacon_state_tree_hash_root_per_sec", "value": world / (st_ms / 1e3), "unit": "roots/s",
                "ms_per_step": st_ms, "scaling": "weak (one independent state per GPU)",
                "config": {"workload": f"cold tree_hash_root of a synthetic {N_VALIDATORS_STATE}-validator Deneb BeaconState "
                                       "(BASELINE configs[1]), resident in HBM", "hash32_concat_units": int(units),
                           "l2": "state (72 MB) fits L2; k_init-free re-hash each step reads the same buffers "
                                 "(ALU-bound kernel, HBM traffic is 2% of time)"},
                "e2e": {"value": world / (st_e2e_ms / 1e3), "unit": "roots/s", "h2d_bytes_per_step": len(ssz),
                        "d2h_bytes_per_step": 32, "ms_per_step": st_e2e_ms},
                "gpu_launches": int(st_launches),
                "roofline": {"kernel": "k_validator_roots", "bound": "hbm", "achieved": st_ach, "peak": peak, "unit": "GB/s",
                             "frac": (st_ach / peak) if st_ach else None, "traffic": profile_traffic("k_validator_roots"),
                             "peak_source": peak_src, "kernel_ms": dom_st},
                "cpu_baseline": cpu_state,
                "sharded_single_state": sharded,
            },
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU implementation of the path, timed on the host cores.  The reference (Rust + blst/sha2
    asm) cannot be built in this image, so this is the CPU oracle ("port"), all host threads, bounded sample."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import numpy as np
    from tests import oracle_lib as O
    from oracle import bls_ref as B
    cores = O.hw_threads()
    O.set_threads(cores)
    K, W = args.steps, args.warmup
    # bounded BLS sample built with the oracle's own signer (no GPU in this arm)
    sample = max(64, 4 * cores)
    keys_unique = 256
    sks = [B.interop_secret_key(i) for i in range(keys_unique)]
    pk_tab = [O.bls_sk_to_pk(s.to_bytes(32, "big")) for s in sks]
    rng = np.random.default_rng(1)
    import hashlib
    sigs, msgs, pks = [], [], []
    base = [rng.permutation(keys_unique)[:KEYS_PER_SET] for _ in range(8)]
    for j in range(sample):
        ids = base[j % 8]
        m = hashlib.sha256(b"ref-arm" + j.to_bytes(8, "little")).digest()
        agg = sum(sks[i] for i in ids) % B.R
        sigs.append(O.bls_sign(agg.to_bytes(32, "big"), m))
        msgs.append(m)
        pks.append(b"".join(pk_tab[i] for i in ids))
    sigs, msgs, pks = b"".join(sigs), b"".join(msgs), b"".join(pks)
    offs = np.arange(sample + 1, dtype=np.uint32) * KEYS_PER_SET
    rands = rng.integers(1, 2 ** 63, size=sample, dtype=np.uint64) * 2 + 1
    for _ in range(min(W, 1)):
        assert O.bls_verify_signature_sets(sigs, msgs, pks, offs, rands)
    t0 = time.perf_counter()
    for _ in range(K):
        assert O.bls_verify_signature_sets(sigs, msgs, pks, offs, rands)
    dt = (time.perf_counter() - t0) / K
    value = sample / dt
    from lighthouse_b200.synthetic import beacon_state_deneb_ssz   # pure numpy byte layout (loads no GPU code path)
    ssz = beacon_state_deneb_ssz(N_VALIDATORS_STATE, seed=42)
    O.beacon_state_root_deneb(ssz)
    t0 = time.perf_counter()
    for _ in range(K):
        O.beacon_state_root_deneb(ssz)
    dts = (time.perf_counter() - t0) / K
    desc = f"{sample} sets x {KEYS_PER_SET} keys per step (bounded sample of the {N_SETS}-set workload), C oracle, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "bls_sig_sets_verified_per_sec", "value": value, "unit": "sets/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (381-bit Montgomery integers)", "data": "synthetic",
        "config": {"workload": f"verify_signature_sets: {N_SETS} aggregate attestations x {KEYS_PER_SET} pubkeys per GPU "
                               "(BASELINE configs[2]) — CPU arm runs a bounded sample", "sample": desc},
        "cpu_baseline": {"value": value, "unit": "sets/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "sets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "tree_hash": {"metric": "beacon_state_tree_hash_root_per_sec", "value": 1.0 / dts, "unit": "roots/s",
                      "ms_per_step": dts * 1e3, "cpu_baseline": {"value": 1.0 / dts, "unit": "roots/s", "cores": cores,
                                                                   "kind": "port", "sample": "full 500k-validator state per step, SHA-NI"}},
        "note": "reference (Rust, blst 0.3.12 / sha2 asm) is unbuildable in this image (no cargo/rustc, crates not vendored); "
                "this is the from-spec C restatement in oracle/ ('port')",
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
