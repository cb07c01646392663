"""Config 2 / config 5 at full size: nzcp_live PLONK proofs (domain 2^21) on the GPU.  The Python
oracle cannot prove at this size, so parity is carried by size-independent properties: every proof
is accepted by the oracle's verifier (known-trapdoor form of plonk.verify), the public signals are
the hashlib expectations, proofs are deterministic under injected blinders, the .wtns path and the
fused path give identical bytes, and a rejected pass does not poison its batch."""
import hashlib
import random

import pytest

from nzcb_circom_b200 import nzcp_helpers as H
from oracle import bn254 as b
from oracle import plonk as oplonk
from oracle.keccak import hash_to_fr

pytestmark = pytest.mark.gpu
TAU = hash_to_fr(b"nzcb-b200-tau")


def test_default_tau_is_the_documented_trapdoor():
    from nzcb_circom_b200.prover import default_tau

    assert default_tau() == TAU


def test_nzcp_live_domain(nzcp_live_prover):
    pr = nzcp_live_prover
    assert pr.zk.domain_size == 1 << 21 and pr.zk.n_public == 3  # README.md:41 (power-21 ptau), nPublic = 3 outputs
    assert pr.vk["power"] == 21 and pr.vk["k1"] == "2" and pr.vk["k2"] == "3"


def test_nzcp_live_proofs_verify(nzcp_live_prover):
    from nzcb_circom_b200.snarkjs import plonk, wtns_from_raw

    pr = nzcp_live_prover
    vk = oplonk.vk_from_json(pr.vk)
    rng = random.Random(5)
    passes = [H.synth_pass(s) for s in (10, 11, 12)]
    bad = bytearray(passes[1]["toBeSigned"])
    bad[30] = 0x65  # claims header is not a map -> the circuit rejects this pass
    items = [(passes[0]["toBeSigned"], passes[0]["data"]), (bytes(bad), passes[1]["data"]),
             (passes[2]["toBeSigned"], passes[2]["data"])]
    blinders = [[rng.randrange(b.R_MOD) for _ in range(9)] for _ in items]
    res = pr.prove_passes(items, blinders)
    assert [s for _, _, s in res] == [0, -6, 0] and res[1][0] is None
    for (proof, public, _), p in ((res[0], passes[0]), (res[2], passes[2])):
        pub = [int(x) for x in public]
        nh, th, exp, data = H.nzcp_decode_outputs(pub)
        assert nh == hashlib.sha512(H.fitBytes(p["nullifier"].encode(), 64)).digest()[:32]
        assert th == hashlib.sha256(p["toBeSigned"]).digest() and exp == p["exp"] and data == p["data"]
        assert oplonk.verify_with_trapdoor(vk, pub, oplonk.proof_from_bytes(proof), TAU)
        tampered = oplonk.proof_from_bytes(proof)
        tampered["eval_zw"] = (tampered["eval_zw"] + 1) % b.R_MOD
        assert not oplonk.verify_with_trapdoor(vk, pub, tampered, TAU)
        assert not oplonk.verify_with_trapdoor(vk, [pub[0], pub[1], pub[2] ^ 1], oplonk.proof_from_bytes(proof), TAU)
    # deterministic under injected blinders; different blinders -> different proof, same public signals
    again = pr.prove_passes([items[0]], [blinders[0]])
    assert again[0][0] == res[0][0]
    other = pr.prove_passes([items[0]], None)
    assert other[0][0] != res[0][0] and other[0][1] == res[0][1]
    # the two API paths agree byte for byte: calculateWitness -> .wtns -> plonk.prove  vs  fused fullProve
    raw, st = pr.tester.calculateWitnessBatch([H.nzcp_input(items[0][0], 351, items[0][1])], True, pr.ctx)
    assert st == [0]
    proof2, pub2 = plonk.prove(pr.zk, wtns_from_raw(raw), blinders=blinders[0], raw=True)
    assert proof2 == res[0][0] and pub2 == res[0][1]


def test_msm_split_latency_mode_two_ranks_one_gpu(nzcp_live_prover):
    """Latency mode: two contexts prove the SAME pass, each committing half of every MSM's point
    range; the partial sums are exchanged (here: an in-process all-gather between two threads sharing the GPU;
    bench.py does it with NCCL between processes) -- both ranks must emit the single-GPU proof byte for byte."""
    import threading

    from nzcb_circom_b200 import Context
    from nzcb_circom_b200.prover import NzcpProver, default_tau

    pr0 = nzcp_live_prover
    p = H.synth_pass(21)
    item = [(p["toBeSigned"], p["data"])]
    rng = random.Random(9)
    bl = [[rng.randrange(b.R_MOD) for _ in range(9)]]
    expect = pr0.prove_passes(item, bl)[0]
    assert expect[2] == 0

    world = 2
    barrier = threading.Barrier(world)
    slots = [None] * world
    results = [None] * world
    errors = []

    def make_allgather(rank):
        def allgather(send):
            slots[rank] = send
            barrier.wait(timeout=120)
            out = b"".join(slots)
            barrier.wait(timeout=120)
            return out
        return allgather

    def run(rank, prover):
        try:
            prover.ctx.set_msm_split(rank, world, make_allgather(rank))
            results[rank] = prover.prove_passes(item, bl)[0]
        except Exception as e:  # pragma: no cover
            errors.append(e)
            barrier.abort()
        finally:
            prover.ctx.set_msm_split(0, 1, None)

    ctx1 = Context(0)
    pr1 = NzcpProver(live=True, tau=default_tau(), ctx=ctx1)
    pr1.setup()
    th = [threading.Thread(target=run, args=(0, pr0)), threading.Thread(target=run, args=(1, pr1))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    assert results[0] == expect and results[1] == expect
    pr1.zk.close()
    ctx1.close()


def test_nzcp_live_proof_bytes_equal_c_oracle_full_size(nzcp_live_prover):
    """Full size (domain 2^21), same zkey bytes, same pass, same blinders: the GPU proof and the proof of the C port
    of the oracle (snarkjs' schedule: monomial-basis commitments, T / Tz halves, two inverse transforms -- none of
    the shortcuts the GPU prover takes) are the same 800 bytes."""
    from oracle import c_oracle as C

    pr = nzcp_live_prover
    C.use_all_cores()
    p = H.synth_pass(33)
    inp = pr.marshal_passes([(p["toBeSigned"], p["data"])])
    bl = list(range(11, 20))
    rc, cproof, cpub = C.fullprove(pr.art.wprog_bytes(), inp, pr.zkey_bytes, bl, 3)
    assert rc == 0
    gproof, gpub, st = pr.prove_raw(inp, 1, [bl])[0]
    assert st == 0 and gproof == cproof and [int(x) for x in gpub] == list(cpub)


def test_nzcp_example_setup_prove_verify(ctx):
    """BASELINE.json configs[0]: the nzcp_example circuit (/root/reference/circuits/nzcp_example.circom:4,
    Makefile:54-57 `plonk setup` / `zkey export verificationkey`) through powersoftau -> plonk setup -> fullProve ->
    verify on the GPU, for the reference's own EXAMPLE_PASS_URI (test/nzcp.js:71): public signals are the golden
    outputs the reference's test reads, the proof is byte-identical to the C oracle's, the device verifier and the
    oracle's verifier accept it and reject a tampered copy."""
    import json
    import os

    from nzcb_circom_b200.prover import NzcpProver, default_tau
    from oracle import c_oracle as C

    pr = NzcpProver(live=False, tau=default_tau(), ctx=ctx)
    zkey = pr.setup(keep_zkey=True)
    assert pr.zk.n_public == 3 and pr.zk.domain_size == 1 << pr.power
    cose = H.getCOSE(H.EXAMPLE_PASS_URI)
    tbs = H.encodeToBeSigned(cose["bodyProtected"], cose["payload"])
    data = bytes(range(1, 21))
    p2 = H.synth_pass(77, live=False)
    items = [(tbs, data), (p2["toBeSigned"], p2["data"])]
    bl = [list(range(21, 30)), list(range(31, 40))]
    res = pr.prove_passes(items, bl)
    assert [r[2] for r in res] == [0, 0]
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.json")) as fh:
        golden = json.load(fh)
    pub0 = [int(x) for x in res[0][1]]
    assert pub0 == [int(x) for x in golden["public_outputs"]["value"]]
    nh, th, exp, d = H.nzcp_decode_outputs(pub0)
    assert th == hashlib.sha256(tbs).digest() and exp == 1951416330 and d == data
    assert nh == hashlib.sha512(H.fitBytes(b"Jack,Sparrow,1960-04-16", 64)).digest()[:32]
    C.use_all_cores()
    inp = pr.marshal_passes(items[:1])
    rc, cproof, cpub = C.fullprove(pr.art.wprog_bytes(), inp, zkey, bl[0], 3)
    assert rc == 0 and cproof == res[0][0] and list(cpub) == pub0
    vk = oplonk.vk_from_json(pr.vk)
    for proof, public, _ in res:
        pub = [int(x) for x in public]
        assert oplonk.verify_with_trapdoor(vk, pub, oplonk.proof_from_bytes(proof), TAU)
    bad = bytearray(res[1][0])
    bad[640] ^= 1
    pubs = [res[0][1], res[1][1], res[1][1]]
    assert pr.verify(pubs, [res[0][0], res[1][0], bytes(bad)]) == [True, True, False]
    pr.zk.close()


def _wires_digest(raw, n_witness):
    """the digest nzcb_witness_batch_ex computes on the device (include/nzcb.h), over host wires"""
    import numpy as np

    w = np.frombuffer(raw, dtype="<u4", count=n_witness * 8).astype(np.uint64)
    z = np.arange(n_witness * 8, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return int(((w + np.uint64(1)) * z).sum(dtype=np.uint64))


def test_witness_batch_65536_passes(ctx):
    """BASELINE.json configs[3] at full size: 65,536 nzcp_live passes through the batched witness program (8 calls of
    8,192), 4,096 DISTINCT synthetic passes plus corrupted ones.  Checked for every pass: the status flag (exactly
    the corrupted passes are rejected, a failed pass never fails its batch), the three public outputs against
    hashlib / the pass itself, and a device-side digest over ALL of the pass's wires -- equal for every copy of a
    pass wherever it sits in the batch, and equal to the digest of the C oracle's witness for a sample of the
    distinct passes.  For a strided sample the whole witness comes back and is compared wire for wire."""
    import ctypes

    from nzcb_circom_b200.circom_tester import wasm_tester
    from oracle import c_oracle as C

    cir = wasm_tester("nzcp_live", ctx)
    art = cir.compiled
    n_out, n_wit = art.n_out, art.n_witness
    D, B, CALLS, STRIDE = 4096, 8192, 8, 1021
    passes = [H.synth_pass(500 + s) for s in range(D)]
    distinct = []
    for p in passes:
        vals = art.flatten_input(H.nzcp_input(p["toBeSigned"], 351, p["data"]))
        distinct.append(b"".join(int(v).to_bytes(32, "little") for v in vals))
    want_pub = []
    for p in passes:
        nh = hashlib.sha512(H.fitBytes(p["nullifier"].encode(), 64)).digest()[:32]
        want_pub.append((nh, hashlib.sha256(p["toBeSigned"]).digest(), p["exp"], p["data"]))
    bad = bytearray(distinct[0])
    bad[0:32] = (2).to_bytes(32, "little")  # first ToBeSigned bit is not boolean (nzcptpl.circom:493-496)
    h = cir._handle(ctx)
    wprog = art.wprog_bytes()
    oracle_digest = {}
    pub_seen = {}
    total_ms = 0.0
    n_sample = (B + STRIDE - 1) // STRIDE
    for call in range(CALLS):
        # a different arrangement of the distinct passes in every call, so a pass meets many batch positions
        idx = [(i * (2 * call + 1) + 37 * call) % D for i in range(B)]
        bad_at = set(range(5 + call, B, 1021))
        rows = [bytes(bad) if i in bad_at else distinct[idx[i]] for i in range(B)]
        buf = b"".join(rows)
        status = (ctypes.c_int32 * B)()
        outputs = ctypes.create_string_buffer(B * n_out * 32)
        digest = (ctypes.c_uint64 * B)()
        sample = ctypes.create_string_buffer(n_sample * n_wit * 32)
        ctx.check(ctx.lib.nzcb_witness_batch_ex(ctx.h, h, buf, B, outputs, digest, STRIDE, sample, status))
        total_ms += ctx.last_device_ms
        st = list(status)
        assert all((st[i] == -6) == (i in bad_at) and st[i] in (0, -6) for i in range(B))
        first_digest = {}
        out_raw = outputs.raw
        for i in range(B):
            if i in bad_at:
                continue
            d = idx[i]
            raw_pub = out_raw[i * n_out * 32:(i + 1) * n_out * 32]
            if d not in pub_seen:    # decoded once per distinct pass; every other copy must carry the same bytes
                pub = [int.from_bytes(raw_pub[k * 32:(k + 1) * 32], "little") for k in range(n_out)]
                assert H.nzcp_decode_outputs(pub) == want_pub[d], (call, i)
                pub_seen[d] = raw_pub
            assert pub_seen[d] == raw_pub, (call, i)
            assert first_digest.setdefault(d, digest[i]) == digest[i], (call, i)   # position independent
        # anchor the digests and the sampled witnesses on the C oracle (a few passes per call: 2 s each on one core)
        for k in range(n_sample):
            i = k * STRIDE
            if i in bad_at or k % 4 != call % 4:
                continue
            d = idx[i]
            if d not in oracle_digest:
                rc, wires = C.witness(wprog, distinct[d], art.n_total)
                assert rc == 0
                oracle_digest[d] = (_wires_digest(wires, n_wit), wires[:n_wit * 32])
            assert digest[i] == oracle_digest[d][0], (call, i)
            assert bytes(memoryview(sample)[k * n_wit * 32:(k + 1) * n_wit * 32]) == oracle_digest[d][1], (call, i)
        for d, (dg, _) in oracle_digest.items():
            if d in first_digest:
                assert first_digest[d] == dg
    assert len(oracle_digest) >= 8 and len(pub_seen) == D
    print(f"65,536 passes in {total_ms:.0f} ms of device time ({65536 / total_ms * 1e3:.0f} passes/s), "
          f"{len(oracle_digest)} witnesses compared wire for wire with the C oracle")


def test_witness_batch_device_resident_inputs_equal_host_inputs(ctx):
    """nzcb_witness_batch_ex_dev (inputs already in HBM: bench.py's roofline_witness) against nzcb_witness_batch_ex
    (host inputs; more than one launch, so the staged upload runs): same status, public outputs and all-wire digests"""
    import ctypes

    from nzcb_circom_b200.circom_tester import wasm_tester

    cir = wasm_tester("nzcp_live", ctx)
    art = cir.compiled
    B = 1000  # > one wave of 888 passes: two launches
    rows = []
    for s_ in range(40):
        p = H.synth_pass(900 + s_)
        vals = art.flatten_input(H.nzcp_input(p["toBeSigned"], 351, p["data"]))
        rows.append(b"".join(int(v).to_bytes(32, "little") for v in vals))
    bad = bytearray(rows[3])
    bad[0:32] = (2).to_bytes(32, "little")
    buf = b"".join(bytes(bad) if i == 777 else rows[i % 40] for i in range(B))
    h = cir._handle(ctx)
    res = []
    d = ctx.dev_alloc(len(buf))
    ctx.dev_upload(d, buf)
    for dev in (False, True):
        status = (ctypes.c_int32 * B)()
        outputs = ctypes.create_string_buffer(B * art.n_out * 32)
        digest = (ctypes.c_uint64 * B)()
        fn = ctx.lib.nzcb_witness_batch_ex_dev if dev else ctx.lib.nzcb_witness_batch_ex
        ctx.check(fn(ctx.h, h, d if dev else buf, B, outputs, digest, 0, None, status))
        res.append((list(status), outputs.raw, list(digest)))
    ctx.dev_free(d)
    assert res[0][0] == res[1][0] and [i for i, s_ in enumerate(res[0][0]) if s_ != 0] == [777]
    keep = [i for i in range(B) if i != 777]
    assert all(res[0][2][i] == res[1][2][i] for i in keep) and all(res[0][2][i] == res[0][2][i % 40] for i in keep if i % 40 != 777 % 40 or True)
    n = art.n_out * 32
    assert all(res[0][1][i * n:(i + 1) * n] == res[1][1][i * n:(i + 1) * n] for i in keep)
