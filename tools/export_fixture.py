"""Writes the file set baseline/snarkjs_baseline.mjs consumes (the snarkjs pin kit, baseline/README.md):

    circuit.zkey  witness.wtns  blinders.json  proof.json  public.json  verification_key.json  calldata.txt
    (+ circuit.r1cs, pot.ptau for the small cases, so `snarkjs plonk setup circuit.r1cs pot.ptau x.zkey` can be
    replayed and its zkey compared with circuit.zkey byte for byte)

  python tools/export_fixture.py --golden small -o DIR           CPU only: the committed tests/golden fixture
  python tools/export_fixture.py --circuit skipValue5_test -o DIR   GPU: powersoftau + plonk setup + fullProve here
  python tools/export_fixture.py --circuit nzcp_live -o DIR         GPU, full size (3.7 GB zkey)

Everything is produced by THIS repository (product path on the GPU, or the committed oracle fixtures); the script
never reads /root/reference.  X_2 = [tau]_2 of the synthetic SRS (tau = keccak("nzcb-b200-tau") mod r) is written
into the zkey header, so the exported key verifies under snarkjs.plonk.verify."""
import argparse
import base64
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _patch_x2(zkey: bytes, x2_lem: bytes) -> bytes:
    """X_2 sits at offset 668 of section 2; the golden fixtures were written with zeros there"""
    import struct
    nsec = struct.unpack_from("<I", zkey, 8)[0]
    pos = 12
    for _ in range(nsec):
        sid, size = struct.unpack_from("<IQ", zkey, pos)
        pos += 12
        if sid == 2:
            off = pos + 156 + 512
            return zkey[:off] + x2_lem + zkey[off + 128:]
        pos += size
    raise ValueError("zkey file: no header section")


def _write(dirname, name, data):
    mode = "wb" if isinstance(data, (bytes, bytearray)) else "w"
    with open(os.path.join(dirname, name), mode) as f:
        f.write(data)


def export_golden(name, out):
    from nzcb_circom_b200.snarkjs import plonk, zKey
    from oracle import pairing as pg
    from oracle.keccak import hash_to_fr

    with open(os.path.join(ROOT, "tests", "golden", f"plonk_{name}.json")) as f:
        fx = json.load(f)
    tau = hash_to_fr(b"nzcb-b200-tau")
    zkey = _patch_x2(base64.b64decode(fx["zkey_b64"]), pg.g2_to_lem(pg.g2_mul(pg.G2_GEN, tau)))
    _write(out, "circuit.zkey", zkey)
    _write(out, "witness.wtns", base64.b64decode(fx["wtns_b64"]))
    _write(out, "blinders.json", json.dumps(fx["blinders"]))
    _write(out, "proof.json", fx["proof_json"])
    _write(out, "public.json", json.dumps(fx["public_signals"]))
    _write(out, "verification_key.json", json.dumps(zKey.exportVerificationKey(zkey), indent=1))
    _write(out, "calldata.txt", plonk.exportSolidityCallData(json.loads(fx["proof_json"]), fx["public_signals"]))
    _write(out, "PROVENANCE.txt", f"tests/golden/plonk_{name}.json ({fx['generator']}); X_2 patched to [tau]_2; domain {fx['domain']}\n")


def export_circuit(name, out, seed):
    import random

    from nzcb_circom_b200 import Context
    from nzcb_circom_b200 import nzcp_helpers as H
    from nzcb_circom_b200.prover import CircuitProver, NzcpProver, default_tau
    from nzcb_circom_b200.snarkjs import R_MOD, plonk, write_wtns, wtns_from_raw

    ctx = Context(0)
    if name in ("nzcp_live", "nzcp_example"):
        pr = NzcpProver(live=(name == "nzcp_live"), tau=default_tau(), ctx=ctx)
        p = H.synth_pass(seed, live=(name == "nzcp_live"))
        inp = H.nzcp_input(p["toBeSigned"], pr.max_len, p["data"])
    else:
        pr = CircuitProver(name, default_tau(), ctx)
        if name != "skipValue5_test":
            raise SystemExit("--circuit: nzcp_live, nzcp_example or skipValue5_test (add an input for other wrappers here)")
        inp = {"bytes": [0x83, 23, 23, 23, 0], "pos": 0}
    zkey = pr.setup(keep_zkey=True)
    rng = random.Random(seed)
    bl = [rng.randrange(R_MOD) for _ in range(9)]
    raw, st = pr.tester.calculateWitnessBatch([inp], True, ctx)
    assert st == [0], st
    wt = wtns_from_raw(raw)
    proof, public = plonk.prove(pr.zk, wt, blinders=bl, raw=True)
    _write(out, "circuit.zkey", bytes(zkey))
    _write(out, "witness.wtns", bytes(wt))
    _write(out, "circuit.r1cs", bytes(pr.art.r1cs_bytes()))
    _write(out, "input.json", json.dumps(inp))
    _write(out, "blinders.json", json.dumps([str(x) for x in bl]))
    pj = plonk.proof_json(proof, ctx)
    _write(out, "proof.json", pj)
    _write(out, "public.json", json.dumps([str(int(x)) for x in public]))
    _write(out, "verification_key.json", json.dumps(pr.vk, indent=1))
    _write(out, "calldata.txt", plonk.exportSolidityCallData(json.loads(pj), [str(int(x)) for x in public]))
    assert pr.verify([[int(x) for x in public]], [proof]) == [True]
    _write(out, "PROVENANCE.txt", f"circuit {name}, seed {seed}, synthetic SRS tau = keccak('nzcb-b200-tau') mod r, domain 2^{pr.power}; "
                                  f"witness and proof from libnzcb.so on {ctx.device_name if hasattr(ctx, 'device_name') else 'cuda:0'}\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--golden", help="tiny | small | nopublic (tests/golden, CPU only)")
    ap.add_argument("--circuit", help="nzcp_live | nzcp_example | skipValue5_test (needs a GPU)")
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("-o", "--out", required=True)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    if a.golden:
        export_golden(a.golden, a.out)
    elif a.circuit:
        export_circuit(a.circuit, a.out, a.seed)
    else:
        ap.error("--golden or --circuit")
    print("wrote", sorted(os.listdir(a.out)))


if __name__ == "__main__":
    main()
