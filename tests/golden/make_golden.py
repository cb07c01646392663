"""Regenerates the fixtures of this directory (run from the repo root: python tests/golden/make_golden.py).

reference_vectors.json  what the reference's own tests pin for the path (values copied with their file:line; the
                        derived ones are recomputed here with hashlib / the host helpers and must agree)
plonk_<name>.json       zkey + wtns + blinders -> proof bytes / proof.json / publicSignals from the Python oracle
                        (oracle/plonk.py, the restatement of snarkjs 0.4.12 plonk.prove), for circuits small enough
                        for big-int Python.  The GPU tests replay them WITHOUT the oracle in the loop.
"""
import base64
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from nzcb_circom_b200 import nzcp_helpers as H  # noqa: E402
from oracle import bn254 as b  # noqa: E402
from oracle import plonk as oplonk  # noqa: E402
from oracle.binfile import write_wtns  # noqa: E402
from oracle.keccak import hash_to_fr, keccak256  # noqa: E402
from tests.circuits_util import random_circuit  # noqa: E402


def reference_vectors():
    cose = H.getCOSE(H.EXAMPLE_PASS_URI)
    tbs = H.encodeToBeSigned(cose["bodyProtected"], cose["payload"])
    nullifier = "Jack,Sparrow,1960-04-16"
    data = bytes(range(1, 21))
    h256 = hashlib.sha256(tbs).digest()
    h512 = hashlib.sha512(H.fitBytes(nullifier.encode(), 64)).digest()
    exp = 1951416330
    out = [int.from_bytes(h512[0:31], "big"), int.from_bytes(h512[31:32] + h256[0:30], "big"),
           int.from_bytes(h256[30:32] + exp.to_bytes(4, "big") + data + bytes(5), "big")]
    return {
        "_source": "noway/nzcb-circom test vectors; see the `cite` fields (paths relative to the reference root)",
        "example_pass_uri": {"value": H.EXAMPLE_PASS_URI, "cite": "test/nzcp.js:71"},
        "example_tobesigned_max": {"value": 314, "cite": "test/nzcp.js:11"},
        "tobesigned_hex": {"value": tbs.hex(), "cite": "test/helpers/nzcp.js:180-206 applied to the example pass"},
        "tobesigned_sha256": {"value": "271ce33d671a2d3b816d788135f4343e14bc66802f8cd841faac939e8c11f3ee", "cite": "test/utils.js:17"},
        "utils_kat_chunks": {"value": ["366677313775235426412199931337625106565467678080892143469223808086055532772", "119"],
                             "cite": "test/utils.js:6-20"},
        "positions": {"claims_first_key": 28, "vc_pos": 76, "cred_subj_map": 246, "cred_subj_first_key": 247,
                      "cite": "test/nzcp.js:108,165,240"},
        "exp": {"value": exp, "cite": "test/nzcp.js:26,66"},
        "nullifier": {"value": nullifier, "cite": "test/nzcp.js:18,21-22"},
        "nullifier_sha512_first32": {"value": h512[:32].hex(), "cite": "test/nzcp.js:22-23,62"},
        "data": {"value": data.hex(), "cite": "test/nzcp.js:36-37"},
        "public_outputs": {"value": [str(x) for x in out], "cite": "circuits/nzcptpl.circom:586-654 (packing), test/nzcp.js:44-68"},
        "keccak256_empty": {"value": keccak256(b"").hex(), "cite": "js-sha3 0.8.0 known answer (SURVEY.md 0.2)"},
    }


def plonk_fixture(name, seed, **kw):
    tau = hash_to_fr(b"nzcb-b200-tau")
    r, w = random_circuit(seed, **kw)
    gates, _, _ = oplonk.r1cs_to_plonk(r)
    n = 1 << max(3, (len(gates) - 1).bit_length())
    srs = b.srs_g1(tau, n + 6)
    zkey = bytes(oplonk.setup(r, srs))
    rng = random.Random(1000 + seed)
    blinders = [rng.randrange(b.R_MOD) for _ in range(9)]
    wt = write_wtns(w)
    proof, pub, _ = oplonk.prove(zkey, wt, blinders)
    vk = oplonk.verification_key(zkey)
    assert oplonk.verify_with_trapdoor(vk, list(pub), proof, tau)
    return {
        "name": name, "domain": n, "n_public": r.n_public,
        "generator": f"tests/golden/make_golden.py: random_circuit(seed={seed}, {kw}); oracle/plonk.py setup + prove",
        "zkey_b64": base64.b64encode(zkey).decode(), "zkey_sha256": hashlib.sha256(zkey).hexdigest(),
        "wtns_b64": base64.b64encode(wt).decode(),
        "blinders": [str(x) for x in blinders],
        "proof_hex": oplonk.proof_to_bytes(proof).hex(),
        "proof_json": oplonk.proof_to_json(proof),
        "public_signals": [str(x) for x in pub],
    }


def main():
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(reference_vectors(), f, indent=1)
    for name, seed, kw in (("tiny", 1, dict(n_out=1, n_in=2, n_mul=3)), ("small", 2, dict(n_out=2, n_in=3, n_mul=20)),
                           ("nopublic", 4, dict(n_out=0, n_in=3, n_mul=25))):
        with open(os.path.join(HERE, f"plonk_{name}.json"), "w") as f:
            json.dump(plonk_fixture(name, seed, **kw), f, indent=1)
        print("wrote", name)


if __name__ == "__main__":
    main()
