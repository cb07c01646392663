"""Regenerates the fixtures of the row-(f) pieces (run from the repo root: python tests/golden/make_golden_late.py).

pass_ingest.json   pass URIs (the reference's example pass, synthetic passes, malformed variants) -> status, fitted
                   ToBeSigned, true length and a SHA-256 of the marshalled inputs, from oracle/pass_ingest.py (the
                   restatement of test/helpers/nzcp.js with JavaScript's number semantics)
verify.json        for the committed plonk_<name>.json proofs: X_2 = [tau]_2 of the fixtures' SRS, mutated proofs and
                   the verdict of oracle/plonk.py verify (pairing form) on each
The GPU tests replay them WITHOUT the oracle in the loop.
"""
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pairing as pg  # noqa: E402
from oracle import pass_ingest as pi  # noqa: E402
from oracle import plonk as op  # noqa: E402
from tests.pass_cases import cases  # noqa: E402
from tests.test_verify import FIXTURES, TAU, _fixture, _mutations  # noqa: E402


def ingest_fixture():
    out = []
    for label, uri in cases(n_synth=12):
        for max_len in (314, 351):
            d20 = bytes((11 * k + len(label)) & 255 for k in range(20))
            st, fitted, n, vals = pi.ingest(uri, max_len, d20)
            out.append({"label": label, "uri": uri, "max_len": max_len, "data": d20.hex(), "status": st,
                        "tobesigned_fitted_sha256": hashlib.sha256(fitted).hexdigest(), "tobesigned_len": n,
                        "inputs_sha256": hashlib.sha256(b"".join(v.to_bytes(32, "little") for v in vals)).hexdigest()})
    return {"generator": "tests/golden/make_golden_late.py: oracle/pass_ingest.py ingest()", "cases": out}


def verify_fixture():
    x2 = pg.g2_mul(pg.G2_GEN, TAU)
    res = {"generator": "tests/golden/make_golden_late.py: oracle/plonk.py verify (pairing form)",
           "X_2": [[str(x2[0][0]), str(x2[0][1])], [str(x2[1][0]), str(x2[1][1])], ["1", "0"]], "fixtures": {}}
    for name in FIXTURES:
        fx, zkey, vk, proof, pub = _fixture(name)
        rows = []
        for label, pr, pu, expected in _mutations(proof, pub, random.Random(2)):
            assert op.verify(vk, pu, op.proof_from_bytes(pr)) == expected
            rows.append({"label": label, "proof_hex": pr.hex(), "public_signals": [str(x) for x in pu], "valid": expected})
        res["fixtures"][name] = rows
    return res


if __name__ == "__main__":
    with open(os.path.join(HERE, "pass_ingest.json"), "w") as f:
        json.dump(ingest_fixture(), f, indent=1)
    with open(os.path.join(HERE, "verify.json"), "w") as f:
        json.dump(verify_fixture(), f, indent=1)
    print("wrote pass_ingest.json, verify.json")
