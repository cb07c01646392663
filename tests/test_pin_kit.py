"""The snarkjs pin kit (baseline/snarkjs_baseline.mjs + tools/export_fixture.py): the fixture
set it consumes is complete and self-consistent -- the exported zkey carries X_2 = [tau]_2, the exported
verification_key.json is the zkey's, and the oracle's pairing verifier accepts the exported proof against it.  Node is
absent here, so the script itself stays unexecuted and bench.py's reference arm must fall back to the C port."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_export_fixture_from_golden(tmp_path):
    from oracle import pairing as pg
    from oracle import plonk as op
    from oracle.keccak import hash_to_fr

    out = tmp_path / "pin"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "export_fixture.py"), "--golden", "tiny", "-o", str(out)], check=True)
    need = {"circuit.zkey", "witness.wtns", "blinders.json", "proof.json", "public.json", "verification_key.json", "calldata.txt"}
    assert need <= set(os.listdir(out))
    with open(os.path.join(ROOT, "baseline", "snarkjs_baseline.mjs")) as f:
        script = f.read()
    assert all(name in script for name in need - {"calldata.txt"})   # the script reads exactly what the exporter writes
    zkey = (out / "circuit.zkey").read_bytes()
    vk_zkey = op.verification_key(zkey)
    assert pg.g2_from_lem(bytes(vk_zkey["X_2"])) == pg.g2_mul(pg.G2_GEN, hash_to_fr(b"nzcb-b200-tau"))
    vkj = json.loads((out / "verification_key.json").read_text())
    vk = op.vk_from_json(vkj)
    assert op.vk_x2(vk) == op.vk_x2(vk_zkey) and vk["Qm"] == vk_zkey["Qm"] and vk["S3"] == vk_zkey["S3"]
    proof = json.loads((out / "proof.json").read_text())
    pub = [int(x) for x in json.loads((out / "public.json").read_text())]
    assert len(json.loads((out / "blinders.json").read_text())) == 9
    pr = {k: (None if v[2] == "0" else (int(v[0]), int(v[1]))) if isinstance(v, list) else int(v)
          for k, v in proof.items() if k not in ("protocol", "curve")}
    assert op.verify(vk, pub, pr)                     # the real pairing check, as snarkjs.plonk.verify would run it
    pr["eval_a"] = (pr["eval_a"] + 1) % (1 << 200)
    assert not op.verify(vk, pub, pr)


def test_reference_arm_prefers_snarkjs_only_when_node_exists():
    import shutil

    sys.path.insert(0, ROOT)
    import bench

    if shutil.which("node") is None:
        assert bench._snarkjs_available() is None      # -> kind "port" (the C restatement), as DESIGN.md says
