"""One pass over the kernels added late in round 1, for ncu: pass ingest (4,096 URIs), the batch variant of the
witness kernel (444 passes), plonk.verify (256 proofs of the committed `small` fixture)."""
import base64
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nzcb_circom_b200 import Context, nzcp_helpers as H  # noqa: E402
from nzcb_circom_b200.circom_tester import wasm_tester  # noqa: E402
from nzcb_circom_b200.pass_ingest import toBeSignedBatch  # noqa: E402
from nzcb_circom_b200.prover import default_tau  # noqa: E402
from nzcb_circom_b200.snarkjs import VKey, plonk, powersoftau, zKey  # noqa: E402

c = Context(0)
passes = [H.synth_pass(i) for i in range(64)]
uris = [passes[i % 64]["uri"] for i in range(4096)]
datas = [passes[i % 64]["data"] for i in range(4096)]
for _ in range(2):
    res, inputs = toBeSignedBatch(uris, 351, datas, ctx=c, want_inputs=True)
print(f"ingest 4096 passes: {c.last_device_ms:.2f} ms incl. copies ({4096 / c.last_device_ms * 1e3:.0f} passes/s), ok={all(r[0] == 0 for r in res)}")

VERIFY_ONLY = os.environ.get("NKP_VERIFY_ONLY") == "1"
cir = wasm_tester("nzcp_live", c) if not VERIFY_ONLY else None
B = 444
if VERIFY_ONLY:
    B = 0
n_in = cir.compiled.n_in if cir else 0
flat = [inputs[(i % 4096) * n_in * 32:((i % 4096) + 1) * n_in * 32] for i in range(B)]
vals = [[int.from_bytes(f[k:k + 32], "little") for k in range(0, len(f), 32)] for f in flat[:64]]
for _ in range(2 if B else 0):
    raw, st = cir.calculateWitnessBatch([vals[i % 64] for i in range(B)], True, c, want_witness=False)
if B:
    print(f"witness {B} passes: {c.last_device_ms:.2f} ms ({B / c.last_device_ms * 1e3:.0f} passes/s), ok={all(s == 0 for s in st)}")

fx = json.load(open(os.path.join(ROOT, "tests", "golden", "plonk_small.json")))
vkj = zKey.exportVerificationKey(base64.b64decode(fx["zkey_b64"]))
x2 = powersoftau.new_g2(default_tau(), c)
rinv = pow(1 << 256, -1, 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47)
co = [str(int.from_bytes(x2[i:i + 32], "little") * rinv % 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47) for i in range(0, 128, 32)]
vkj["X_2"] = [[co[0], co[1]], [co[2], co[3]], ["1", "0"]]
vk = VKey(vkj, c)
proof = bytes.fromhex(fx["proof_hex"])
pub = [int(x) for x in fx["public_signals"]]
for n in (1, 256, 1024, 4096, 32768):
    for _ in range(2):
        ok = plonk.verify_batch(vk, [pub] * n, [proof] * n, c)
    print(f"verify {n} proofs: {c.last_device_ms:.2f} ms ({n / c.last_device_ms * 1e3:.0f} proofs/s), all valid={all(ok)}")
for g in ("1", "4", "16"):
    os.environ["NZCB_VERIFY_GROUP"] = g
    n = 32768 if g != "1" else 4096
    for _ in range(2):
        ok = plonk.verify_batch(vk, [pub] * n, [proof] * n, c)
    print(f"verify {n} proofs, {g} per warp: {c.last_device_ms:.2f} ms ({n / c.last_device_ms * 1e3:.0f} proofs/s), all valid={all(ok)}")
