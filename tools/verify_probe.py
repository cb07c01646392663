"""GPU probe: prove a batch of distinct nzcp_live passes, verify each on the device (warp kernel and serial kernel)
and with the oracle's pairing verifier; report disagreements."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nzcb_circom_b200 import Context, nzcp_helpers as H
from nzcb_circom_b200.prover import NzcpProver, default_tau
from oracle import plonk as op

ctx = Context(0)
pr = NzcpProver(live=True, tau=default_tau(), ctx=ctx)
pr.setup()
passes = [H.synth_pass(1000 + i) for i in range(32)]
res = pr.prove_passes([(p["toBeSigned"], p["data"]) for p in passes])
assert all(s == 0 for _, _, s in res)
pubs = [[int(x) for x in r[1]] for r in res]
prfs = [r[0] for r in res]
warp = pr.verify(pubs, prfs)
warp2 = pr.verify(pubs, prfs)
single = [pr.verify(pubs[i:i + 1], prfs[i:i + 1])[0] for i in range(32)]
os.environ["NZCB_VERIFY_SERIAL"] = "1"
t0 = time.time()
serial = pr.verify(pubs, prfs)
print("serial kernel 32 proofs: %.1f ms" % ctx.last_device_ms)
os.environ["NZCB_VERIFY_SERIAL"] = "0"
vk = op.vk_from_json(pr.vk)
orc = [op.verify(vk, pubs[i], op.proof_from_bytes(prfs[i])) for i in range(32)]
print("warp  ", "".join("1" if x else "0" for x in warp))
print("warp2 ", "".join("1" if x else "0" for x in warp2))
print("single", "".join("1" if x else "0" for x in single))
print("serial", "".join("1" if x else "0" for x in serial))
print("oracle", "".join("1" if x else "0" for x in orc))
for i in range(32):
    if not warp[i]:
        print("failing proof", i, prfs[i].hex()[:64], pubs[i])
        open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", f"fail_{i}.bin"), "wb").write(prfs[i] + b"".join(x.to_bytes(32, "little") for x in pubs[i]))
        break
import json
json.dump(pr.vk, open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "vk_live.json"), "w"))
