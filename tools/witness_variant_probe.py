"""GPU check of the slim witness program (NZCB_WITNESS_DROP_IMPLIED=1): witnesses against the C oracle on the default
program, a rejected pass, and the rate at a few batch sizes."""
import os
import sys

os.environ["NZCB_WITNESS_DROP_IMPLIED"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nzcb_circom_b200 import Context, nzcp_helpers as H  # noqa: E402
from nzcb_circom_b200.circom_tester import compile_circuit, wasm_tester  # noqa: E402
from oracle import c_oracle as C  # noqa: E402

c = Context(0)
cir = wasm_tester("nzcp_live", c)
slim = cir.compiled
os.environ["NZCB_WITNESS_DROP_IMPLIED"] = "0"
full = compile_circuit("nzcp_live")
print("instructions", slim.n_instr, "vs", full.n_instr)
base = []
for i in range(64):
    p = H.synth_pass(i)
    base.append(slim.flatten_input(H.nzcp_input(p["toBeSigned"], 351, p["data"])))
bad = bytearray(H.synth_pass(1)["toBeSigned"])
bad[30] = 0x65
inputs = base[:3] + [slim.flatten_input(H.nzcp_input(bytes(bad), 351, bytes(20)))]
raw, st = cir.calculateWitnessBatch(inputs, True, c)
nw = slim.n_witness * 32
ok = st == [0, 0, 0, -6]
for i in range(3):
    inp = b"".join(int(v).to_bytes(32, "little") for v in inputs[i])
    rc, w = C.witness(full.wprog_bytes(), inp, full.n_total)
    ok = ok and rc == 0 and raw[i * nw:(i + 1) * nw] == w[:nw]
print("witness == oracle(default program), rejection kept:", ok, st)
for B in (1, 148, 2048):
    for _ in range(2):
        _, s = cir.calculateWitnessBatch([base[i % 64] for i in range(B)], True, c, want_witness=False)
    print(f"B={B}: {c.last_device_ms:.2f} ms  {B / c.last_device_ms * 1e3:.0f} passes/s ok={all(x == 0 for x in s)}", flush=True)
