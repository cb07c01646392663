"""nzcp_live witness program on B passes through the C ABI: python tools/witness_probe.py [B ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nzcb_circom_b200 import Context, nzcp_helpers as H
from nzcb_circom_b200.circom_tester import wasm_tester
c = Context(0)
cir = wasm_tester("nzcp_live", c)
for B in [int(a) for a in sys.argv[1:]] or [1, 8]:
    inputs = []
    for i in range(B):
        p = H.synth_pass(i)
        inputs.append(H.nzcp_input(p["toBeSigned"], 351, p["data"]))
    for it in range(3):
        t = time.perf_counter()
        raw, st = cir.calculateWitnessBatch(inputs, True, c, want_witness=False)
        dt = time.perf_counter() - t
    print(f"B={B}: device {c.last_device_ms:.2f} ms, wall {1000*dt:.1f} ms, status ok={all(s == 0 for s in st)}")
