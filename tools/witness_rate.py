import os, sys, time
sys.path.insert(0, '/root/repo')
from nzcb_circom_b200 import Context, nzcp_helpers as H
from nzcb_circom_b200.circom_tester import wasm_tester
c = Context(0)
cir = wasm_tester("nzcp_live", c)
base=[]
for i in range(64):
    p = H.synth_pass(i)
    base.append(cir.compiled.flatten_input(H.nzcp_input(p["toBeSigned"], 351, p["data"])))
for B in [int(x) for x in os.environ.get("WR_B", "1,148,444,2048,8192").split(",")]:
    inputs=[base[i%64] for i in range(B)]
    for it in range(2):
        raw, st = cir.calculateWitnessBatch(inputs, True, c, want_witness=False)
    print(f"B={B}: device {c.last_device_ms:.2f} ms  {B/c.last_device_ms*1e3:.0f} passes/s ok={all(s==0 for s in st)}", flush=True)
