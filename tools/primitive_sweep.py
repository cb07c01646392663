"""BASELINE.json configs[2] and [3]: BN254 G1 MSM and Fr NTT at 2^16..2^22 on one B200 (device-resident inputs, CUDA
events on the ctx stream, best of 3 after a warm-up), plus the batched witness program at B = 4096 passes.
Writes gpurun_out/primitive_sweep.json; roofline fractions use the IMAD32 peak measured in the same run.
  python tools/primitive_sweep.py [max_log_n]"""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from nzcb_circom_b200 import Context, nzcp_helpers as H
from nzcb_circom_b200.circom_tester import wasm_tester
from nzcb_circom_b200.ffjavascript import G1Table
from nzcb_circom_b200.snarkjs import powersoftau

c = Context(0)
max_log = int(sys.argv[1]) if len(sys.argv) > 1 else 22
imad = c.microbench(0, 4000, 8)
fq = c.microbench(3, 1000, 8)
res = {"imad32_per_s": imad, "fq_modmul_per_s": fq, "ntt": {}, "msm_table": {}, "msm_window": {}}
tau = 0xae4711c826850d09ad8857707a9efce27474fb4937e510dc529a1baf89b6f59
srs = powersoftau.new_g1(tau, 1 << max_log, c)
rng = np.random.default_rng(1)
for log_n in range(16, max_log + 1):
    n = 1 << log_n
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    raw[:, 31] &= 0x1F
    d = c.dev_alloc(n * 32)
    c.dev_upload(d, raw.tobytes())
    ts = []
    for it in range(4):
        c.check(c.lib.nzcb_ntt_fr_dev(c.h, d, log_n, it & 1))
        ts.append(c.last_device_ms)
    t = min(ts[1:])
    alg = (n // 2) * log_n * 264.0
    res["ntt"][log_n] = {"ms": t, "butterflies_per_s": (n // 2) * log_n / t * 1e3, "roofline_frac_imad": alg / (t * 1e-3) / imad}
    # MSM, uniform scalars: fixed-base table (the prover's mode) and one-shot windows (nzcb_msm_g1_dev)
    tab = G1Table(srs[:64 * n], c)
    out = (ctypes.c_uint8 * 64)()
    arr = (ctypes.c_void_p * 1)(d.value)
    ns = (ctypes.c_size_t * 1)(n)
    ts = []
    for it in range(4):
        c.check(c.lib.nzcb_msm_g1_table_dev(c.h, tab.h, arr, ns, 1, out))
        ts.append(c.last_device_ms)
    t = min(ts[1:])
    res["msm_table"][log_n] = {"ms": t, "points_per_s": n / t * 1e3, "roofline_frac_imad": 160.0 * n * 264 / (t * 1e-3) / imad}
    tab.close()
    db = c.dev_alloc(n * 64)
    c.dev_upload(db, srs[:64 * n])
    ts = []
    for it in range(3):
        c.check(c.lib.nzcb_msm_g1_dev(c.h, db, d, n, out))
        ts.append(c.last_device_ms)
    t = min(ts[1:])
    res["msm_window"][log_n] = {"ms": t, "points_per_s": n / t * 1e3, "roofline_frac_imad": 160.0 * n * 264 / (t * 1e-3) / imad}
    c.dev_free(db)
    c.dev_free(d)
    print(log_n, res["ntt"][log_n], res["msm_table"][log_n], res["msm_window"][log_n], flush=True)

# batched witness (configs[3] is 65,536 passes; 4,096 distinct passes are timed and the rate reported)
cir = wasm_tester("nzcp_live", c)
B = 4096
inputs = []
for i in range(B):
    p = H.synth_pass(i)
    inputs.append(cir.compiled.flatten_input(H.nzcp_input(p["toBeSigned"], 351, p["data"])))
t0 = time.perf_counter()
raw, st = cir.calculateWitnessBatch(inputs, True, c, want_witness=False)
wall = time.perf_counter() - t0
res["witness_batch"] = {"passes": B, "device_ms": c.last_device_ms, "passes_per_s_device": B / c.last_device_ms * 1e3,
                        "wall_s_incl_python_marshalling": wall, "all_accepted": all(s == 0 for s in st),
                        "wires_per_pass": cir.compiled.n_total, "hbm_write_gbs": B * cir.compiled.n_total * 32 / c.last_device_ms / 1e6}
print(res["witness_batch"])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/primitive_sweep.json", "w"), indent=1)
