"""Dense fixed-base MSM batch (K jobs of 2^log_n uniform scalars, device resident) with the batched-affine halving
rounds switched by NZCB_MSM_AFFINE: results must be identical, device times side by side.
python tools/msm_affine_probe.py [log_n] [K] [reps] [modes, e.g. 0,3]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nzcb_circom_b200 import Context
from nzcb_circom_b200.ffjavascript import G1Table
from nzcb_circom_b200.snarkjs import powersoftau

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 21
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["0", "3"]
c = Context(0)
n = 1 << log_n
srs = powersoftau.new_g1(0xae4711c826850d09ad8857707a9efce27474fb4937e510dc529a1baf89b6f59, n, c)
tab = G1Table(srs, c)
rng = np.random.default_rng(1)
ds = []
for k in range(K):
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    raw[:, 31] &= 0x1F
    d = c.dev_alloc(n * 32)
    c.dev_upload(d, raw.tobytes())
    ds.append(d)
arr = (ctypes.c_void_p * K)(*[d.value for d in ds])
ns = (ctypes.c_size_t * K)(*([n] * K))
res = {}
for mode in modes:
    os.environ["NZCB_MSM_AFFINE"] = mode
    out = (ctypes.c_uint8 * (64 * K))()
    ts = []
    for it in range(reps):
        c.check(c.lib.nzcb_msm_g1_table_dev(c.h, tab.h, arr, ns, K, out))
        ts.append(c.last_device_ms)
    res[mode] = bytes(out)
    print(f"log_n {log_n} K {K} NZCB_MSM_AFFINE={mode}: device ms {['%.3f' % t for t in ts]}", flush=True)
vals = list(res.values())
assert all(v == vals[0] for v in vals), "results differ between modes"
print("results identical across modes")
