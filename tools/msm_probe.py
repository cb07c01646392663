"""one fixed-base MSM of 2^log_n uniform scalars, device resident: python tools/msm_probe.py log_n [reps]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nzcb_circom_b200 import Context
from nzcb_circom_b200.ffjavascript import G1Table
from nzcb_circom_b200.snarkjs import powersoftau
c = Context(0)
log_n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 1 << log_n
srs = powersoftau.new_g1(0xae4711c826850d09ad8857707a9efce27474fb4937e510dc529a1baf89b6f59, n, c)
tab = G1Table(srs, c)
raw = np.random.default_rng(1).integers(0, 256, size=(n, 32), dtype=np.uint8); raw[:, 31] &= 0x1F
d = c.dev_alloc(n * 32); c.dev_upload(d, raw.tobytes())
out = (ctypes.c_uint8 * 64)(); arr = (ctypes.c_void_p * 1)(d.value); ns = (ctypes.c_size_t * 1)(n)
for it in range(reps):
    c.check(c.lib.nzcb_msm_g1_table_dev(c.h, tab.h, arr, ns, 1, out))
    print(log_n, c.last_device_ms)
