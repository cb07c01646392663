"""Device-resident Fr NTT timing through the C ABI: python tools/ntt_probe.py [log_n ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nzcb_circom_b200 import Context
c = Context(0)
for log_n in [int(a) for a in sys.argv[1:]] or [21, 23]:
    n = 1 << log_n
    raw = np.random.default_rng(log_n).integers(0, 256, size=(n, 32), dtype=np.uint8)
    raw[:, 31] &= 0x1F
    d = c.dev_alloc(n * 32)
    c.dev_upload(d, raw.tobytes())
    ts = []
    for it in range(6):
        c.check(c.lib.nzcb_ntt_fr_dev(c.h, d, log_n, it & 1))
        ts.append(c.last_device_ms)
    mm = (n // 2) * log_n
    best = min(ts[2:])
    print(f"log_n {log_n}: {best:.3f} ms  ({mm / best / 1e6:.1f} G butterflies/s, {2 * n * 32 / best / 1e6:.0f} GB/s per-pass-equivalent)")
    c.dev_free(d)
