import json, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nzcb_circom_b200 import Context
c = Context(0)
res = {}
print('selftest mismatches', c.selftest_mul())
for kind, name in ((0, 'imad32'), (1, 'imad_wide'), (4, 'imad_hi'), (2, 'fr_mul'), (3, 'fq_mul'), (5, 'fr_mul_portable')):
    for bps in (4, 8):
        v = c.microbench(kind, 4000 if kind < 2 else 1000, bps)
        res[f'{name}_bps{bps}'] = {'ops_per_s': v, 'ms': c.last_device_ms}
        print(name, bps, f'{v:.4e}', c.last_device_ms)
json.dump(res, open('gpurun_out/microbench.json', 'w'), indent=1)
