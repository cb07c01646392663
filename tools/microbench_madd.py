import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nzcb_circom_b200 import Context
c = Context(0)
res = {}
names = {0: "inline 128x4", 1: "call 128x4", 2: "inline 128x3", 3: "inline 256x2", 4: "call 256x3", 5: "inline 128x5", 6: "inline 128x6", 7: "inline 128x8"}
fq = c.microbench(3, 1000, 8)
print(f"fq_mul peak {fq:.4e}/s")
for log_table in (10,):
    for v in range(8):
        r = c.microbench_madd(v, 2000, log_table)
        res[f"v{v}_t{log_table}"] = {"variant": names[v], "madds_per_s": r, "modmul_per_s": 10 * r, "frac_of_fq_mul_peak": 10 * r / fq, "ms": c.last_device_ms}
        print(names[v], log_table, f"{r:.4e} madd/s  {10*r/fq:.3f} of mul peak", c.last_device_ms)
json.dump(res, open("gpurun_out/microbench_madd.json", "w"), indent=1)
