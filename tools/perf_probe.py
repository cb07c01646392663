"""Full-size probe: a synthetic chain circuit x' = x*x + x + 5 sized to a 2^k PLONK domain ->
SRS + plonk setup + prove on the GPU, proof checked with the oracle verifier (known trapdoor)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bn254 as b
from oracle import plonk as oplonk
from oracle.keccak import hash_to_fr
from nzcb_circom_b200 import Context
from nzcb_circom_b200.snarkjs import ZKey, plonk, powersoftau, write_wtns

R = b.R_MOD


def chain_circuit(power):
    m = (1 << (power - 1)) - 8  # 2 gates per constraint
    # wires: 0 = 1, 1 = out (= last x), 2 = x0, 3.. = x1..
    w = [1, 0, 3]
    x = 3
    for _ in range(m - 1):
        x = (x * x + x + 5) % R
        w.append(x)
    out = (x * x + x + 5) % R
    w[1] = out
    nw = len(w)
    rec = np.dtype([("na", "<u4"), ("aw", "<u4"), ("ac", "V32"), ("nb", "<u4"), ("bw", "<u4"), ("bc", "V32"),
                    ("nc", "<u4"), ("c0w", "<u4"), ("c0c", "V32"), ("c1w", "<u4"), ("c1c", "V32"), ("c2w", "<u4"),
                    ("c2c", "V32")])
    arr = np.zeros(m, dtype=rec)
    one = np.frombuffer((1).to_bytes(32, "little"), dtype="V32")[0]
    m1 = np.frombuffer((R - 1).to_bytes(32, "little"), dtype="V32")[0]
    m5 = np.frombuffer((R - 5).to_bytes(32, "little"), dtype="V32")[0]
    src = np.arange(2, 2 + m, dtype=np.uint32)
    dst = src + 1
    dst[-1] = 1
    arr["na"] = 1; arr["aw"] = src; arr["ac"] = one
    arr["nb"] = 1; arr["bw"] = src; arr["bc"] = one
    arr["nc"] = 3; arr["c0w"] = 0; arr["c0c"] = m5
    # keep wires ascending inside the LC where possible (order is normalised by the setup anyway)
    arr["c1w"] = src; arr["c1c"] = m1; arr["c2w"] = dst; arr["c2c"] = one
    body = arr.tobytes()
    import struct
    hdr = struct.pack("<I", 32) + R.to_bytes(32, "little") + struct.pack("<IIIIQI", nw, 1, 0, 1, nw, m)
    wmap = np.arange(nw, dtype="<u8").tobytes()
    r1cs = b"r1cs" + struct.pack("<II", 1, 3)
    for sid, pl in ((1, hdr), (2, body), (3, wmap)):
        r1cs += struct.pack("<IQ", sid, len(pl)) + pl
    return r1cs, w


def main():
    power = int(sys.argv[1]) if len(sys.argv) > 1 else 21
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    tau = hash_to_fr(b"nzcb-b200-tau")
    t = time.time(); r1cs, w = chain_circuit(power); print(f"circuit: {time.time()-t:.1f}s nw={len(w)}", flush=True)
    wt = write_wtns(w)
    ctx = Context(0)
    n = 1 << power
    t = time.time(); srs = powersoftau.new_g1(tau, n + 6, ctx); print(f"srs: {time.time()-t:.2f}s", flush=True)
    t = time.time(); zkey = plonk.setup(r1cs, srs, powersoftau.new_g2(tau, ctx), ctx); print(f"setup: {time.time()-t:.2f}s zkey={len(zkey)/2**30:.2f} GiB", flush=True)
    t = time.time(); zk = ZKey(zkey, ctx); print(f"zkey load: {time.time()-t:.2f}s n={zk.domain_size} nVars={zk.n_vars} nAdd={zk.n_additions}", flush=True)
    vk = oplonk.verification_key(zkey[:4096 + 2000])if False else None
    bl = list(range(101, 110))
    for i in range(reps):
        t = time.time(); proof, pub = plonk.prove(zk, wt, blinders=bl, raw=True)
        print(f"prove[{i}]: wall {time.time()-t:.3f}s device {ctx.last_device_ms:.1f} ms launches {ctx.launches}", flush=True)
    # verify (header-only parse of the zkey is enough for the vk)
    from oracle.binfile import read_zkey_header
    vk = oplonk.verification_key(zkey)
    ok = oplonk.verify_with_trapdoor(vk, [int(x) for x in pub], oplonk.proof_from_bytes(proof), tau)
    print("verifies:", ok, "public:", pub)
    assert ok and int(pub[0]) == w[1]


if __name__ == "__main__":
    main()
