"""PLONK setup / prove / verify -- oracle restatement of snarkjs 0.4.12
``plonk_setup.js``, ``plonk_prove.js``, ``plonk_verify.js`` (un-vendored,
/root/reference/yarn.lock:7279; the only reference call site is the Makefile
recipe /root/reference/Makefile:54-62).  Algorithm per SURVEY.md A.2/A.3/A.5 as
recalled; PARITY UNPINNED at the snarkjs level (no Node in this image).  The
prover and the verifier below are written independently of each other (prover
from the snarkjs data flow, verifier from the PLONK verification equation) and
are cross-checked by a known-trapdoor KZG identity.  Pure Python big-int: use on
domains up to ~2^12.  Test infrastructure only.
"""
import struct

from . import bn254 as b
from .binfile import R1CS, read_binfile, read_wtns, read_zkey_header, section, write_binfile
from .keccak import hash_to_fr
from .ntt import fft, ifft

R = b.R_MOD


# =========================================================================
# setup (A.3)
# =========================================================================
def r1cs_to_plonk(r1cs: R1CS):
    """R1CS -> (gates, additions, plonk_n_vars).  gate = [sl, sr, so, qm, ql, qr, qo, qc]."""
    gates = []
    additions = []
    nvars = [r1cs.n_vars]

    def norm(lc):
        return {s: c % R for s, c in lc.items() if c % R != 0}

    def lc_type(lc):
        k = 0
        n = 0
        for s, c in lc.items():
            if c % R == 0:
                continue
            if s == 0:
                k = (k + c) % R
            else:
                n += 1
        if n > 0:
            return str(n)
        return "k" if k != 0 else "0"

    def reduce_coefs(lc, max_c):
        k = 0
        cs = []
        for s in sorted(lc):  # JS objects iterate integer keys in ascending order
            c = lc[s] % R
            if s == 0:
                k = (k + c) % R
            elif c != 0:
                cs.append((s, c))
        while len(cs) > max_c:
            c1 = cs.pop(0)
            c2 = cs.pop(0)
            so = nvars[0]
            nvars[0] += 1
            gates.append([c1[0], c2[0], so, 0, (-c1[1]) % R, (-c2[1]) % R, 1, 0])
            additions.append((c1[0], c2[0], c1[1], c2[1]))
            cs.append((so, 1))
        ss = [c[0] for c in cs]
        cf = [c[1] for c in cs]
        while len(cf) < max_c:
            ss.append(0)
            cf.append(0)
        return k, ss, cf

    def add_sum(lc):
        k, ss, cf = reduce_coefs(lc, 3)
        gates.append([ss[0], ss[1], ss[2], 0, cf[0], cf[1], cf[2], k])

    def add_mul(la, lb, lc):
        ka, sa, ca = reduce_coefs(la, 1)
        kb, sb, cb = reduce_coefs(lb, 1)
        kc, sc, cc = reduce_coefs(lc, 1)
        gates.append([sa[0], sb[0], sc[0], ca[0] * cb[0] % R, ca[0] * kb % R, ka * cb[0] % R, (-cc[0]) % R,
                      (ka * kb - kc) % R])

    def join(lc1, k, lc2):
        # k * lc1 - lc2  (A = const k:  k * B - C = 0)
        res = {}
        for s, c in lc1.items():
            res[s] = (res.get(s, 0) + k * c) % R
        for s, c in lc2.items():
            res[s] = (res.get(s, 0) - c) % R
        return norm(res)

    for s in range(1, r1cs.n_public + 1):
        gates.append([s, 0, 0, 0, 1, 0, 0, 0])
    for la, lb, lc in r1cs.constraints:
        ta, tb = lc_type(la), lc_type(lb)
        if ta == "0" or tb == "0":
            add_sum(norm(lc))
        elif ta == "k":
            add_sum(join(lb, la.get(0, 0) % R, lc))
        elif tb == "k":
            add_sum(join(la, lb.get(0, 0) % R, lc))
        else:
            add_mul(la, lb, lc)
    return gates, additions, nvars[0]


def pick_k1_k2(power):
    n = 1 << power
    w = b.fr_root(power)
    dom = set()
    x = 1
    for _ in range(n):
        dom.add(x)
        x = x * w % R
    k1 = 2
    while k1 in dom:
        k1 += 1
    k2 = k1 + 1
    while k2 in dom or (k2 * pow(k1, -1, R) % R) in dom:
        k2 += 1
    return k1, k2


def _poly_block(evals, n):
    """n evaluations -> n coefficients + 4n evaluations, as LEM bytes (zkey layout)."""
    coefs = ifft(evals)
    ext = fft(coefs + [0] * (3 * n))
    return coefs, ext


def _lem_list(vals):
    return b"".join(b.to_lem(v) for v in vals)


def setup(r1cs: R1CS, srs_g1, x2_g2_lem: bytes = bytes(128)):
    """snarkjs `plonk setup`: returns zkey bytes.  srs_g1: list of affine [tau^i]G1, len >= n+6."""
    gates, additions, nvars = r1cs_to_plonk(r1cs)
    n_pub = r1cs.n_public
    power = max(3, (len(gates) - 1).bit_length())
    n = 1 << power
    assert len(srs_g1) >= n + 6, "circuit too big for this power of tau ceremony"
    k1, k2 = pick_k1_k2(power) if power <= 12 else (2, 3)
    w = b.fr_root(power)

    sel = []
    for k in range(3, 8):
        ev = [g[k] for g in gates] + [0] * (n - len(gates))
        sel.append(_poly_block(ev, n))

    # sigma: each position gets the id of the signal's previous appearance, the
    # first one closes the cycle with the last
    sigma = [0] * (3 * n)
    last = {}
    first = {}
    x = 1
    for i in range(n):
        sigs = gates[i][:3] if i < len(gates) else (0, 0, 0)
        for col, s in enumerate(sigs):
            p = col * n + i
            if s not in last:
                first[s] = p
            else:
                sigma[p] = last[s]
            last[s] = x if col == 0 else (x * k1 % R if col == 1 else x * k2 % R)
        x = x * w % R
    for s, p in first.items():
        sigma[p] = last[s]
    sig = [_poly_block(sigma[c * n:(c + 1) * n], n) for c in range(3)]

    lag = []
    for i in range(max(n_pub, 1)):
        ev = [0] * n
        ev[i] = 1
        lag.append(_poly_block(ev, n))

    def commit(coefs):
        return b.g1_msm(srs_g1[:n], coefs)

    hdr = struct.pack("<I", 32) + b.P_MOD.to_bytes(32, "little") + struct.pack("<I", 32) + R.to_bytes(32, "little")
    hdr += struct.pack("<IIIII", nvars, n_pub, n, len(additions), len(gates))
    hdr += b.to_lem(k1) + b.to_lem(k2)
    for coefs, _ in sel:
        hdr += b.g1_to_lem(commit(coefs))
    for coefs, _ in sig:
        hdr += b.g1_to_lem(commit(coefs))
    hdr += x2_g2_lem

    sec3 = b"".join(struct.pack("<II", a, bb) + b.to_lem(ac) + b.to_lem(bc) for a, bb, ac, bc in additions)
    maps = [b"".join(struct.pack("<I", g[c]) for g in gates) for c in range(3)]
    secs = [(1, struct.pack("<I", 2)), (2, hdr), (3, sec3), (4, maps[0]), (5, maps[1]), (6, maps[2])]
    for i, (coefs, ext) in enumerate(sel):
        secs.append((7 + i, _lem_list(coefs) + _lem_list(ext)))
    secs.append((12, b"".join(_lem_list(c) + _lem_list(e) for c, e in sig)))
    secs.append((13, b"".join(_lem_list(c) + _lem_list(e) for c, e in lag)))
    secs.append((14, b"".join(b.g1_to_lem(P) for P in srs_g1[:n + 6])))
    return write_binfile(b"zkey", 1, secs)


# =========================================================================
# prove (A.2)
# =========================================================================
class ProverError(Exception):
    pass


def _read_fr_block(buf, count, off=0):
    return [b.from_lem(buf[(off + i) * 32:(off + i + 1) * 32]) for i in range(count)]


def _eval_pol(coefs, x):
    acc = 0
    for c in reversed(coefs):
        acc = (acc * x + c) % R
    return acc


def _div_pol1(P, d):
    n = len(P)
    res = [0] * n
    res[n - 2] = P[n - 1]
    for i in range(n - 3, -1, -1):
        res[i] = (P[i + 1] + d * res[i + 1]) % R
    if P[0] % R != (-d * res[0]) % R:
        raise ProverError("Polinomial does not divide")
    return res


def prove(zkey: bytes, wtns: bytes, blinders):
    """blinders: list of 9 ints b1..b9.  Returns (proof dict, public signals, debug dict)."""
    zk = read_zkey_header(zkey)
    q, w = read_wtns(wtns)
    if q != R:
        raise ProverError("Curve of the witness does not match the curve of the proving key")
    if len(w) != zk.n_vars - zk.n_additions:
        raise ProverError(f"Invalid witness length. Circuit: {zk.n_vars}, witness: {len(w)}, {zk.n_additions}")
    n = zk.domain_size
    power = zk.power
    n_pub = zk.n_public
    k1 = b.from_lem(zk.k1_lem)
    k2 = b.from_lem(zk.k2_lem)
    secs = zk.sections
    bl = [None] + [x % R for x in blinders]

    w = list(w)
    w[0] = 0
    # internal ("addition") signals, file order
    add_sec = section(zkey, secs, 3)
    iw = []
    nw = len(w)

    def get_w(idx):
        if idx < nw:
            return w[idx]
        if idx < zk.n_vars:
            return iw[idx - nw]
        return 0

    for i in range(zk.n_additions):
        a, bb = struct.unpack_from("<II", add_sec, i * 72)
        ac = b.from_lem(add_sec[i * 72 + 8:i * 72 + 40])
        bc = b.from_lem(add_sec[i * 72 + 40:i * 72 + 72])
        iw.append((ac * get_w(a) + bc * get_w(bb)) % R)

    maps = [section(zkey, secs, 4 + c) for c in range(3)]
    A, B, C = ([get_w(struct.unpack_from("<I", maps[c], 4 * i)[0]) for i in range(zk.n_constraints)] +
               [0] * (n - zk.n_constraints) for c in range(3))

    def sel_block(sid):
        s = section(zkey, secs, sid)
        return _read_fr_block(s, n), _read_fr_block(s, 4 * n, n)

    (pol_qm, QM4), (pol_ql, QL4), (pol_qr, QR4), (pol_qo, QO4), (pol_qc, QC4) = (sel_block(7 + i) for i in range(5))
    s12 = section(zkey, secs, 12)
    pol_s1, S14 = _read_fr_block(s12, n, 0), _read_fr_block(s12, 4 * n, n)
    pol_s2, S24 = _read_fr_block(s12, n, 5 * n), _read_fr_block(s12, 4 * n, 6 * n)
    pol_s3, S34 = _read_fr_block(s12, n, 10 * n), _read_fr_block(s12, 4 * n, 11 * n)
    s13 = section(zkey, secs, 13)
    L4 = [_read_fr_block(s13, 4 * n, j * 5 * n + n) for j in range(max(n_pub, 1))]
    s14 = section(zkey, secs, 14)
    ptau = [b.g1_from_lem(s14[i * 64:(i + 1) * 64]) for i in range(n + 6)]

    def exp_tau(coefs):
        return b.g1_msm(ptau[:len(coefs)], coefs)

    def to4t(ev, pz):
        a = ifft(ev)
        a4 = fft(a + [0] * (3 * n))
        a1 = a + [0] * len(pz)
        for i, p in enumerate(pz):
            a1[n + i] = (a1[n + i] + p) % R
            a1[i] = (a1[i] - p) % R
        return a1, a4

    proof = {}
    dbg = {}
    # ---- round 1
    pol_a, A4 = to4t(A, [bl[2], bl[1]])
    pol_b, B4 = to4t(B, [bl[4], bl[3]])
    pol_c, C4 = to4t(C, [bl[6], bl[5]])
    proof["A"], proof["B"], proof["C"] = exp_tau(pol_a), exp_tau(pol_b), exp_tau(pol_c)

    # ---- round 2
    t1 = b"".join(b.to_be(A[i]) for i in range(n_pub))
    t1 += b.g1_to_be(proof["A"]) + b.g1_to_be(proof["B"]) + b.g1_to_be(proof["C"])
    beta = hash_to_fr(t1)
    gamma = hash_to_fr(b.to_be(beta))
    wn = b.fr_root(power)
    num = [1] * n
    den = [1] * n
    x = 1
    for i in range(n):
        nn = (A[i] + beta * x + gamma) * (B[i] + k1 * beta * x + gamma) % R * (C[i] + k2 * beta * x + gamma) % R
        dd = (A[i] + beta * S14[4 * i] + gamma) * (B[i] + beta * S24[4 * i] + gamma) % R * \
             (C[i] + beta * S34[4 * i] + gamma) % R
        num[(i + 1) % n] = num[i] * nn % R
        den[(i + 1) % n] = den[i] * dd % R
        x = x * wn % R
    Z = [num[i] * b.fr_inv(den[i]) % R for i in range(n)]
    if Z[0] != 1:
        raise ProverError("Copy constraints does not match")
    pol_z, Z4 = to4t(Z, [bl[9], bl[8], bl[7]])
    proof["Z"] = exp_tau(pol_z)

    # ---- round 3
    alpha = hash_to_fr(b.g1_to_be(proof["Z"]))
    alpha2 = alpha * alpha % R
    w4 = b.fr_root(2)
    Z1 = [0, (w4 - 1) % R, (-2) % R, (-1 - w4) % R]
    Z2 = [0, (-2 * w4) % R, 4, (2 * w4) % R]
    Z3 = [0, (2 + 2 * w4) % R, (-8) % R, (2 - 2 * w4) % R]

    def mul2(a, bb, ap, bp, p):
        r = a * bb % R
        a0 = (a * bp + ap * bb) % R
        a1 = ap * bp % R
        return r, (a0 + Z1[p] * a1) % R

    def mul4(a, bb, c, d, ap, bp, cp, dp, p):
        a_b, a_bp, ap_b, ap_bp = a * bb % R, a * bp % R, ap * bb % R, ap * bp % R
        c_d, c_dp, cp_d, cp_dp = c * d % R, c * dp % R, cp * d % R, cp * dp % R
        r = a_b * c_d % R
        a0 = (ap_b * c_d + a_bp * c_d + a_b * cp_d + a_b * c_dp) % R
        a1 = (ap_bp * c_d + ap_b * cp_d + ap_b * c_dp + a_bp * cp_d + a_bp * c_dp + a_b * cp_dp) % R
        a2 = (a_bp * cp_dp + ap_b * cp_dp + ap_bp * c_dp + ap_bp * cp_d) % R
        a3 = ap_bp * cp_dp % R
        return r, (a0 + Z1[p] * a1 + Z2[p] * a2 + Z3[p] * a3) % R

    w4n = b.fr_root(power + 2)
    T = [0] * (4 * n)
    Tz = [0] * (4 * n)
    x = 1
    for i in range(4 * n):
        p = i % 4
        a, bb, c, z = A4[i], B4[i], C4[i], Z4[i]
        zw = Z4[(i + 4) % (4 * n)]
        ap = (bl[2] + bl[1] * x) % R
        bp = (bl[4] + bl[3] * x) % R
        cp = (bl[6] + bl[5] * x) % R
        zp = (bl[7] * x * x + bl[8] * x + bl[9]) % R
        xw = x * wn % R
        zwp = (bl[7] * xw * xw + bl[8] * xw + bl[9]) % R
        pl = 0
        for j in range(n_pub):
            pl = (pl - L4[j][i] * A[j]) % R
        e1, e1z = mul2(a, bb, ap, bp, p)
        e1 = (e1 * QM4[i] + a * QL4[i] + bb * QR4[i] + c * QO4[i] + pl + QC4[i]) % R
        e1z = (e1z * QM4[i] + ap * QL4[i] + bp * QR4[i] + cp * QO4[i]) % R
        bx = beta * x % R
        e2, e2z = mul4((a + bx + gamma) % R, (bb + bx * k1 + gamma) % R, (c + bx * k2 + gamma) % R, z, ap, bp, cp, zp, p)
        e3, e3z = mul4((a + beta * S14[i] + gamma) % R, (bb + beta * S24[i] + gamma) % R,
                       (c + beta * S34[i] + gamma) % R, zw, ap, bp, cp, zwp, p)
        e4 = (z - 1) * L4[0][i] % R * alpha2 % R
        e4z = zp * L4[0][i] % R * alpha2 % R
        T[i] = (e1 + e2 * alpha - e3 * alpha + e4) % R
        Tz[i] = (e1z + e2z * alpha - e3z * alpha + e4z) % R
        x = x * w4n % R
    t = ifft(T)
    for i in range(n):
        t[i] = (-t[i]) % R
    for i in range(n, 4 * n):
        t[i] = (t[i - n] - t[i]) % R
        if i > 3 * n - 4 and t[i] != 0:
            raise ProverError("T Polynomial is not divisible")
    tz = ifft(Tz)
    for i in range(4 * n):
        if i > 3 * n + 5:
            if tz[i] != 0:
                raise ProverError("Tz Polynomial is not well calculated")
        else:
            t[i] = (t[i] + tz[i]) % R
    pol_t = t[:3 * n + 6]
    proof["T1"], proof["T2"], proof["T3"] = exp_tau(t[:n]), exp_tau(t[n:2 * n]), exp_tau(t[2 * n:3 * n + 6])

    # ---- round 4
    xi = hash_to_fr(b.g1_to_be(proof["T1"]) + b.g1_to_be(proof["T2"]) + b.g1_to_be(proof["T3"]))
    ev = {}
    ev["a"], ev["b"], ev["c"] = _eval_pol(pol_a, xi), _eval_pol(pol_b, xi), _eval_pol(pol_c, xi)
    ev["s1"], ev["s2"] = _eval_pol(pol_s1, xi), _eval_pol(pol_s2, xi)
    ev["t"] = _eval_pol(pol_t, xi)
    ev["zw"] = _eval_pol(pol_z, xi * wn % R)
    coef_ab = ev["a"] * ev["b"] % R
    bxi = beta * xi % R
    e2 = (ev["a"] + bxi + gamma) * (ev["b"] + bxi * k1 + gamma) % R * (ev["c"] + bxi * k2 + gamma) % R * alpha % R
    e3 = (ev["a"] + beta * ev["s1"] + gamma) * (ev["b"] + beta * ev["s2"] + gamma) % R * beta % R * ev["zw"] % R * alpha % R
    xim = xi
    for _ in range(power):
        xim = xim * xim % R
    eval_l1 = (xim - 1) * b.fr_inv((xi - 1) * n % R) % R
    e4 = eval_l1 * alpha2 % R
    coefz = (e2 + e4) % R
    pol_r = [0] * (n + 3)
    for i in range(n + 3):
        v = coefz * pol_z[i] % R
        if i < n:
            v = (v + coef_ab * pol_qm[i] + ev["a"] * pol_ql[i] + ev["b"] * pol_qr[i] + ev["c"] * pol_qo[i] + pol_qc[i]
                 - e3 * pol_s3[i]) % R
        pol_r[i] = v
    ev["r"] = _eval_pol(pol_r, xi)

    # ---- round 5
    v = [None, hash_to_fr(b"".join(b.to_be(ev[k]) for k in ("a", "b", "c", "s1", "s2", "zw", "r")))]
    for i in range(2, 7):
        v.append(v[i - 1] * v[1] % R)
    xi2m = xim * xim % R
    pol_wxi = [0] * (n + 6)
    for i in range(n + 6):
        ww = xi2m * pol_t[2 * n + i] % R
        if i < n + 3:
            ww = (ww + v[1] * pol_r[i]) % R
        if i < n + 2:
            ww = (ww + v[2] * pol_a[i] + v[3] * pol_b[i] + v[4] * pol_c[i]) % R
        if i < n:
            ww = (ww + pol_t[i] + xim * pol_t[n + i] + v[5] * pol_s1[i] + v[6] * pol_s2[i]) % R
        pol_wxi[i] = ww
    pol_wxi[0] = (pol_wxi[0] - ev["t"] - v[1] * ev["r"] - v[2] * ev["a"] - v[3] * ev["b"] - v[4] * ev["c"]
                  - v[5] * ev["s1"] - v[6] * ev["s2"]) % R
    pol_wxi = _div_pol1(pol_wxi, xi)
    proof["Wxi"] = exp_tau(pol_wxi)
    pol_wxiw = list(pol_z)
    pol_wxiw[0] = (pol_wxiw[0] - ev["zw"]) % R
    pol_wxiw = _div_pol1(pol_wxiw, xi * wn % R)
    proof["Wxiw"] = exp_tau(pol_wxiw)

    for k in ("a", "b", "c", "s1", "s2", "zw", "r"):
        proof["eval_" + k] = ev[k]
    public = [w_ for w_ in read_wtns(wtns)[1][1:n_pub + 1]]
    dbg.update(beta=beta, gamma=gamma, alpha=alpha, xi=xi, v=v, eval_t=ev["t"], pol_t=pol_t, pol_a=pol_a, pol_z=pol_z,
               A=A, B=B, C=C, Z=Z)
    return proof, public, dbg


PROOF_POINTS = ("A", "B", "C", "Z", "T1", "T2", "T3", "Wxi", "Wxiw")
PROOF_EVALS = ("eval_a", "eval_b", "eval_c", "eval_s1", "eval_s2", "eval_zw", "eval_r")


def proof_to_bytes(proof):
    """nzcb_proof binary layout (include/nzcb.h)."""
    return b"".join(b.g1_to_be(proof[k]) for k in PROOF_POINTS) + b"".join(b.to_be(proof[k]) for k in PROOF_EVALS)


def proof_from_bytes(buf):
    proof = {}
    for i, k in enumerate(PROOF_POINTS):
        raw = buf[i * 64:(i + 1) * 64]
        proof[k] = None if raw == bytes(64) else (int.from_bytes(raw[:32], "big"), int.from_bytes(raw[32:], "big"))
    for i, k in enumerate(PROOF_EVALS):
        proof[k] = int.from_bytes(buf[576 + i * 32:576 + (i + 1) * 32], "big")
    return proof


def proof_to_json(proof):
    """proof.json text as snarkjs prints it (JSON.stringify(proof, null, 1) of stringifyBigInts)."""
    import json
    order = ["A", "B", "C", "Z", "T1", "T2", "T3", "eval_a", "eval_b", "eval_c", "eval_s1", "eval_s2", "eval_zw",
             "eval_r", "Wxi", "Wxiw"]
    obj = {}
    for k in order:
        v = proof[k]
        if k in PROOF_POINTS:
            obj[k] = ["0", "1", "0"] if v is None else [str(v[0]), str(v[1]), "1"]
        else:
            obj[k] = str(v)
    obj["protocol"] = "plonk"
    obj["curve"] = "bn128"
    return json.dumps(obj, indent=1)


# =========================================================================
# verify (A.5) -- written from the PLONK verification equation
# =========================================================================
def verification_key(zkey: bytes):
    zk = read_zkey_header(zkey)
    vk = {"nPublic": zk.n_public, "power": zk.power, "k1": b.from_lem(zk.k1_lem), "k2": b.from_lem(zk.k2_lem),
          "w": b.fr_root(zk.power), "X_2": zk.X_2_lem}
    for nm, raw in zk.commits_lem.items():
        vk[nm] = b.g1_from_lem(raw)
    return vk


def vk_from_json(j):
    """verification_key.json object (decimal strings, `snarkjs zkey export verificationkey`) -> vk dict"""
    vk = {"nPublic": int(j["nPublic"]), "power": int(j["power"]), "k1": int(j["k1"]), "k2": int(j["k2"]),
          "w": int(j["w"]), "X_2": j.get("X_2")}
    for nm in ("Qm", "Ql", "Qr", "Qo", "Qc", "S1", "S2", "S3"):
        x, y, z = j[nm]
        vk[nm] = None if z == "0" else (int(x), int(y))
    assert vk["w"] == b.fr_root(vk["power"])
    return vk


def _verify_points(vk, public, proof):
    """everything of plonk.verify up to the pairing: -> (A1, B1) with the proof accepted iff
    e(A1, X_2) == e(B1, [1]_2); None when the proof is not well constructed"""
    n = 1 << vk["power"]
    for k in PROOF_POINTS:
        if proof[k] is not None and not (0 <= proof[k][0] < b.P_MOD and 0 <= proof[k][1] < b.P_MOD):
            return None
        if not b.g1_is_on_curve(proof[k]):
            return None
    for k in PROOF_EVALS:
        if not (0 <= proof[k] < R):
            return None
    if len(public) != vk["nPublic"]:
        return None
    g = b.g1_to_be
    beta = hash_to_fr(b"".join(b.to_be(p % R) for p in public) + g(proof["A"]) + g(proof["B"]) + g(proof["C"]))
    gamma = hash_to_fr(b.to_be(beta))
    alpha = hash_to_fr(g(proof["Z"]))
    xi = hash_to_fr(g(proof["T1"]) + g(proof["T2"]) + g(proof["T3"]))
    v1 = hash_to_fr(b"".join(b.to_be(proof[k]) for k in PROOF_EVALS))
    v = [None, v1]
    for i in range(2, 7):
        v.append(v[-1] * v1 % R)
    u = hash_to_fr(g(proof["Wxi"]) + g(proof["Wxiw"]))
    xin = pow(xi, n, R)
    zh = (xin - 1) % R
    w = vk["w"]
    lag = []
    for i in range(max(1, vk["nPublic"])):
        wi = pow(w, i, R)
        lag.append(wi * zh % R * b.fr_inv(n * (xi - wi) % R) % R)
    pl = 0
    for i, p in enumerate(public):
        pl = (pl - p * lag[i]) % R
    ea, eb, ec = proof["eval_a"], proof["eval_b"], proof["eval_c"]
    es1, es2, ezw, er = proof["eval_s1"], proof["eval_s2"], proof["eval_zw"], proof["eval_r"]
    alpha2 = alpha * alpha % R
    t = (er + pl - alpha * (ea + beta * es1 + gamma) % R * (eb + beta * es2 + gamma) % R * (ec + gamma) % R * ezw
         - alpha2 * lag[0]) % R * b.fr_inv(zh) % R
    k1, k2 = vk["k1"], vk["k2"]
    # D = v1 * [r(X) without the constant parts]  + u * Z
    dz = (v1 * (alpha * (ea + beta * xi + gamma) % R * (eb + beta * k1 * xi + gamma) % R * (ec + beta * k2 * xi + gamma)
                + alpha2 * lag[0]) + u) % R
    ds3 = v1 * alpha % R * beta % R * ezw % R * (ea + beta * es1 + gamma) % R * (eb + beta * es2 + gamma) % R
    D = b.g1_msm_naive(
        [vk["Qm"], vk["Ql"], vk["Qr"], vk["Qo"], vk["Qc"], proof["Z"], vk["S3"]],
        [v1 * ea * eb % R, v1 * ea % R, v1 * eb % R, v1 * ec % R, v1, dz, (-ds3) % R])
    F = b.g1_msm_naive([proof["T1"], proof["T2"], proof["T3"], proof["A"], proof["B"], proof["C"], vk["S1"], vk["S2"]],
                       [1, xin, xin * xin % R, v[2], v[3], v[4], v[5], v[6]])
    F = b.g1_add(F, D)
    e = (t + v1 * er + v[2] * ea + v[3] * eb + v[4] * ec + v[5] * es1 + v[6] * es2 + u * ezw) % R
    E = b.g1_mul(b.G1_GEN, e)
    lhs = b.g1_add(proof["Wxi"], b.g1_mul(proof["Wxiw"], u))
    rhs = b.g1_msm_naive([proof["Wxi"], proof["Wxiw"]], [xi, u * xi % R * w % R])
    rhs = b.g1_add(rhs, b.g1_sub(F, E))
    return lhs, rhs


def verify_with_trapdoor(vk, public, proof, tau):
    """plonk.verify with the final pairing e(W1, [tau]_2) == e(W2, [1]_2) replaced by the
    equivalent G1 identity tau * W1 == W2, valid because the SRS trapdoor tau is known."""
    pts = _verify_points(vk, public, proof)
    return pts is not None and b.g1_mul(pts[0], tau) == pts[1]


def vk_x2(vk):
    """X_2 of a vk as an oracle G2 point: 128 LEM bytes (zkey header) or the JSON triple of Fq2 pairs"""
    from . import pairing as pg

    x2 = vk["X_2"]
    if isinstance(x2, (bytes, bytearray)):
        return pg.g2_from_lem(bytes(x2))
    if x2[2] == ["0", "0"]:
        return None
    return ((int(x2[0][0]), int(x2[0][1])), (int(x2[1][0]), int(x2[1][1])))


def verify(vk, public, proof):
    """snarkjs plonk.verify with the real pairing check
    curve.pairingEq(-A1, X_2, B1, G2.one): e(-A1, X_2) e(B1, [1]_2) == 1"""
    from . import pairing as pg

    pts = _verify_points(vk, public, proof)
    if pts is None:
        return False
    X2 = vk_x2(vk)
    if X2 is not None and not pg.g2_is_on_curve(X2):
        return False
    return pg.pairing_eq([(b.g1_neg(pts[0]), X2), (pts[1], pg.G2_GEN)])
