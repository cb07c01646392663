"""TEST INFRASTRUCTURE -- writes a small, insecure (known-trapdoor) .ptau the way `snarkjs powersoftau new` +
`prepare phase2` lay it out (iden3 binfile "ptau" v1; the role of /root/reference/Makefile:64-67), so
the product's reader (csrc/ptau.cu, nzcb_plonk_setup_ptau) has a file to read.  Sections: 1 header (n8, q, power,
ceremonyPower), 2 tauG1 (2^(power+1) - 1 points), 3 tauG2 (2^power), 4 alphaTauG1, 5 betaTauG1, 6 betaG2,
7 contributions (none), 12 Lagrange tauG1 for every 2^p, p <= power (13-15 omitted: plonk setup never reads them).
Format recalled from snarkjs 0.4.12 (un-vendored): parity unpinned."""
import struct

from . import bn254 as b
from . import pairing as pg
from .binfile import write_binfile


def write_ptau(tau, power, alpha=3, beta=5, prepared=True):
    n = 1 << power
    hdr = struct.pack("<I", 32) + b.P_MOD.to_bytes(32, "little") + struct.pack("<II", power, power)
    tau_g1 = b.srs_g1(tau, 2 * n - 1)
    s2 = b"".join(b.g1_to_lem(p) for p in tau_g1)
    g2, acc, pts2 = pg.G2_GEN, 1, []
    for _ in range(n):
        pts2.append(pg.g2_mul(g2, acc))
        acc = acc * tau % b.R_MOD
    s3 = b"".join(pg.g2_to_lem(p) for p in pts2)
    s4 = b"".join(b.g1_to_lem(b.g1_mul(p, alpha)) for p in tau_g1[:n])
    s5 = b"".join(b.g1_to_lem(b.g1_mul(p, beta)) for p in tau_g1[:n])
    s6 = pg.g2_to_lem(pg.g2_mul(g2, beta))
    s7 = struct.pack("<I", 0)
    sections = [(1, hdr), (2, s2), (3, s3), (4, s4), (5, s5), (6, s6), (7, s7)]
    if prepared:
        lag = []
        for p in range(power + 1):
            m = 1 << p
            w = b.fr_root(p)
            # L_i(tau) = (tau^m - 1) w^i / (m (tau - w^i))
            zt = (pow(tau, m, b.R_MOD) - 1) % b.R_MOD
            for i in range(m):
                wi = pow(w, i, b.R_MOD)
                li = zt * wi % b.R_MOD * b.fr_inv(m * (tau - wi) % b.R_MOD) % b.R_MOD
                lag.append(b.g1_to_lem(b.g1_mul(b.G1_GEN, li)))
        sections.append((12, b"".join(lag)))
    return write_binfile(b"ptau", 1, sections)
