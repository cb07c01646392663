"""iden3 sectioned binary files (.wtns, .zkey, .r1cs) -- oracle reader/writer.

Restates @iden3/binfileutils 0.0.10 + r1csfile 0.0.35 + the snarkjs wtns/zkey
headers (un-vendored, /root/reference/yarn.lock:843,6692,7279) per SURVEY.md
A.4: 4-byte magic, u32 version, u32 nSections, then {u32 id, u64 size,
payload}; all integers little-endian.  Test infrastructure only.
"""
import struct

from .bn254 import P_MOD, R_MOD


def write_binfile(magic: bytes, version: int, sections):
    """sections: list of (id, bytes) in file order."""
    out = bytearray()
    out += magic
    out += struct.pack("<II", version, len(sections))
    for sid, payload in sections:
        out += struct.pack("<IQ", sid, len(payload))
        out += payload
    return bytes(out)


def read_binfile(data: bytes, magic: bytes):
    assert data[:4] == magic, f"bad magic {data[:4]!r}"
    version, nsec = struct.unpack_from("<II", data, 4)
    pos = 12
    sections = {}
    for _ in range(nsec):
        sid, size = struct.unpack_from("<IQ", data, pos)
        pos += 12
        sections.setdefault(sid, []).append((pos, size))
        pos += size
    return version, sections


def section(data, sections, sid):
    pos, size = sections[sid][0]
    return data[pos:pos + size]


# ---- .wtns ---------------------------------------------------------------
def write_wtns(witness):
    """witness: list of ints (canonical).  v2, section 1 header, section 2 values LE."""
    hdr = struct.pack("<I", 32) + R_MOD.to_bytes(32, "little") + struct.pack("<I", len(witness))
    body = b"".join(w.to_bytes(32, "little") for w in witness)
    return write_binfile(b"wtns", 2, [(1, hdr), (2, body)])


def read_wtns(data):
    _, secs = read_binfile(data, b"wtns")
    hdr = section(data, secs, 1)
    n8 = struct.unpack_from("<I", hdr, 0)[0]
    q = int.from_bytes(hdr[4:4 + n8], "little")
    nw = struct.unpack_from("<I", hdr, 4 + n8)[0]
    body = section(data, secs, 2)
    return q, [int.from_bytes(body[i * n8:(i + 1) * n8], "little") for i in range(nw)]


# ---- .r1cs ---------------------------------------------------------------
class R1CS:
    """n_vars wires (wire 0 == 1), outputs then public inputs then private inputs
    directly after wire 0; constraints = list of (A, B, C) dicts wire -> coef."""

    def __init__(self, n_vars, n_pub_out, n_pub_in, n_prv_in, constraints):
        self.n_vars = n_vars
        self.n_pub_out = n_pub_out
        self.n_pub_in = n_pub_in
        self.n_prv_in = n_prv_in
        self.constraints = constraints

    @property
    def n_public(self):
        return self.n_pub_out + self.n_pub_in


def write_r1cs(r: R1CS):
    hdr = struct.pack("<I", 32) + R_MOD.to_bytes(32, "little")
    hdr += struct.pack("<IIIIQI", r.n_vars, r.n_pub_out, r.n_pub_in, r.n_prv_in, r.n_vars, len(r.constraints))
    body = bytearray()
    for lcs in r.constraints:
        for lc in lcs:
            items = sorted(lc.items())
            body += struct.pack("<I", len(items))
            for w, c in items:
                body += struct.pack("<I", w) + (c % R_MOD).to_bytes(32, "little")
    wmap = b"".join(struct.pack("<Q", i) for i in range(r.n_vars))
    return write_binfile(b"r1cs", 1, [(1, hdr), (2, bytes(body)), (3, wmap)])


def read_r1cs(data):
    _, secs = read_binfile(data, b"r1cs")
    hdr = section(data, secs, 1)
    n8 = struct.unpack_from("<I", hdr, 0)[0]
    assert int.from_bytes(hdr[4:4 + n8], "little") == R_MOD
    n_vars, n_out, n_pub, n_prv, _nlabels, n_cons = struct.unpack_from("<IIIIQI", hdr, 4 + n8)
    body = section(data, secs, 2)
    pos = 0
    cons = []
    for _ in range(n_cons):
        lcs = []
        for _k in range(3):
            nt = struct.unpack_from("<I", body, pos)[0]
            pos += 4
            lc = {}
            for _t in range(nt):
                w = struct.unpack_from("<I", body, pos)[0]
                c = int.from_bytes(body[pos + 4:pos + 4 + n8], "little")
                pos += 4 + n8
                lc[w] = c
            lcs.append(lc)
        cons.append(tuple(lcs))
    return R1CS(n_vars, n_out, n_pub, n_prv, cons)


# ---- .zkey (PLONK, v1, 14 sections) ----------------------------------------
ZKEY_PLONK_PROTOCOL = 2


class ZkeyHeader:
    pass


def read_zkey_header(data):
    _, secs = read_binfile(data, b"zkey")
    assert struct.unpack("<I", section(data, secs, 1))[0] == ZKEY_PLONK_PROTOCOL, "zkey file is not plonk"
    h = section(data, secs, 2)
    z = ZkeyHeader()
    pos = 0
    n8q = struct.unpack_from("<I", h, pos)[0]
    pos += 4
    z.q = int.from_bytes(h[pos:pos + n8q], "little")
    pos += n8q
    n8r = struct.unpack_from("<I", h, pos)[0]
    pos += 4
    z.r = int.from_bytes(h[pos:pos + n8r], "little")
    pos += n8r
    assert z.q == P_MOD and z.r == R_MOD
    z.n_vars, z.n_public, z.domain_size, z.n_additions, z.n_constraints = struct.unpack_from("<IIIII", h, pos)
    pos += 20
    z.k1_lem = h[pos:pos + 32]
    z.k2_lem = h[pos + 32:pos + 64]
    pos += 64
    names = ["Qm", "Ql", "Qr", "Qo", "Qc", "S1", "S2", "S3"]
    z.commits_lem = {}
    for nm in names:
        z.commits_lem[nm] = h[pos:pos + 64]
        pos += 64
    z.X_2_lem = h[pos:pos + 128]
    z.power = z.domain_size.bit_length() - 1
    z.sections = secs
    return z
