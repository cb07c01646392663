"""CPU restatement of the GPU witness interpreter (csrc/witness.cu): executes a
``.wprog`` witness program the way circom_runtime's WitnessCalculator runs a
circom-generated WASM (un-vendored, /root/reference/yarn.lock:2496; call sites
/root/reference/test/nzcp.js:42 and every ``calculateWitness`` in test/cbor.js,
test/quinSelector.js).  Pure Python ints.  Test infrastructure only.

Independent checks that do NOT go through this program: ``check_r1cs`` (every
constraint of the .r1cs holds for the witness) and the hashlib / cbor
expectations in the tests."""
import struct

from .binfile import read_r1cs
from .bn254 import R_MOD as R

OP_LIN, OP_MUL, OP_BITS, OP_INV, OP_ASSERT, OP_BITSLC = 1, 2, 3, 4, 5, 6


class AssertFailed(Exception):
    """circom_runtime error 4: "Assert Failed" """


class Program:
    def __init__(self, data: bytes):
        assert data[:4] == b"NZWP"
        (ver, self.n_total, self.n_witness, self.n_out, self.n_in, n_consts, self.n_instr, self.n_levels,
         n_code) = struct.unpack_from("<IIIIIIIII", data, 4)
        assert ver == 1
        pos = 40
        self.consts = [int.from_bytes(data[pos + 32 * i:pos + 32 * i + 32], "little") for i in range(n_consts)]
        pos += 32 * n_consts
        self.ioff = struct.unpack_from(f"<{self.n_instr}I", data, pos)
        pos += 4 * self.n_instr
        self.lstart = struct.unpack_from(f"<{self.n_levels + 1}I", data, pos)
        pos += 4 * (self.n_levels + 1)
        self.code = struct.unpack_from(f"<{n_code}I", data, pos)
        assert pos + 4 * n_code == len(data)


def run(prog: Program, inputs, check=True, full=False):
    """inputs: n_in ints in declaration order.  Returns the witness (n_witness ints; full=True: the program temps too).
    Raises AssertFailed when a `===` of the circuit does not hold (sanityCheck = true)."""
    assert len(inputs) == prog.n_in
    w = [0] * prog.n_total
    w[0] = 1
    for i, v in enumerate(inputs):
        w[1 + prog.n_out + i] = v % R
    code, consts = prog.code, prog.consts

    def lc(p):
        n = code[p]
        ci = code[p + 1]
        acc = consts[ci] if ci != 0xFFFFFFFF else 0
        p += 2
        for _ in range(n):
            acc += consts[code[p + 1]] * w[code[p]]
            p += 2
        return acc % R, p

    for off in prog.ioff:
        op = code[off]
        if op == OP_LIN:
            v, _ = lc(off + 2)
            w[code[off + 1]] = v
        elif op == OP_MUL:
            a, p = lc(off + 2)
            b, p = lc(p)
            c, p = lc(p)
            w[code[off + 1]] = (a * b + c) % R
        elif op == OP_BITS:
            dst, src, n = code[off + 1], code[off + 2], code[off + 3]
            v = w[src]
            for i in range(n):
                w[dst + i] = (v >> i) & 1
        elif op == OP_INV:
            v = w[code[off + 2]]
            w[code[off + 1]] = pow(v, R - 2, R) if v else 0
        elif op == OP_BITSLC:
            dst, n = code[off + 1], code[off + 2]
            v, _ = lc(off + 3)
            for i in range(n):
                w[dst + i] = (v >> i) & 1
        elif op == OP_ASSERT:
            a, p = lc(off + 1)
            b, p = lc(p)
            c, p = lc(p)
            if check and (a * b - c) % R != 0:
                raise AssertFailed("Assert Failed")
        else:
            raise ValueError(f"bad opcode {op}")
    return w if full else w[:prog.n_witness]


def check_r1cs(r1cs_bytes: bytes, witness):
    """index of the first violated constraint, or -1 if the witness satisfies the whole R1CS"""
    r = read_r1cs(r1cs_bytes)
    assert len(witness) == r.n_vars

    def ev(lc):
        return sum(c * witness[s] for s, c in lc.items()) % R

    for i, (a, b, c) in enumerate(r.constraints):
        if (ev(a) * ev(b) - ev(c)) % R != 0:
            return i
    return -1
