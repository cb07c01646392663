"""Word-level SHA-2 steps of the native witness program (builder.OP_SHAROUND / OP_SHASCHED, csrc/witness.cu
sha_step_warp; circuits/nzcptpl.circom:509-516 Sha256Var, :577-580 Sha512).  The generic program -- one instruction per
wire, what the oracle VMs run -- is the reference: a Python restatement of the word-level semantics must reproduce,
from the generic witness, every wire of every fused step at its place (CPU), and the GPU running the native program
must return the generic witness wire for wire (GPU).  Digests are checked against hashlib on the way."""
import hashlib
import random

import pytest

from nzcb_circom_b200.circom import sha2
from nzcb_circom_b200.circom.builder import OP_SHAROUND, OP_SHASCHED, R, Circuit
from oracle import witness_vm as vm


def _sha256_circuit(block_space=1):
    c = Circuit(f"sha256var{block_space}_test")
    nbits = 512 << block_space
    out = c.output("out", 256)
    bits = c.input("in", nbits)
    ln = c.input("len")
    dig = sha2.sha256_var(c, bits, ln, block_space)
    for i in range(256):
        c.assign_output(out[i], dig[i])
    return c


def _sha512_circuit():
    c = Circuit("sha512_test")
    out = c.output("out", 512)
    bits = c.input("in", 512)
    dig = sha2.sha512_fixed(c, bits)
    for i in range(512):
        c.assign_output(out[i], dig[i])
    return c


def _rotr(x, r, n):
    r %= n
    return ((x >> r) | (x << (n - r))) & ((1 << n) - 1) if r else x


def _words(pl, W, nw):
    def bit(lc):
        v = lc.k
        for w, cf in lc.t.items():
            v += cf * W[w if w >= 0 else nw + (-w - 1)]
        return (v % R) & 1

    return [sum(bit(b) << i for i, b in enumerate(word)) for word in pl["words"]]


def _emulate(op, pl, W, nw, x=None):
    """the wires [w0, w0 + size) a word-level step writes, from the witness W (csrc/witness.cu sha_step_warp);
    x: the step's input words when they do not come from the witness (inside a block instruction)"""
    n, mask = pl["n"], (1 << pl["n"]) - 1
    if x is None:
        x = _words(pl, W, nw)
    r1a, r1b, r1c = pl["rot1"]
    r0a, r0b, r0c = pl["rot0"]
    out = {}

    def put_xor3(base, mid, o, n_mid):
        for i in range(n):
            if i < n_mid:
                out[base + 2 * i], out[base + 2 * i + 1] = (mid >> i) & 1, (o >> i) & 1
            else:
                out[base + 2 * n_mid + i - n_mid] = (o >> i) & 1

    def put_bits(base, v, count):
        for j in range(count):
            out[base + j] = (v >> j) & 1

    w0 = pl["w0"]
    if op == OP_SHAROUND:
        A, B, C, D, E, F, G, H, Wt = x
        e2, e3 = _rotr(E, r1b, n), _rotr(E, r1c, n)
        S1 = _rotr(E, r1a, n) ^ e2 ^ e3
        ch = (E & F) ^ (~E & G & mask)
        t1 = H + S1 + ch + pl["K"] + Wt
        a2, a3 = _rotr(A, r0b, n), _rotr(A, r0c, n)
        S0 = _rotr(A, r0a, n) ^ a2 ^ a3
        mj = (A & B) ^ (A & C) ^ (B & C)
        t2 = S0 + mj
        put_xor3(w0, e2 & e3, S1, n)
        put_bits(w0 + 2 * n, ch, n)
        put_bits(w0 + 3 * n, t1, n + 3)
        put_xor3(w0 + 4 * n + 3, a2 & a3, S0, n)
        put_xor3(w0 + 6 * n + 3, B & C, mj, n)
        put_bits(w0 + 8 * n + 3, t2, n + 1)
        put_bits(w0 + 9 * n + 4, D + (t1 & mask), n + 1)
        put_bits(w0 + 10 * n + 5, (t1 & mask) + (t2 & mask), n + 1)
    else:
        b1, c1 = _rotr(x[0], r1b, n), x[0] >> r1c
        s1 = _rotr(x[0], r1a, n) ^ b1 ^ c1
        b0, c0 = _rotr(x[2], r0b, n), x[2] >> r0c
        s0 = _rotr(x[2], r0a, n) ^ b0 ^ c0
        o_s0 = 2 * n - r1c
        o_w = o_s0 + 2 * n - r0c
        put_xor3(w0, b1 & c1, s1, n - r1c)
        put_xor3(w0 + o_s0, b0 & c0, s0, n - r0c)
        put_bits(w0 + o_w, s1 + x[1] + s0 + x[3], n + 2)
    assert sorted(out) == list(range(w0, w0 + pl["size"]))
    return out


def _emulate_block(pl, W, nw):
    """OP_SHABLOCK (csrc/witness.cu sha_block_warp): state and message words gathered once, then the schedule steps
    and rounds on words, each writing its wires where the per-step instructions would"""
    from nzcb_circom_b200.circom.builder import OP_SHAROUND as RND, OP_SHASCHED as SCH

    n, mask = pl["n"], (1 << pl["n"]) - 1
    x = _words(pl, W, nw)
    a, b, c, d, e, f, g, h = x[:8]
    w = list(x[8:24])
    out = {}
    s_size = 5 * n + 2 - pl["srot1"][2] - pl["srot0"][2]
    for t in range(pl["r_start"], pl["rounds"]):
        if t >= 16:
            st = {"n": n, "rot1": pl["srot1"], "rot0": pl["srot0"], "w0": pl["sched_w0"][t - 16], "size": s_size}
            o = _emulate(SCH, st, W, nw, [w[t - 2], w[t - 7], w[t - 15], w[t - 16]])
            out.update(o)
            o_w = st["w0"] + 4 * n - pl["srot1"][2] - pl["srot0"][2]
            w.append(sum(o[o_w + i] << i for i in range(n)))
        rd = {"n": n, "rot1": pl["rot1"], "rot0": pl["rot0"], "K": pl["K"][t - pl["r_start"]],
              "w0": pl["round_w0"][t - pl["r_start"]], "size": 11 * n + 6}
        o = _emulate(RND, rd, W, nw, [a, b, c, d, e, f, g, h, w[t]])
        out.update(o)
        en = sum(o[rd["w0"] + 9 * n + 4 + i] << i for i in range(n))
        an = sum(o[rd["w0"] + 10 * n + 5 + i] << i for i in range(n))
        h, g, f, e = g, f, e, en
        d, c, b, a = c, b, a, an
    return out


def _emulate_rounds(pl, W, nw):
    """OP_SHAROUNDS (csrc/witness.cu sha_rounds_warp): state gathered once, w[t] read round by round"""
    from nzcb_circom_b200.circom.builder import OP_SHAROUND as RND

    n = pl["n"]
    x = _words(pl, W, nw)
    a, b, c, d, e, f, g, h = x[:8]
    out = {}
    for i in range(pl["rounds"] - pl["r_start"]):
        rd = {"n": n, "rot1": pl["rot1"], "rot0": pl["rot0"], "K": pl["K"][i], "w0": pl["round_w0"][i], "size": 11 * n + 6}
        o = _emulate(RND, rd, W, nw, [a, b, c, d, e, f, g, h, x[8 + i]])
        out.update(o)
        en = sum(o[rd["w0"] + 9 * n + 4 + k] << k for k in range(n))
        an = sum(o[rd["w0"] + 10 * n + 5 + k] << k for k in range(n))
        h, g, f, e = g, f, e, en
        d, c, b, a = c, b, a, an
    return out


def _msb_bits(data, nbits):
    bits = [(byte >> (7 - k)) & 1 for byte in data for k in range(8)]
    return bits + [0] * (nbits - len(bits))


def _cases():
    rng = random.Random(5)
    c256 = _sha256_circuit(1)
    msgs = [bytes(rng.randrange(256) for _ in range(L)) for L in (0, 55, 56, 100, 119)]
    yield c256, [({"in": _msb_bits(m, 1024), "len": 8 * len(m)}, hashlib.sha256(m).digest()) for m in msgs]
    c512 = _sha512_circuit()
    msgs = [bytes(rng.randrange(256) for _ in range(64)) for _ in range(2)]
    yield c512, [({"in": _msb_bits(m, 512)}, hashlib.sha512(m).digest()) for m in msgs]


def test_native_steps_reproduce_the_generic_witness():
    for c, cases in _cases():
        art = c.finalize().artifact()
        from nzcb_circom_b200.circom.builder import OP_SHABLOCK, OP_SHAROUNDS
        fused = [(i[0], i[1]) for i in c.prog if i[0] in (OP_SHAROUND, OP_SHASCHED, OP_SHABLOCK, OP_SHAROUNDS)]
        n_round = sum(1 for op, _ in fused if op == OP_SHAROUND) + \
            sum(pl["rounds"] - pl["r_start"] for op, pl in fused if op in (OP_SHABLOCK, OP_SHAROUNDS))
        # all but the four rounds whose state still holds constants of the initial hash value, per hash call
        assert n_round >= (2 * 64 - 4 if "256" in c.name else 80 - 4)
        if "256" in c.name:   # both compressions of the two-block SHA-256 are single instructions
            assert [pl["r_start"] for op, pl in fused if op == OP_SHABLOCK] == [4, 0]
        else:                 # SHA-512's padded block: the schedule keeps its steps, the 76 regular rounds are one instruction
            assert [(pl["n"], pl["r_start"]) for op, pl in fused if op == OP_SHAROUNDS] == [(64, 4)]
        assert art.n_instr_native < art.n_instr // 8 and art.n_levels_native < art.n_levels // 2
        prog = vm.Program(art.wprog_bytes())
        for inp, digest in cases:
            W = vm.run(prog, art.flatten_input(inp), full=True)
            nout = len(digest) * 8
            assert bytes(sum(W[1 + 8 * k + j] << (7 - j) for j in range(8)) for k in range(nout // 8)) == digest
            for op, pl in fused:
                got = (_emulate_block(pl, W, art.n_witness) if op == OP_SHABLOCK else
                       _emulate_rounds(pl, W, art.n_witness) if op == OP_SHAROUNDS else _emulate(op, pl, W, art.n_witness))
                for w, v in got.items():
                    assert W[w] == v, (c.name, op, w)


@pytest.mark.gpu
def test_gpu_native_program_equals_the_generic_witness(ctx):
    from nzcb_circom_b200.circom_tester import WasmTester
    from oracle import c_oracle as C

    for c, cases in _cases():
        art = c.finalize().artifact()
        assert art.wprog_bytes(native=True) != art.wprog_bytes()
        cir = WasmTester(art, ctx)
        raw, st = cir.calculateWitnessBatch([inp for inp, _ in cases], True, ctx)
        assert st == [0] * len(cases)
        for k, (inp, digest) in enumerate(cases):
            flat = b"".join(int(v).to_bytes(32, "little") for v in art.flatten_input(inp))
            rc, wires = C.witness(art.wprog_bytes(), flat, art.n_total)      # the generic program on the C oracle
            assert rc == 0
            assert raw[k * art.n_witness * 32:(k + 1) * art.n_witness * 32] == wires[:art.n_witness * 32]
        cir.close()


# ---- QuinSelector as one instruction (builder.OP_QUINSEL, csrc/witness.cu quinsel_warp) -----------------------------
def _emulate_quinsel(pl, W, nw):
    from nzcb_circom_b200.circom.builder import R as RR

    def val(lc):
        v = lc.k
        for w, cf in lc.t.items():
            v += cf * W[w if w >= 0 else nw + (-w - 1)]
        return v % RR

    index, N = val(pl["index"]), len(pl["ins"])
    valid = index < N
    picked = val(pl["ins"][index]) if valid else 0
    out = {}
    for i in range(N):
        d = (i - index) % RR
        out[pl["eq_w"][i] - 1] = pow(d, RR - 2, RR) if d else 0
        out[pl["eq_w"][i]] = 1 if d == 0 else 0
        if pl["sum_w"][i] is not None:
            out[pl["sum_w"][i]] = picked if valid and i >= index else 0
    return out


def test_native_quinselector_reproduces_the_generic_witness():
    from nzcb_circom_b200.circom.builder import OP_QUINSEL
    from nzcb_circom_b200.circom_tester import MAINS

    rng = random.Random(9)
    for name, make in (("quinSelector5_test", lambda: {"in": [rng.randrange(R) for _ in range(5)], "index": rng.randrange(5)}),
                       ("getV5_test", None), ("skipValue5_test", None), ("readCredSubj_exampleTest", None),
                       ("constructNullifier_test", None)):
        c = Circuit(name)
        MAINS[name](c)
        art = c.finalize().artifact()
        fused = [i[1] for i in c.prog if i[0] == OP_QUINSEL]
        assert fused and art.n_instr_native < art.n_instr
        prog = vm.Program(art.wprog_bytes())
        if make is not None:
            inputs = [make() for _ in range(6)]
        elif name == "getV5_test":
            inputs = [{"bytes": [9, 8, 7, 6, 5], "pos": k} for k in range(5)]
        elif name == "skipValue5_test":
            inputs = [{"bytes": [0x83, 23, 23, 23, 0], "pos": 0}, {"bytes": [0x62, 65, 66, 0, 0], "pos": 0}]
        elif name == "constructNullifier_test":   # selectors over buffers with constant (zero) tails, signal indices
            from nzcb_circom_b200 import nzcp_helpers as H
            inputs = [{"givenName": H.padArray(H.stringToArray(g), 64), "givenNameLen": len(g),
                       "familyName": H.padArray(H.stringToArray(f), 64), "familyNameLen": len(f),
                       "dob": H.padArray(H.stringToArray("1960-04-16"), 64), "dobLen": 10}
                      for g, f in (("Jack", "Sparrow"), ("A", "B" * 21))]
        else:
            from nzcb_circom_b200 import nzcp_helpers as H
            cose = H.getCOSE(H.EXAMPLE_PASS_URI)
            tbs = H.encodeToBeSigned(cose["bodyProtected"], cose["payload"])
            inputs = [{"mapLen": 3, "bytes": list(H.fitBytes(tbs, 314)), "pos": 247}]
        for inp in inputs:
            W = vm.run(prog, art.flatten_input(inp), full=True)
            for pl in fused:
                for w, v in _emulate_quinsel(pl, W, art.n_witness).items():
                    assert W[w] == v, (name, pl["w0"], w)


@pytest.mark.gpu
def test_gpu_native_quinselector_out_of_range_index_is_rejected_not_miscomputed(ctx):
    """index >= choices: the LessThan assert rejects the pass; in range: every wire equals the C oracle's"""
    from nzcb_circom_b200.circom_tester import wasm_tester
    from oracle import c_oracle as C

    cir = wasm_tester("quinSelector5_test", ctx)
    art = cir.compiled
    assert art.wprog_bytes(native=True) != art.wprog_bytes()
    inputs = [{"in": [3, 1, 4, 1, 5], "index": k} for k in range(5)] + [{"in": [3, 1, 4, 1, 5], "index": 7}]
    raw, st = cir.calculateWitnessBatch(inputs, True, ctx)
    assert st == [0] * 5 + [-6]
    for k in range(5):
        flat = b"".join(int(v).to_bytes(32, "little") for v in art.flatten_input(inputs[k]))
        rc, wires = C.witness(art.wprog_bytes(), flat, art.n_total)
        assert rc == 0 and raw[k * art.n_witness * 32:(k + 1) * art.n_witness * 32] == wires[:art.n_witness * 32]
