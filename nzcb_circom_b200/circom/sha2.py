"""SHA-256 (variable length) and SHA-512 gadgets.

The reference pulls these from two *unpinned* GitHub zips at build time
(noway/sha256-var-circom@main and Electron-Labs/sha512@master,
/root/reference/Makefile:20-43) which are not in /root/reference; only their call
sites are (circuits/nzcptpl.circom:509-516 ``Sha256Var(3)``, :577-580
``Sha512(512)``).  SURVEY.md A.6 fixes the semantics: bit inputs MSB-first per
byte, standard FIPS 180-4 padding, done in-circuit for the variable length.  The
gadgets below are restated in the circomlib style (Xor3 / Ch / Maj per bit,
BinSum for the modular additions) so constraint counts are comparable with the
real thing; the round constants are derived from first principles (fractional
parts of square / cube roots of primes) and pinned by the hashlib cross-checks of
tests/test_nzcp.py (the digests of real and synthetic passes) and tests/test_sha_native.py."""
from math import isqrt

from .builder import LC, OP_SHABLOCK, OP_SHAROUND, OP_SHAROUNDS, OP_SHASCHED, Circuit
from .circomlib import is_equal, num2bits


def _primes(n):
    out = []
    k = 2
    while len(out) < n:
        if all(k % p for p in out if p * p <= k):
            out.append(k)
        k += 1
    return out


def _icbrt(x):
    lo, hi = 0, 1 << ((x.bit_length() + 2) // 3 + 1)
    while lo < hi:
        mid = (lo + hi + 1) // 2
        if mid ** 3 <= x:
            lo = mid
        else:
            hi = mid - 1
    return lo


def _frac_sqrt(p, bits):
    return isqrt(p << (2 * bits)) & ((1 << bits) - 1)


def _frac_cbrt(p, bits):
    return _icbrt(p << (3 * bits)) & ((1 << bits) - 1)


SHA256_H = [_frac_sqrt(p, 32) for p in _primes(8)]
SHA256_K = [_frac_cbrt(p, 32) for p in _primes(64)]
SHA512_H = [_frac_sqrt(p, 64) for p in _primes(8)]
SHA512_K = [_frac_cbrt(p, 64) for p in _primes(80)]
assert SHA256_H[0] == 0x6A09E667 and SHA256_K[63] == 0xC67178F2
assert SHA512_H[0] == 0x6A09E667F3BCC908 and SHA512_K[79] == 0x6C44198C4A475817


# words are lists of LCs, index 0 = least significant bit (circomlib convention)
def _const_word(v, n):
    return [LC(None, (v >> i) & 1) for i in range(n)]


def _rotr(w, r):
    n = len(w)
    return [w[(i + r) % n] for i in range(n)]


def _shr(w, r):
    n = len(w)
    return [w[i + r] if i + r < n else LC() for i in range(n)]


def _xor3(c: Circuit, a, b, cc):
    """circomlib sha256/xor3.circom: mid <== b*c; out <== a*(1 - 2b - 2c + 4mid) + b + c - 2mid"""
    mid = c.mul(b, cc)
    return c.quad(a * (1 - b * 2 - cc * 2 + mid * 4) + b + cc - mid * 2)


def _ch(c, a, b, cc):
    """ch.circom: out <== a*(b - c) + c"""
    return c.quad(a * (b - cc) + cc)


def _maj(c, a, b, cc):
    """maj.circom: mid <== b*c; out <== a*(b + c - 2mid) + mid"""
    mid = c.mul(b, cc)
    return c.quad(a * (b + cc - mid * 2) + mid)


def _binsum(c: Circuit, n, ops):
    """circomlib binsum.circom: sum of words mod 2^n.  out bits are hints, each boolean, and
    their weighted sum is constrained to the linear sum of the operands."""
    lin = LC()
    for w in ops:
        for i, b in enumerate(w):
            lin = lin + b * (1 << i)
    if lin.is_const():
        return _const_word(lin.k & ((1 << n) - 1), n)
    nout = ((len(ops) * ((1 << n) - 1))).bit_length()
    bits = c.hint_bits(lin, nout)
    acc = LC()
    for i, b in enumerate(bits):
        c.assert_zero(b * (b - 1), implied=True)  # b was just written by the decomposition
        acc = acc + b * (1 << i)
    c.assert_eq(acc, lin)
    return bits[:n]


class _Sha2:
    def __init__(self, n, rounds, K, big0, big1, small0, small1):
        self.n, self.rounds, self.K = n, rounds, K
        self.big0, self.big1, self.small0, self.small1 = big0, big1, small0, small1

    def _sigma(self, c, x, rot):
        r1, r2, r3, shift = rot
        third = _shr(x, r3) if shift else _rotr(x, r3)
        a, b = _rotr(x, r1), _rotr(x, r2)
        return [_xor3(c, a[i], b[i], third[i]) for i in range(self.n)]

    def compress(self, c: Circuit, state, block_words):
        """state: 8 words; block_words: 16 words (LSB-first bit lists).  Returns the new state.

        Every message-schedule step and every round whose wires come out in the regular layout (no constant
        folding: all state bits are signals) is ALSO recorded as one word-level instruction of the native witness
        program (builder.OP_SHASCHED / OP_SHAROUND): the same wires, computed on n-bit words."""
        n = self.n
        w = list(block_words)
        r3_1, r3_0 = self.small1[2], self.small0[2]
        p_block = len(c.prog)
        sched_w0, round_w0, round_state = {}, {}, {}
        for t in range(16, self.rounds):
            p0, w0 = len(c.prog), c.n_wires
            s1 = self._sigma(c, w[t - 2], self.small1)
            s0 = self._sigma(c, w[t - 15], self.small0)
            wt = _binsum(c, n, [s1, w[t - 7], s0, w[t - 16]])
            w.append(wt)
            # layout of a regular step: sigma1 (mid, out interleaved; out only where the shifted operand is 0),
            # sigma0 likewise, then the n + 2 sum bits
            o_s0 = 2 * n - r3_1
            o_w = o_s0 + 2 * n - r3_0
            size = o_w + n + 2

            def sig_out(off, r3, i):
                return w0 + off + (2 * i + 1 if i < n - r3 else 2 * (n - r3) + i - (n - r3))

            regular = (c.n_wires - w0 == size and
                       all(s1[i].single_wire() == sig_out(0, r3_1, i) for i in range(n)) and
                       all(s0[i].single_wire() == sig_out(o_s0, r3_0, i) for i in range(n)) and
                       all(wt[i].single_wire() == w0 + o_w + i for i in range(n)))
            if regular:
                c.fuse(p0, OP_SHASCHED, {"n": n, "rot1": list(self.small1[:3]), "rot0": list(self.small0[:3]), "w0": w0,
                                         "size": size, "words": [w[t - 2], w[t - 7], w[t - 15], w[t - 16]]})
                sched_w0[t] = (w0, size)
        a, b, cc, d, e, f, g, h = state
        for t in range(self.rounds):
            p0, w0 = len(c.prog), c.n_wires
            ins = [a, b, cc, d, e, f, g, h, w[t]]
            S1 = self._sigma(c, e, self.big1)
            chv = [_ch(c, e[i], f[i], g[i]) for i in range(n)]
            t1 = _binsum(c, n, [h, S1, chv, _const_word(self.K[t], n), w[t]])
            S0 = self._sigma(c, a, self.big0)
            mj = [_maj(c, a[i], b[i], cc[i]) for i in range(n)]
            t2 = _binsum(c, n, [S0, mj])
            h, g, f = g, f, e
            e = _binsum(c, n, [d, t1])
            d, cc, b = cc, b, a
            a = _binsum(c, n, [t1, t2])
            # regular layout: S1 2n | ch n | t1 n+3 | S0 2n | maj 2n | t2 n+1 | e n+1 | a n+1   (11 n + 6 wires)
            o_ch, o_t1, o_S0, o_mj, o_t2, o_e, o_a = 2 * n, 3 * n, 4 * n + 3, 6 * n + 3, 8 * n + 3, 9 * n + 4, 10 * n + 5
            size = 11 * n + 6
            regular = (c.n_wires - w0 == size and
                       all(x.single_wire() is not None for word in ins[:8] for x in word) and
                       all(S1[i].single_wire() == w0 + 2 * i + 1 and chv[i].single_wire() == w0 + o_ch + i and
                           t1[i].single_wire() == w0 + o_t1 + i and S0[i].single_wire() == w0 + o_S0 + 2 * i + 1 and
                           mj[i].single_wire() == w0 + o_mj + 2 * i + 1 and t2[i].single_wire() == w0 + o_t2 + i and
                           e[i].single_wire() == w0 + o_e + i and a[i].single_wire() == w0 + o_a + i for i in range(n)))
            if regular:
                c.fuse(p0, OP_SHAROUND, {"n": n, "rot1": list(self.big1[:3]), "rot0": list(self.big0[:3]), "K": self.K[t],
                                         "w0": w0, "size": size, "words": ins})
                round_w0[t] = (w0, size)
                round_state[t] = ins[:8]
        # A 32-bit compression whose every schedule step is regular and whose rounds are regular from some round on
        # becomes ONE instruction of the native program (state in registers across the rounds): the step instructions
        # recorded above, from that round on, are folded into it.  (SHA-512's padded block has constant message words,
        # so its schedule is not regular: it keeps one instruction per step.)
        r_start = next((t for t in range(self.rounds) if all(u in round_w0 for u in range(t, self.rounds))), None)
        if n == 32 and r_start is not None and r_start <= 16 and len(sched_w0) == self.rounds - 16:
            keep, folded = [], []
            for ins_ in c.prog[p_block:]:
                if ins_[0] == OP_SHASCHED or (ins_[0] == OP_SHAROUND and ins_[1]["w0"] >= round_w0[r_start][0]):
                    folded.append(ins_)
                else:
                    keep.append(ins_)
            generic = [g_ for f_ in folded for g_ in f_[2]]
            del c.prog[p_block:]
            c.prog.extend(keep)
            c.prog.append((OP_SHABLOCK, {
                "n": n, "rot1": list(self.big1[:3]), "rot0": list(self.big0[:3]), "srot1": list(self.small1[:3]),
                "srot0": list(self.small0[:3]), "r_start": r_start, "rounds": self.rounds,
                "sched_w0": [sched_w0[t][0] for t in range(16, self.rounds)],
                "round_w0": [round_w0[t][0] for t in range(r_start, self.rounds)],
                "K": [self.K[t] for t in range(r_start, self.rounds)],
                "regions": [sched_w0[t] for t in range(16, self.rounds)] + [round_w0[t] for t in range(r_start, self.rounds)],
                "words": list(round_state[r_start]) + list(block_words)}, generic))
        elif r_start is not None and self.rounds - r_start >= 8:
            # the schedule is not regular (constant message bits) but the rounds are: ONE instruction for all of them,
            # w[t] read from its wires (or constants) round by round; the schedule keeps its own instructions
            keep, folded = [], []
            for ins_ in c.prog[p_block:]:
                if ins_[0] == OP_SHAROUND and ins_[1]["w0"] >= round_w0[r_start][0]:
                    folded.append(ins_)
                else:
                    keep.append(ins_)
            generic = [g_ for f_ in folded for g_ in f_[2]]
            del c.prog[p_block:]
            c.prog.extend(keep)
            c.prog.append((OP_SHAROUNDS, {
                "n": n, "rot1": list(self.big1[:3]), "rot0": list(self.big0[:3]), "r_start": r_start, "rounds": self.rounds,
                "round_w0": [round_w0[t][0] for t in range(r_start, self.rounds)],
                "K": [self.K[t] for t in range(r_start, self.rounds)],
                "regions": [round_w0[t] for t in range(r_start, self.rounds)],
                "words": list(round_state[r_start]) + [w[t] for t in range(r_start, self.rounds)]}, generic))
        new = [a, b, cc, d, e, f, g, h]
        return [_binsum(c, n, [state[i], new[i]]) for i in range(8)]


_SHA256 = _Sha2(32, 64, SHA256_K, (2, 13, 22, False), (6, 11, 25, False), (7, 18, 3, True), (17, 19, 10, True))
_SHA512 = _Sha2(64, 80, SHA512_K, (28, 34, 39, False), (14, 18, 41, False), (1, 8, 7, True), (19, 61, 6, True))


def _words_from_msb_bits(bits, n):
    """message bits, MSB-first, -> words as LSB-first bit lists"""
    return [list(reversed(bits[k * n:(k + 1) * n])) for k in range(len(bits) // n)]


def sha256_var(c: Circuit, in_bits, length_bits, block_space):
    """Sha256Var(BlockSpace) -- call site nzcptpl.circom:509-516.  in_bits: 512 * 2^BlockSpace
    message bits (MSB-first per byte), length_bits: message length in bits (a multiple of 8,
    at most 512 * 2^BlockSpace - 72).  Returns the 256 digest bits, MSB first.

    Padding is done in-circuit at byte granularity: L = len / 8; eq_k = (k == L) marks the
    0x80 byte, lt_k = 1 - sum_{j<=k} eq_j masks the message; the 64-bit length goes into the
    last 8 bytes of block nb = (L + 8) >> 6, and the digest is the chaining value after that
    block, selected out of the 2^BlockSpace candidates."""
    nblocks = 1 << block_space
    nbytes = 64 * nblocks
    assert len(in_bits) == 8 * nbytes
    lbits_n = (8 * nbytes).bit_length()  # enough for any length in range
    length_bits = LC.of(length_bits)
    lb = num2bits(c, length_bits, lbits_n)
    for i in range(3):
        c.assert_zero(lb[i])  # whole bytes only
    L = LC()
    for i in range(3, lbits_n):
        L = L + lb[i] * (1 << (i - 3))
    # nb = (L + 8) >> 6, must fit the block space
    lp = num2bits(c, L + 8, 6 + block_space)
    nb = LC()
    for i in range(block_space):
        nb = nb + lp[6 + i] * (1 << i)
    nb = c.wire(nb)
    sel = [is_equal(c, nb, j) for j in range(nblocks)]
    # per byte masks.  lt_k is bound to its own signal (lt_k = lt_{k-1} - eq_k) so the eight
    # products per byte see a one-term factor instead of a k-term prefix sum
    msg = []
    lt = LC(None, 1)
    L_w = c.wire(L)
    # lt_k <== lt_(k-1) - eq_k is a chain of nbytes dependent signals.  The R1CS keeps that form; the witness
    # program computes lt_k = 1 - (eq's of the earlier blocks of BLK bytes) - (eq's of this block up to k): three
    # levels whatever the length (the chain was hidden behind the hash rounds until those became one instruction)
    BLK = 23
    prefix, block_eqs = LC(), LC()
    for k in range(nbytes):
        eq = is_equal(c, k, L_w)
        if k % BLK == 0 and k:
            prefix = c.temp(prefix + block_eqs) if not (prefix + block_eqs).is_const() else prefix + block_eqs
            block_eqs = LC()
        block_eqs = block_eqs + eq
        lt = c.wire_as(lt - eq, 1 - prefix - block_eqs)
        for bpos in range(8):
            m = c.mul(in_bits[8 * k + bpos], lt)
            if bpos == 0:
                m = m + eq
            msg.append(m)
    c.assert_zero(lt)  # the 0x80 byte exists: L < nbytes
    # 64-bit big-endian bit length in the last 8 bytes of block nb (those bytes are otherwise 0)
    for j in range(nblocks):
        base = 512 * (j + 1) - 64
        for i in range(lbits_n):
            pos = base + 63 - i
            msg[pos] = msg[pos] + c.mul(sel[j], lb[i])
    state = [_const_word(h, 32) for h in SHA256_H]
    outs = []
    for j in range(nblocks):
        state = _SHA256.compress(c, state, _words_from_msb_bits(msg[512 * j:512 * (j + 1)], 32))
        outs.append(state)
    digest = []
    for wi in range(8):
        for bit in range(31, -1, -1):
            acc = LC()
            for j in range(nblocks):
                acc = c.quad(sel[j] * outs[j][wi][bit] + acc)
            digest.append(acc)
    return digest


def sha512_fixed(c: Circuit, in_bits):
    """Sha512(nBits) -- call site nzcptpl.circom:577-580, nBits a compile-time constant.
    Padding is constant; returns 512 digest bits, MSB first."""
    nbits = len(in_bits)
    total = ((nbits + 1 + 128 + 1023) // 1024) * 1024
    msg = [LC.of(b) for b in in_bits] + [LC(None, 1)] + [LC()] * (total - nbits - 1 - 128)
    msg += [LC(None, (nbits >> (127 - i)) & 1) for i in range(128)]
    state = [_const_word(h, 64) for h in SHA512_H]
    for j in range(total // 1024):
        state = _SHA512.compress(c, state, _words_from_msb_bits(msg[1024 * j:1024 * (j + 1)], 64))
    return [state[wi][bit] for wi in range(8) for bit in range(63, -1, -1)]
