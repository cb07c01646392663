"""Circuit builder: the role `circom --r1cs --wasm --O2` plays for the reference
(/root/reference/Makefile:5-15).  circom itself (a Rust binary) is absent from
this image and so are the SHA gadgets it would compile, so
the templates are restated on this small eDSL, which emits

  * the R1CS (``.r1cs`` bytes) that `plonk setup` consumes, and
  * a *witness program*: one instruction per wire (the straight-line program a
    circom-generated WASM runs), level-scheduled so the GPU interpreter
    (csrc/witness.cu) can run every level in parallel.

Semantics follow circom: ``<==`` is "assign and constrain", ``<--`` is a hint
(BITS / INV instructions), ``===`` adds a constraint plus a run-time assert.
Like ``--O2`` the builder never materialises a signal that is a linear
combination of others: linear expressions are carried symbolically (class LC)
and only quadratic assignments, hints, inputs and outputs become wires.  Wire
order is circom's convention where it is observable: 1, outputs, inputs in
declaration order (test/nzcp.js:44, test/cbor.js:191-193), then internals.
"""
import os
import struct

R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001

OP_LIN, OP_MUL, OP_BITS, OP_INV, OP_ASSERT, OP_BITSLC = 1, 2, 3, 4, 5, 6
# Word-level ("fat") instructions of the NATIVE witness program (csrc/witness.cu): one SHA-2 round / one message
# schedule step computed on 32- or 64-bit words, every bit wire of the step written out afterwards.  They replace
# the generic instructions the same step emits (kept alongside: the oracle's program is the generic one, so the
# oracle VM checks the native semantics wire for wire).  Signals, numbering and R1CS are untouched.
OP_SHAROUND, OP_SHASCHED = 7, 8
# QuinSelector(N) (circuits/quinSelector.circom:26-41) as one instruction: the N IsZero inverses, the N equality flags
# and the N running sums written from the index and the one selected input
OP_QUINSEL = 9
# a whole SHA-256 compression from round r_start on (message schedule included): the state stays in registers across
# the rounds, one level instead of one per round.  Stands for the OP_SHASCHED / OP_SHAROUND steps it is made of.
OP_SHABLOCK = 10
# the rounds of a compression from round r_start on, with the message-schedule words read from their wires (for
# compressions whose schedule is not regular: constant message bits fold some of its steps away)
OP_SHAROUNDS = 11
FAT_OPS = (OP_SHAROUND, OP_SHASCHED, OP_QUINSEL, OP_SHABLOCK, OP_SHAROUNDS)


class LC:
    """Linear combination  k + sum coef * wire.  Immutable by convention."""
    __slots__ = ("t", "k")

    def __init__(self, t=None, k=0):
        self.t = t if t is not None else {}
        self.k = k % R

    @staticmethod
    def of(x):
        if isinstance(x, LC):
            return x
        if isinstance(x, int):
            return LC(None, x)
        raise TypeError(f"cannot use {type(x)} as a linear combination")

    def is_const(self):
        return not self.t

    def single_wire(self):
        """wire id if this is exactly 1 * wire, else None"""
        if self.k == 0 and len(self.t) == 1:
            (w, c), = self.t.items()
            if c == 1:
                return w
        return None

    def __add__(self, o):
        if isinstance(o, Quad):
            return o + self
        o = LC.of(o)
        if not o.t:
            return LC(self.t, self.k + o.k)
        if not self.t:
            return LC(o.t, self.k + o.k)
        t = dict(self.t)
        for w, c in o.t.items():
            v = (t.get(w, 0) + c) % R
            if v:
                t[w] = v
            else:
                t.pop(w, None)
        return LC(t, self.k + o.k)

    __radd__ = __add__

    def __neg__(self):
        return LC({w: R - c for w, c in self.t.items()}, -self.k)

    def __sub__(self, o):
        if isinstance(o, Quad):
            return (-o) + self
        return self + (-LC.of(o))

    def __rsub__(self, o):
        return LC.of(o) + (-self)

    def __mul__(self, o):
        if isinstance(o, int):
            o %= R
            if o == 0:
                return LC()
            if o == 1:
                return self
            return LC({w: c * o % R for w, c in self.t.items()}, self.k * o)
        if isinstance(o, LC):
            if not o.t:
                return self * o.k
            if not self.t:
                return o * self.k
            return Quad(self, o, LC())
        return NotImplemented

    __rmul__ = __mul__


class Quad:
    """a * b + c with a, b, c linear -- what one R1CS constraint can express."""
    __slots__ = ("a", "b", "c")

    def __init__(self, a, b, c):
        self.a, self.b, self.c = a, b, c

    def __add__(self, o):
        if isinstance(o, Quad):
            raise ValueError("non-quadratic expression: sum of two products")
        return Quad(self.a, self.b, self.c + o)

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, Quad):
            raise ValueError("non-quadratic expression: difference of two products")
        return Quad(self.a, self.b, self.c - o)

    def __rsub__(self, o):
        return (-self) + o

    def __neg__(self):
        return Quad(-self.a, self.b, -self.c)

    def __mul__(self, o):
        if isinstance(o, int):
            return Quad(self.a * o, self.b, self.c * o)
        if isinstance(o, LC) and o.is_const():
            return self * o.k
        raise ValueError("non-quadratic expression: product of degree > 2")

    __rmul__ = __mul__


def _terms(lc: LC):
    """constraint-side representation: dict wire -> coef with the constant under wire 0"""
    d = dict(lc.t)
    if lc.k:
        d[0] = lc.k
    return d


class Circuit:
    def __init__(self, name):
        self.name = name
        self.n_wires = 1          # wire 0 == 1
        self.n_temps = 0
        self.outputs = []         # (name, dims, first wire)
        self.inputs = []
        self.n_out = 0
        self.n_in = 0
        self._io_closed = False
        self.constraints = []     # (A, B, C) dicts
        self.prog = []            # instruction tuples
        self.output_assigned = set()
        self.drop_implied = os.environ.get("NZCB_WITNESS_DROP_IMPLIED", "1") != "0"  # see assert_zero

    # ---- declaration -------------------------------------------------
    @staticmethod
    def _size(dims):
        n = 1
        for d in dims:
            n *= d
        return n

    def output(self, name, *dims):
        assert not self.inputs and not self._io_closed, "outputs come first in the witness"
        n = self._size(dims)
        first = self.n_wires
        self.n_wires += n
        self.n_out += n
        self.outputs.append((name, dims, first))
        return self._shape([LC({first + i: 1}) for i in range(n)], dims)

    def input(self, name, *dims):
        assert not self._io_closed
        n = self._size(dims)
        first = self.n_wires
        self.n_wires += n
        self.n_in += n
        self.inputs.append((name, dims, first))
        return self._shape([LC({first + i: 1}) for i in range(n)], dims)

    @staticmethod
    def _shape(flat, dims):
        if not dims:
            return flat[0]
        if len(dims) == 1:
            return flat
        step = len(flat) // dims[0]
        return [Circuit._shape(flat[i * step:(i + 1) * step], dims[1:]) for i in range(dims[0])]

    # ---- wires ---------------------------------------------------------
    def _new_wire(self):
        self._io_closed = True
        w = self.n_wires
        self.n_wires += 1
        return w

    def _new_temp(self):
        self.n_temps += 1
        return -self.n_temps  # remapped behind the witness wires at finalisation

    def mul(self, a, b):
        """signal s; s <== a * b"""
        return self.quad(LC.of(a) * LC.of(b))

    def quad(self, q):
        """signal s; s <== a*b + c   (one constraint).  Linear input: returned as is."""
        if isinstance(q, (LC, int)):
            return LC.of(q)
        if q.a.is_const():
            return q.b * q.a.k + q.c
        if q.b.is_const():
            return q.a * q.b.k + q.c
        w = self._new_wire()
        self.prog.append((OP_MUL, w, q.a, q.b, q.c))
        self.constraints.append((_terms(q.a), _terms(q.b), _terms(LC({w: 1}) - q.c)))
        return LC({w: 1})

    def chain(self):
        """Running sum  s_i <== a_i * b_i + s_(i-1)  (s_(-1) = 0): see MulSumChain."""
        return MulSumChain(self)

    def wire(self, x):
        """force a real signal equal to x (s <== x)"""
        if isinstance(x, Quad):
            return self.quad(x)
        x = LC.of(x)
        if x.single_wire() is not None:
            return x
        w = self._new_wire()
        self.prog.append((OP_LIN, w, x))
        self.constraints.append(({}, {}, _terms(x - LC({w: 1}))))
        return LC({w: 1})

    def wire_as(self, x, prog_lc):
        """signal s; s <== x, exactly like wire(x) for the R1CS and the numbering -- but the witness PROGRAM computes
        s from prog_lc, a linear combination of equal value with a shorter dependency chain (e.g. a prefix sum
        spelled out instead of the previous element of a running difference)."""
        x = LC.of(x)
        w = self._new_wire()
        self.prog.append((OP_LIN, w, LC.of(prog_lc)))
        self.constraints.append(({}, {}, _terms(x - LC({w: 1}))))
        return LC({w: 1})

    def temp(self, x):
        """program temp (not a signal, no constraint) holding the value of the linear combination x"""
        t = self._new_temp()
        self.prog.append((OP_LIN, t, LC.of(x)))
        return LC({t: 1})

    def assign_output(self, out_lc, x):
        """out <== x for a declared main output"""
        w = out_lc.single_wire()
        assert w is not None and 1 <= w <= self.n_out and w not in self.output_assigned
        self.output_assigned.add(w)
        self._io_closed = True
        if isinstance(x, Quad) and not (x.a.is_const() or x.b.is_const()):
            self.prog.append((OP_MUL, w, x.a, x.b, x.c))
            self.constraints.append((_terms(x.a), _terms(x.b), _terms(out_lc - x.c)))
            return
        if isinstance(x, Quad):
            x = x.a * x.b + x.c
        x = LC.of(x)
        self.prog.append((OP_LIN, w, x))
        self.constraints.append(({}, {}, _terms(x - out_lc)))

    def _src(self, x):
        """single id (wire or temp) holding the value of x, for hint instructions"""
        x = LC.of(x)
        w = x.single_wire()
        if w is not None:
            return w
        t = self._new_temp()
        self.prog.append((OP_LIN, t, x))
        return t

    def hint_bits(self, x, n):
        """out[i] <-- (x >> i) & 1, i < n   (no constraints)"""
        self._io_closed = True
        x = LC.of(x)
        first = self.n_wires
        self.n_wires += n
        w = x.single_wire()
        if w is not None:
            self.prog.append((OP_BITS, first, w, n))
        else:  # the sum is evaluated and decomposed by ONE instruction: no temp, no extra dependency level
            self.prog.append((OP_BITSLC, first, n, x))
        return [LC({first + i: 1}) for i in range(n)]

    def hint_inv(self, x):
        """inv <-- x != 0 ? 1/x : 0"""
        src = self._src(x)
        w = self._new_wire()
        self.prog.append((OP_INV, w, src))
        return LC({w: 1})

    # ---- constraints ---------------------------------------------------
    def assert_zero(self, x, implied=False):
        """x === 0 (constraint + run-time assert, circom_runtime error 4 "Assert Failed").
        implied=True marks a check that holds by construction of the witness program itself -- the booleanity of
        bits a decomposition instruction has just written, `in * out === 0` of IsZero after its inverse hint: it
        stays a constraint of the R1CS and costs no run-time instruction (it can never fire).  On by default since the
        slimmer program went through the whole GPU parity suite (round 2: 139 tests, witnesses == C oracle);
        NZCB_WITNESS_DROP_IMPLIED=0 emits the checks again."""
        if isinstance(x, Quad) and not (x.a.is_const() or x.b.is_const()):
            a, b, c = x.a, x.b, -x.c
        else:
            if isinstance(x, Quad):
                x = x.a * x.b + x.c
            x = LC.of(x)
            if x.is_const():
                if x.k != 0:
                    raise ValueError("constraint is never satisfiable")
                return
            a, b, c = LC(), LC(), x
        if not (implied and self.drop_implied):
            self.prog.append((OP_ASSERT, a, b, c))
        self.constraints.append((_terms(a), _terms(b), _terms(c)))

    def assert_eq(self, x, y):
        self.assert_zero(x - y)

    def fuse(self, p0, op, payload):
        """the instructions emitted since prog index p0 can be replaced, in the native program, by ONE word-level
        instruction (op, payload); the generic ones are kept for the generic (oracle) program"""
        generic = self.prog[p0:]
        del self.prog[p0:]
        self.prog.append((op, payload, generic))

    # ---- outputs -------------------------------------------------------
    def finalize(self):
        assert len(self.output_assigned) == self.n_out, f"{self.name}: unassigned main outputs"
        return Compiled(self)


class MulSumChain:
    """s_i <== a_i * b_i + s_(i-1), the running sum of quinSelector.circom:37.  step() allocates the signal and the
    constraint of one step exactly where a plain ``quad`` would (same wire numbering, same R1CS); only the witness
    PROGRAM differs: finish() emits products -> block totals -> block prefixes -> prefix sums, four dependency levels
    whatever the length, instead of one level per step (351 for GetV(351), the bulk of the circuit's critical path).
    The intermediate values live in program temps, which are not part of the witness."""

    def __init__(self, c):
        self.c = c
        self.steps = []      # (wire, a, b)
        self.acc = LC()
        self.plain = False   # a degenerate (linear) step was met: fall back to one instruction per step

    def step(self, a, b):
        c = self.c
        a, b = LC.of(a), LC.of(b)
        if self.plain or a.is_const() or b.is_const():
            if not self.plain:
                self._emit_plain()
                self.plain = True
            self.acc = c.quad(a * b + self.acc)
            return self.acc
        w = c._new_wire()
        c.constraints.append((_terms(a), _terms(b), _terms(LC({w: 1}) - self.acc)))
        self.steps.append((w, a, b))
        self.acc = LC({w: 1})
        return self.acc

    def _emit_plain(self):
        prev = LC()
        for w, a, b in self.steps:
            self.c.prog.append((OP_MUL, w, a, b, prev))
            prev = LC({w: 1})
        self.steps = []

    def finish(self):
        c = self.c
        n = len(self.steps)
        if self.plain or n <= 8:
            self._emit_plain()
            return self.acc
        prods = []
        for _, a, b in self.steps:
            t = c._new_temp()
            c.prog.append((OP_MUL, t, a, b, LC()))
            prods.append(t)
        blk = 1
        while blk * blk < n:
            blk += 1
        totals = []      # block totals
        for k in range(0, n, blk):
            t = c._new_temp()
            c.prog.append((OP_LIN, t, LC({p: 1 for p in prods[k:k + blk]})))
            totals.append(t)
        prefixes = [None]  # prefixes[k] = sum of the totals of blocks < k
        for k in range(1, len(totals)):
            t = c._new_temp()
            c.prog.append((OP_LIN, t, LC({x: 1 for x in totals[:k]})))
            prefixes.append(t)
        for i, (w, _, _) in enumerate(self.steps):
            k = i // blk
            terms = {p: 1 for p in prods[k * blk:i + 1]}
            if prefixes[k] is not None:
                terms[prefixes[k]] = 1
            c.prog.append((OP_LIN, w, LC(terms)))
        self.steps = []
        return self.acc


class Compiled:
    """Finalised circuit: R1CS + witness program with temps remapped and levels computed.  native=True keeps the
    word-level instructions (the GPU's program); native=False expands them into the generic instructions they stand
    for (the oracle's program).  Both write the same wires."""

    def __init__(self, c: Circuit, native=False):
        self.native = native
        self._c = c
        self.name = c.name
        self.n_witness = c.n_wires
        self.n_total = c.n_wires + c.n_temps
        self.n_out, self.n_in = c.n_out, c.n_in
        self.outputs, self.inputs = c.outputs, c.inputs
        self.constraints = c.constraints
        nw = c.n_wires

        def rid(w):
            return w if w >= 0 else nw + (-w - 1)

        def rlc(lc):
            if any(w < 0 for w in lc.t):
                return LC({rid(w): v for w, v in lc.t.items()}, lc.k)
            return lc

        level = [0] * self.n_total
        prog = []
        lv = []
        flat = []
        for ins in c.prog:
            if ins[0] in FAT_OPS and not native:
                flat.extend(ins[2])
            else:
                flat.append(ins)
        for ins in flat:
            op = ins[0]
            if op in (OP_SHAROUND, OP_SHASCHED):
                pl = dict(ins[1])
                pl["words"] = [[rlc(b) for b in word] for word in pl["words"]]
                l = 1 + max((level[w] for word in pl["words"] for b in word for w in b.t), default=0)
                for w in range(pl["w0"], pl["w0"] + pl["size"]):
                    level[w] = l
                prog.append((op, pl))
            elif op in (OP_SHABLOCK, OP_SHAROUNDS):
                pl = dict(ins[1])
                pl["words"] = [[rlc(b) for b in word] for word in pl["words"]]
                l = 1 + max((level[w] for word in pl["words"] for b in word for w in b.t), default=0)
                for w0, size in pl["regions"]:
                    for w in range(w0, w0 + size):
                        level[w] = l
                prog.append((op, pl))
            elif op == OP_QUINSEL:
                pl = dict(ins[1])
                pl["index"] = rlc(pl["index"])
                pl["ins"] = [rlc(b) for b in pl["ins"]]
                l = 1 + max((level[w] for b in pl["ins"] + [pl["index"]] for w in b.t), default=0)
                for w in pl["eq_w"] + [x - 1 for x in pl["eq_w"]] + [x for x in pl["sum_w"] if x is not None]:
                    level[w] = l
                prog.append((op, pl))
            elif op == OP_LIN:
                lc = rlc(ins[2])
                dst = rid(ins[1])
                l = 1 + max((level[w] for w in lc.t), default=0)
                level[dst] = l
                prog.append((op, dst, lc))
            elif op == OP_MUL:
                a, b, cc = rlc(ins[2]), rlc(ins[3]), rlc(ins[4])
                dst = rid(ins[1])
                l = 1 + max((level[w] for lc in (a, b, cc) for w in lc.t), default=0)
                level[dst] = l
                prog.append((op, dst, a, b, cc))
            elif op == OP_BITS:
                src = rid(ins[2])
                l = 1 + level[src]
                for i in range(ins[3]):
                    level[ins[1] + i] = l
                prog.append((op, ins[1], src, ins[3]))
            elif op == OP_INV:
                src = rid(ins[2])
                l = 1 + level[src]
                level[ins[1]] = l
                prog.append((op, ins[1], src))
            elif op == OP_BITSLC:
                lc = rlc(ins[3])
                l = 1 + max((level[w] for w in lc.t), default=0)
                for i in range(ins[2]):
                    level[ins[1] + i] = l
                prog.append((op, ins[1], ins[2], lc))
            else:
                a, b, cc = rlc(ins[1]), rlc(ins[2]), rlc(ins[3])
                l = 1 + max((level[w] for lc in (a, b, cc) for w in lc.t), default=0)
                prog.append((op, a, b, cc))
            lv.append(l)
        order = sorted(range(len(prog)), key=lambda i: lv[i])
        self.prog = [prog[i] for i in order]
        self.levels = [lv[i] for i in order]
        self.n_levels = max(lv, default=0)

    # ---- .r1cs ------------------------------------------
    def r1cs_bytes(self):
        body = bytearray()
        pack_i = struct.Struct("<I").pack
        cache = {}

        def coef(c):
            v = cache.get(c)
            if v is None:
                v = cache[c] = c.to_bytes(32, "little")
            return v

        for lcs in self.constraints:
            for lc in lcs:
                items = sorted(lc.items())
                body += pack_i(len(items))
                for w, c in items:
                    body += pack_i(w)
                    body += coef(c)
        hdr = struct.pack("<I", 32) + R.to_bytes(32, "little")
        hdr += struct.pack("<IIIIQI", self.n_witness, self.n_out, 0, self.n_in, self.n_witness, len(self.constraints))
        import numpy as np
        wmap = np.arange(self.n_witness, dtype="<u8").tobytes()
        out = b"r1cs" + struct.pack("<II", 1, 3)
        for sid, pl in ((1, hdr), (2, bytes(body)), (3, wmap)):
            out += struct.pack("<IQ", sid, len(pl)) + pl
        return out

    # ---- witness program (.wprog) -----------------------------------------
    def wprog_bytes(self):
        """Layout (all u32 little-endian unless noted):
        "NZWP", version=1, n_total, n_witness, n_out, n_in, n_consts, n_instr, n_levels, n_code
        consts   n_consts x 32 B canonical LE   (index 0 = 1, index 1 = r-1)
        ioff     n_instr offsets into code (instructions sorted by level)
        lstart   n_levels + 1 indices into ioff
        code     instruction words:
           LIN    op, dst, <lc>
           MUL    op, dst, <lc a>, <lc b>, <lc c>
           BITS   op, dst0, src, n
           INV    op, dst, src
           ASSERT op, <lc a>, <lc b>, <lc c>
           BITSLC op, dst0, n, <lc>          (bits of the value of an LC)
           SHAROUND op, n, r1a r1b r1c (Sigma1), r0a r0b r0c (Sigma0), K_lo, K_hi, w0, size, 9 n x <bit>, <lc>...
                     (bits of a b c d e f g h w, LSB first)        -- native program only
           SHASCHED op, n, r1a r1b r1c (sigma1, r1c a shift), r0a r0b r0c (sigma0), w0, size, 4 n x <bit>, <lc>...
                     (bits of w[t-2], w[t-7], w[t-15], w[t-16])    -- native program only
           QUINSEL  op, N, N x (eq wire, sum wire | 0xffffffff), N x <bit> (the inputs), <lc index>, <lc>...
                     -- native program only; the IsZero inverse of choice i is the wire before its eq wire
           SHABLOCK op, n, 12 rotation counts (Sigma1, Sigma0, sigma1, sigma0), r_start, rounds, (rounds - 16) x w0 of
                     the schedule steps, (rounds - r_start) x w0 of the rounds, (rounds - r_start) x K (low word),
                     24 n x <bit> (state a..h at round r_start, message words 0..15), <lc>...  -- native program only
           SHAROUNDS op, n, 6 rotation counts (Sigma1, Sigma0), r_start, rounds, (rounds - r_start) x w0 of the rounds,
                     (rounds - r_start) x (K_lo, K_hi), (8 + rounds - r_start) n x <bit> (state a..h at round r_start,
                     then w[t] for every round), <lc>...  -- native program only
           <bit> = wire id, or 0x80000000 | offset (from the instruction's first word) of the <lc> giving the value
           <lc> = n_terms, const_idx (0xffffffff = no constant), then n_terms x (wire, coef_idx)"""
        consts = {1: 0, R - 1: 1}
        clist = [1, R - 1]

        def cidx(v):
            i = consts.get(v)
            if i is None:
                i = consts[v] = len(clist)
                clist.append(v)
            return i

        code = []
        ioff = []

        def emit_lc(lc):
            code.append(len(lc.t))
            code.append(cidx(lc.k) if lc.k else 0xFFFFFFFF)
            for w in sorted(lc.t):
                code.append(w)
                code.append(cidx(lc.t[w]))

        for ins in self.prog:
            ioff.append(len(code))
            op = ins[0]
            code.append(op)
            if op == OP_SHAROUNDS:
                pl = ins[1]
                start = ioff[-1]
                code.append(pl["n"])
                code.extend(pl["rot1"])
                code.extend(pl["rot0"])
                code.extend((pl["r_start"], pl["rounds"]))
                code.extend(pl["round_w0"])
                for k in pl["K"]:
                    code.extend((k & 0xFFFFFFFF, k >> 32))
                refs = len(code)
                bits = [b for word in pl["words"] for b in word]
                code.extend([0] * len(bits))
                for k, b in enumerate(bits):
                    w = b.single_wire()
                    if w is not None:
                        code[refs + k] = w
                    else:
                        code[refs + k] = 0x80000000 | (len(code) - start)
                        emit_lc(b)
            elif op == OP_SHABLOCK:
                pl = ins[1]
                start = ioff[-1]
                code.append(pl["n"])
                for key in ("rot1", "rot0", "srot1", "srot0"):
                    code.extend(pl[key])
                code.extend((pl["r_start"], pl["rounds"]))
                code.extend(pl["sched_w0"])
                code.extend(pl["round_w0"])
                code.extend(k & 0xFFFFFFFF for k in pl["K"])
                refs = len(code)
                bits = [b for word in pl["words"] for b in word]
                code.extend([0] * len(bits))
                for k, b in enumerate(bits):
                    w = b.single_wire()
                    if w is not None:
                        code[refs + k] = w
                    else:
                        code[refs + k] = 0x80000000 | (len(code) - start)
                        emit_lc(b)
            elif op == OP_QUINSEL:
                pl = ins[1]
                start = ioff[-1]
                code.append(len(pl["ins"]))
                for ew, sw in zip(pl["eq_w"], pl["sum_w"]):
                    code.extend((ew, 0xFFFFFFFF if sw is None else sw))
                refs = len(code)
                code.extend([0] * len(pl["ins"]))
                emit_lc(pl["index"])
                for k, b in enumerate(pl["ins"]):
                    w = b.single_wire()
                    if w is not None:
                        code[refs + k] = w
                    else:
                        code[refs + k] = 0x80000000 | (len(code) - start)
                        emit_lc(b)
            elif op in (OP_SHAROUND, OP_SHASCHED):
                pl = ins[1]
                code.append(pl["n"])
                code.extend(pl["rot1"])
                code.extend(pl["rot0"])
                if op == OP_SHAROUND:
                    code.extend((pl["K"] & 0xFFFFFFFF, pl["K"] >> 32))
                code.extend((pl["w0"], pl["size"]))
                start = ioff[-1]
                refs = len(code)
                bits = [b for word in pl["words"] for b in word]
                code.extend([0] * len(bits))
                for k, b in enumerate(bits):
                    w = b.single_wire()
                    if w is not None:
                        code[refs + k] = w
                    else:
                        code[refs + k] = 0x80000000 | (len(code) - start)
                        emit_lc(b)
            elif op == OP_LIN:
                code.append(ins[1])
                emit_lc(ins[2])
            elif op == OP_MUL:
                code.append(ins[1])
                emit_lc(ins[2]); emit_lc(ins[3]); emit_lc(ins[4])
            elif op == OP_BITS:
                code.extend((ins[1], ins[2], ins[3]))
            elif op == OP_INV:
                code.extend((ins[1], ins[2]))
            elif op == OP_BITSLC:
                code.extend((ins[1], ins[2]))
                emit_lc(ins[3])
            else:
                emit_lc(ins[1]); emit_lc(ins[2]); emit_lc(ins[3])
        lstart = [0] * (self.n_levels + 1)
        for l in self.levels:
            lstart[l] += 1  # count at index l (levels are 1-based) -> prefix below
        acc = 0
        starts = []
        for l in range(1, self.n_levels + 1):
            starts.append(acc)
            acc += lstart[l]
        starts.append(acc)
        import numpy as np
        hdr = b"NZWP" + struct.pack("<IIIIIIIII", 1, self.n_total, self.n_witness, self.n_out, self.n_in, len(clist),
                                    len(self.prog), self.n_levels, len(code))
        return (hdr + b"".join(v.to_bytes(32, "little") for v in clist) + np.asarray(ioff, dtype="<u4").tobytes() +
                np.asarray(starts, dtype="<u4").tobytes() + np.asarray(code, dtype="<u4").tobytes())

    def artifact(self):
        """the build products a circom run leaves on disk: .r1cs, witness program (generic, and the native one the
        GPU loads when the circuit has word-level steps), I/O table"""
        art = Artifact(self.name, self.n_witness, self.n_total, self.n_out, self.n_in, list(self.inputs),
                       list(self.outputs), len(self.constraints), self.n_levels, len(self.prog), self.wprog_bytes(),
                       self.r1cs_bytes())
        if not self.native and os.environ.get("NZCB_WITNESS_NATIVE", "1") != "0" and \
                any(i[0] in FAT_OPS for i in self._c.prog):
            nat = Compiled(self._c, native=True)
            art.wprog_native = nat.wprog_bytes()
            art.n_levels_native, art.n_instr_native = nat.n_levels, len(nat.prog)
        return art

    def flatten_input(self, inp: dict):
        return flatten_input(self.inputs, inp)


class Artifact:
    def __init__(self, name, n_witness, n_total, n_out, n_in, inputs, outputs, n_constraints, n_levels, n_instr, wprog,
                 r1cs):
        self.name, self.n_witness, self.n_total, self.n_out, self.n_in = name, n_witness, n_total, n_out, n_in
        self.inputs, self.outputs = inputs, outputs
        self.n_constraints, self.n_levels, self.n_instr = n_constraints, n_levels, n_instr
        self.wprog, self.r1cs = wprog, r1cs

    def wprog_bytes(self, native=False):
        """the witness program: generic (what the oracle VMs run; always present) or, native=True, the one with
        word-level SHA-2 steps that the GPU loads (falls back to the generic one when the circuit has none)"""
        if native:
            return getattr(self, "wprog_native", None) or self.wprog
        return self.wprog

    def r1cs_bytes(self):
        return self.r1cs

    def flatten_input(self, inp: dict):
        return flatten_input(self.inputs, inp)

    def sym_bytes(self):
        """the circuit's input table for nzcb_inputs_resolve (include/nzcb.h): circom_runtime addresses main's input
        signals by the FNV-1a-64 hash of their name; offsets are flattened declaration order"""
        return sym_bytes(self.inputs)


def fnv1a64(name: str) -> int:
    h = 0xCBF29CE484222325
    for ch in name.encode():
        h = ((h ^ ch) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def sym_bytes(inputs):
    import struct
    out, off = b"", 0
    for name, dims, _first in inputs:
        n = Circuit._size(dims)
        out += struct.pack("<QII", fnv1a64(name), off, n)
        off += n
    return b"NZSY" + struct.pack("<III", 1, len(inputs), off) + out


# ---- input marshalling (circom_runtime: names, arrays flattened row-major) --
def flatten_input(inputs, inp: dict):
    vals = []
    for name, dims, _first in inputs:
        if name not in inp:
            raise KeyError(f"Signal not found: {name}")  # circom_runtime error 1
        flat = []

        def walk(x):
            if isinstance(x, (list, tuple)):
                for y in x:
                    walk(y)
            else:
                flat.append(int(x) % R)

        walk(inp[name])
        n = Circuit._size(dims)
        if len(flat) > n:
            raise ValueError(f"Too many values for input signal {name}")  # error 2 / 6
        if len(flat) < n:
            raise ValueError(f"Not enough values for input signal {name}")
        vals.extend(flat)
    extra = set(inp) - {n for n, _, _ in inputs}
    if extra:
        raise KeyError(f"Signal not found: {sorted(extra)[0]}")
    return vals
