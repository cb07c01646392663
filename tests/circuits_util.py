"""Random satisfiable R1CS circuits + witnesses for prover parity tests.
Wire layout follows circom: wire 0 = 1, then outputs, then inputs, then internals."""
import random

from oracle.binfile import R1CS
from oracle.bn254 import R_MOD as R


def random_circuit(seed, n_out=2, n_in=3, n_mul=40, max_terms=6, public_inputs=0):
    """Returns (R1CS, witness list).  Mix of: general A*B=C with long LCs (forces PLONK
    "additions"), constant-A / zero-A linear constraints, boolean constraints."""
    rng = random.Random(seed)
    n_pub_in = public_inputs
    first_int = 1 + n_out + n_in
    w = [1] + [0] * n_out + [rng.randrange(R) for _ in range(n_in)]
    cons = []
    avail = list(range(1 + n_out, 1 + n_out + n_in))  # wires whose value is known

    def ev(lc):
        return sum(c * w[s] for s, c in lc.items()) % R

    def rand_lc(nt):
        lc = {}
        for _ in range(nt):
            s = rng.choice(avail + [0])
            c = rng.choice([1, R - 1, 2, rng.randrange(R), rng.randrange(1 << 16)])
            lc[s] = (lc.get(s, 0) + c) % R
        return {s: c for s, c in lc.items() if c}

    for k in range(n_mul):
        kind = rng.randrange(6)
        new = len(w)
        if kind <= 2:  # new = LCa * LCb + LCc
            la, lb, lc = rand_lc(rng.randrange(1, max_terms)), rand_lc(rng.randrange(1, max_terms)), rand_lc(
                rng.randrange(0, max_terms))
            w.append((ev(la) * ev(lb) + ev(lc)) % R)
            c_side = {new: 1}
            for s, c in lc.items():
                c_side[s] = (c_side.get(s, 0) - c) % R
            cons.append((la, lb, c_side))
        elif kind == 3:  # constant A:  k * LCb = new
            kk = rng.randrange(1, R)
            lb = rand_lc(rng.randrange(1, max_terms + 3))
            w.append(kk * ev(lb) % R)
            cons.append(({0: kk}, lb, {new: 1}))
        elif kind == 4:  # zero A: 0 = LC  (new = linear combination)
            lb = rand_lc(rng.randrange(1, max_terms + 3))
            w.append(ev(lb))
            lc = dict(lb)
            lc[new] = (lc.get(new, 0) - 1) % R
            cons.append(({}, rand_lc(2), lc))
        else:  # boolean-ish: bit * (bit - 1) = 0
            w.append(rng.randrange(2))
            cons.append(({new: 1}, {new: 1, 0: R - 1}, {}))
        avail.append(new)
    # outputs: out_i = LCa * LCb
    for i in range(n_out):
        la, lb = rand_lc(3), rand_lc(3)
        w[1 + i] = ev(la) * ev(lb) % R
        cons.append((la, lb, {1 + i: 1}))
    r = R1CS(len(w), n_out, n_pub_in, n_in - n_pub_in, cons)
    for la, lb, lc in cons:
        assert ev(la) * ev(lb) % R == ev(lc)
    assert first_int <= len(w)
    return r, w
