"""Static profile of a circuit's witness program (CPU only): instructions per opcode, LC terms, level widths.
The numbers behind DESIGN.md's discussion of the witness kernel (what a pass has to execute, how wide its levels
are, how much of it is the SHA round chain).   python tools/witness_stats.py [circuit]"""
import os
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nzcb_circom_b200.circom_tester import compile_circuit  # noqa: E402
from oracle import witness_vm as vm  # noqa: E402

NAMES = {1: "LIN", 2: "MUL", 3: "BITS", 4: "INV", 5: "ASSERT", 6: "BITSLC"}


def main(name):
    art = compile_circuit(name)
    p = vm.Program(art.wprog_bytes())
    code = p.code
    ops, terms, bits_out = Counter(), Counter(), 0

    def lc_terms(q):
        return code[q], q + 2 + 2 * code[q]

    for off in p.ioff:
        op = code[off]
        ops[op] += 1
        if op == 1:
            terms[op] += lc_terms(off + 2)[0]
        elif op in (2, 5):
            q = off + (2 if op == 2 else 1)
            for _ in range(3):
                n, q = lc_terms(q)
                terms[op] += n
        elif op == 3:
            bits_out += code[off + 3]
        elif op == 6:
            bits_out += code[off + 2]
            terms[op] += lc_terms(off + 3)[0]
    widths = [p.lstart[i + 1] - p.lstart[i] for i in range(p.n_levels)]
    print(f"{name}: {p.n_instr} instructions, {p.n_levels} levels, {p.n_total} wires incl. temps ({p.n_witness} witness), "
          f"{p.n_in} inputs, {len(p.consts)} constants, {len(code) * 4 / 1e6:.1f} MB of code")
    for op, n in sorted(ops.items()):
        print(f"  {NAMES[op]:7s} {n:8d} instructions  {terms[op]:9d} LC terms  ({terms[op] / max(n, 1):.1f} per instruction)")
    print(f"  bits written by BITS / BITSLC: {bits_out}")
    ws = sorted(widths)
    print(f"  level width: min {ws[0]}, median {ws[len(ws) // 2]}, mean {sum(ws) / len(ws):.0f}, p90 {ws[int(0.9 * len(ws))]}, max {ws[-1]}")
    for lim in (32, 128, 384, 1024):
        n = sum(1 for w in widths if w <= lim)
        print(f"  levels of width <= {lim:4d}: {n:5d} ({100 * n / len(widths):.0f} %), holding {100 * sum(w for w in widths if w <= lim) / p.n_instr:.0f} % of the instructions")
    # work at T threads: sum over levels of ceil(width / T) "instruction slots" on the critical path
    for T in (128, 384):
        slots = sum((w + T - 1) // T for w in widths)
        print(f"  critical path at {T} threads: {slots} instruction slots ({slots / p.n_levels:.2f} per level)")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "nzcp_live")
