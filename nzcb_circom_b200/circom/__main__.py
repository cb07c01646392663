"""`python -m nzcb_circom_b200.circom <main> [<main> ...] -o <dir>` -- the circom stand-in's command line: what
`circom circuits/x.circom --r1cs --wasm --sym` (/root/reference/Makefile:45-52 and circom_tester's own compile step)
leaves on disk, for the mains of /root/reference/circuits restated in circom_tester.MAINS:

    <dir>/<main>.r1cs     iden3 r1cs v1
    <dir>/<main>.wprog    the witness program nzcb_circuit_load takes (the role of <main>_js/<main>.wasm)
    <dir>/<main>.sym      input table for nzcb_inputs_resolve (name hash -> offset, size)
    <dir>/<main>.json     names, dimensions and counts, for humans and for integration/nzcb.js

`--all` writes every main; a path like circuits/nzcp_live.circom is accepted and reduced to its base name, so the
reference's own file names can be passed (test/nzcp.js:104,118,347)."""
import argparse
import json
import os
import sys


def main(argv=None):
    from ..circom_tester import MAINS, compile_circuit

    ap = argparse.ArgumentParser(prog="python -m nzcb_circom_b200.circom")
    ap.add_argument("mains", nargs="*")
    ap.add_argument("--all", action="store_true")
    ap.add_argument("-o", "--out", required=True)
    a = ap.parse_args(argv)
    names = sorted(MAINS) if a.all else a.mains
    if not names:
        ap.error("name at least one main component, or --all")
    os.makedirs(a.out, exist_ok=True)
    for name in names:
        key = os.path.splitext(os.path.basename(name))[0]
        art = compile_circuit(key)
        for ext, data in ((".r1cs", art.r1cs_bytes()), (".wprog", art.wprog_bytes(native=True)), (".sym", art.sym_bytes())):
            with open(os.path.join(a.out, key + ext), "wb") as f:
                f.write(data)
        meta = {"main": key, "nWitness": art.n_witness, "nOutputs": art.n_out, "nInputs": art.n_in,
                "nConstraints": art.n_constraints,
                "inputs": [{"name": n, "dims": list(d), "firstWire": f} for n, d, f in art.inputs],
                "outputs": [{"name": o[0], "dims": list(o[1]), "firstWire": o[2]} if isinstance(o, (tuple, list)) else o for o in art.outputs]}
        with open(os.path.join(a.out, key + ".json"), "w") as f:
            json.dump(meta, f, indent=1)
        print(f"{key}: {art.n_constraints} constraints, {art.n_witness} wires, {art.n_in} inputs -> {a.out}/{key}.{{r1cs,wprog,sym,json}}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
