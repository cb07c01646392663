"""Run a circuit's witness calculation on either backend, the way the reference's Mocha
tests call ``cir.calculateWitness(input, true)``:

  "oracle" -- CPU restatement (oracle/witness_vm.py) of the witness program, CPU-only suite
  "gpu"    -- the product: nzcb_witness_batch through the C ABI (tests marked gpu)

A rejected input (``assert.isRejected``) comes back as None."""
import pytest

from nzcb_circom_b200.circom_tester import compile_circuit, wasm_tester
from oracle import witness_vm as vm

BACKENDS = ["oracle", pytest.param("gpu", marks=pytest.mark.gpu)]
_programs = {}
_testers = {}


def calc_many(name, inputs, backend, check_r1cs=False):
    art = compile_circuit(name)
    flat = [art.flatten_input(i) for i in inputs]
    if backend == "oracle":
        if name not in _programs:
            _programs[name] = vm.Program(art.wprog_bytes())
        out = []
        for f in flat:
            try:
                out.append(vm.run(_programs[name], f))
            except vm.AssertFailed:
                out.append(None)
    else:
        if name not in _testers:
            _testers[name] = wasm_tester(name)
        raw, st = _testers[name].calculateWitnessBatch(flat, True)
        nw = art.n_witness
        out = []
        for i, s in enumerate(st):
            if s != 0:
                out.append(None)
            else:
                blk = raw[i * nw * 32:(i + 1) * nw * 32]
                out.append([int.from_bytes(blk[k:k + 32], "little") for k in range(0, len(blk), 32)])
    if check_r1cs:
        for w in out:
            if w is not None:
                assert vm.check_r1cs(art.r1cs_bytes(), w) == -1, "witness violates the circuit's own R1CS"
    return out


def calc(name, inp, backend, check_r1cs=False):
    return calc_many(name, [inp], backend, check_r1cs)[0]
