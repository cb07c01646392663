"""CPU oracle for the nzcb PLONK / witness hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``nzcb_circom_b200/`` (the product)
may import this package; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do, and there only
as the checker.

PARITY UNPINNED at the snarkjs / ffjavascript / circom-WASM level: those
packages (snarkjs 0.4.12, ffjavascript 0.2.48, wasmcurves 0.1.0,
circom_runtime 0.1.17, js-sha3 0.8.0 -- /root/reference/yarn.lock:7279,3905,
8173,2496,5074) are un-vendored npm dependencies and no Node runtime exists in
this image, so the oracle restates their published algorithms (SURVEY.md
Appendix A) in Python big-int arithmetic.  What IS pinned: the witness public
outputs of the reference's own tests (test/nzcp.js:62-68, test/utils.js:17,
test/cbor.js, test/quinSelector.js), Keccak/SHA known-answer tests, and the
algebraic self-consistency of prover vs. an independently written verifier
plus a known-trapdoor KZG check.

Modules: bn254 / ntt / keccak (field, curve, transform, transcript hash), binfile (iden3 file formats), plonk (setup,
prove, verify), witness_vm (the witness program interpreter), c/ (the same path in C with OpenMP: the CPU baseline),
pairing (BN254 optimal ate pairing: pinned by bilinearity, non-degeneracy and order r, and by agreeing with the
known-trapdoor form of plonk.verify), pass_ingest (test/helpers/nzcp.js with JavaScript's number semantics: pinned
by the reference's EXAMPLE_PASS_URI -> ToBeSigned -> SHA-256 of test/utils.js:17), ptau (writer of a small
known-trapdoor .ptau for the reader's tests: format recalled, unpinned).
"""
