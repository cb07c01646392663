// snarkjs pin kit -- the ONE route by which "parity unpinned" (DESIGN.md row (c)) turns green, and the real CPU
// baseline B1 of BASELINE.md section 3.  Needs Node >= 14 with the reference's own dependencies installed
// (`yarn install` in noway/nzcb-circom: snarkjs ^0.4.12, package.json:18; ffjavascript 0.2.48, yarn.lock:3905).
// None of that exists in the build image, so this file is shipped unexecuted; it prints one JSON line and exits 0
// in every case (missing modules -> {"unavailable": ...}).
//
//   node baseline/snarkjs_baseline.mjs <fixture_dir> [--reps N] [--node-modules <dir>]
//
// <fixture_dir> is written by `python tools/export_fixture.py` from THIS repository's outputs:
//   circuit.zkey  witness.wtns  blinders.json  proof.json  public.json  verification_key.json
// What it does, for the nine blinding scalars of blinders.json injected through curve.Fr.random
// (snarkjs 0.4.12 src/plonk_prove.js draws b_1..b_9 with nine Fr.random() calls, in that order):
//   1. snarkjs.plonk.prove(circuit.zkey, witness.wtns)      -> proof, publicSignals
//   2. byte comparison of JSON.stringify(proof, null, 1) / publicSignals with proof.json / public.json
//   3. snarkjs.plonk.verify(verification_key.json, ...) of snarkjs' proof AND of this repository's proof
//   4. zKey.exportVerificationKey(circuit.zkey) deep-compared with verification_key.json
//   5. timing: N further plonk.prove calls (ffjavascript's worker pool = all host cores) -> proofs/s
import { createRequire } from "module";
import fs from "fs";
import os from "os";
import path from "path";

const args = process.argv.slice(2);
const dir = args.find((a) => !a.startsWith("--")) || ".";
const opt = (name, dflt) => {
    const i = args.indexOf(name);
    return i >= 0 && i + 1 < args.length ? args[i + 1] : dflt;
};
const reps = parseInt(opt("--reps", "3"), 10);
const nm = opt("--node-modules", null);
const out = (o) => {
    console.log(JSON.stringify(o));
    process.exit(0);
};

let snarkjs, ffjs;
try {
    const require = createRequire(nm ? path.resolve(nm, "_") : import.meta.url);
    snarkjs = require("snarkjs");
    ffjs = require("ffjavascript");
} catch (e) {
    out({ impl: "snarkjs", unavailable: "snarkjs / ffjavascript not resolvable: " + String(e.message).split("\n")[0] });
}

const need = ["circuit.zkey", "witness.wtns", "blinders.json", "proof.json", "public.json", "verification_key.json"];
for (const f of need) if (!fs.existsSync(path.join(dir, f))) out({ impl: "snarkjs", unavailable: `fixture file missing: ${f}` });
const rd = (f) => JSON.parse(fs.readFileSync(path.join(dir, f), "utf8"));
const blinders = rd("blinders.json").map((x) => BigInt(x));
const ourProofText = fs.readFileSync(path.join(dir, "proof.json"), "utf8");
const ourProof = JSON.parse(ourProofText);
const ourPublic = rd("public.json");
const vk = rd("verification_key.json");

// ffjavascript caches the curve object (globalThis.curve_bn128): the prover gets this very instance
const curve = await ffjs.getCurveFromName("bn128");
const realRandom = curve.Fr.random.bind(curve.Fr);
let draw = 0;
const inject = () => {
    draw = 0;
    curve.Fr.random = () => (draw < blinders.length ? curve.Fr.e(blinders[draw++]) : realRandom());
};

const zkeyPath = path.join(dir, "circuit.zkey");
const wtnsPath = path.join(dir, "witness.wtns");
const res = { impl: "snarkjs", node: process.version, cores: os.cpus().length, fixture: path.resolve(dir) };
try {
    res.snarkjs_version = createRequire(nm ? path.resolve(nm, "_") : import.meta.url)("snarkjs/package.json").version;
} catch (e) { /* exports map may hide package.json */ }

inject();
const { proof, publicSignals } = await snarkjs.plonk.prove(zkeyPath, wtnsPath);
res.blinders_drawn = draw;                                              // must be 9
res.proof_json_bytes_equal = JSON.stringify(proof, null, 1) === ourProofText.trimEnd();
res.proof_fields_equal = JSON.stringify(proof) === JSON.stringify(ourProof);
res.public_signals_equal = JSON.stringify(publicSignals) === JSON.stringify(ourPublic);
if (!res.proof_fields_equal) {
    res.first_difference = Object.keys(proof).find((k) => JSON.stringify(proof[k]) !== JSON.stringify(ourProof[k]));
}
res.snarkjs_verifies_snarkjs_proof = await snarkjs.plonk.verify(vk, publicSignals, proof);
res.snarkjs_verifies_our_proof = await snarkjs.plonk.verify(vk, ourPublic, ourProof);
try {
    const vk2 = await snarkjs.zKey.exportVerificationKey(zkeyPath);
    res.verification_key_equal = JSON.stringify(vk2) === JSON.stringify(vk);
} catch (e) {
    res.verification_key_equal = "exportVerificationKey failed: " + e.message;
}
if (fs.existsSync(path.join(dir, "calldata.txt"))) {
    const cd = await snarkjs.plonk.exportSolidityCallData(ourProof, ourPublic);
    res.calldata_equal = cd.trim() === fs.readFileSync(path.join(dir, "calldata.txt"), "utf8").trim();
}

const t0 = process.hrtime.bigint();
for (let i = 0; i < reps; i++) {
    inject();
    await snarkjs.plonk.prove(zkeyPath, wtnsPath);
}
const secs = Number(process.hrtime.bigint() - t0) / 1e9;
res.metric = "PLONK proofs/s (snarkjs.plonk.prove, witness given)";
res.steps = reps;
res.value = reps / secs;
res.unit = "proofs/s";
res.ms_per_proof = (1000 * secs) / reps;
await curve.terminate();
out(res);
