// nzcb.js -- the JavaScript face of libnzcb.so: the same shapes as the three npm packages the reference drives this
// path through, so test/nzcp.js, test/cbor.js and the Makefile recipes run unchanged on top of it:
//
//   const { wasm: wasm_tester } = require("circom_tester")   ->  require("./integration/nzcb").wasm
//   snarkjs.plonk.{prove, fullProve, verify}                  ->  require("./integration/nzcb").plonk
//   getCOSE + encodeToBeSigned (test/helpers/nzcp.js)         ->  require("./integration/nzcb").toBeSigned
//
// NOT RUN IN THIS REPOSITORY'S ENVIRONMENT (no Node in the image, SURVEY.md 0.2): the Python mirrors
// nzcb_circom_b200/{circom_tester,snarkjs,pass_ingest}.py are line-for-line the same logic over the same C ABI and
// are what the tests execute.
"use strict";
const fs = require("fs");
const n = require("./build/Release/nzcb_napi.node");

const R = 21888242871839275222246405745257275088548364400416034343698204186575808495617n;
const le32 = (x) => { const b = Buffer.alloc(32); let v = ((BigInt(x) % R) + R) % R; for (let i = 0; i < 32; i++) { b[i] = Number(v & 255n); v >>= 8n; } return b; };
const fromLe = (b, o) => { let v = 0n; for (let i = 31; i >= 0; i--) v = (v << 8n) | BigInt(b[o + i]); return v; };
const be32 = (x) => Buffer.from(BigInt(x).toString(16).padStart(64, "0"), "hex");
const flatten = (v, out) => { if (Array.isArray(v)) v.forEach((e) => flatten(e, out)); else out.push(le32(v)); return out; };

const keys = new Map(); // zkey made device resident once per process
const zkeyOf = (f) => { const k = typeof f === "string" ? f : f.data; if (!keys.has(k)) keys.set(k, n.loadZkey(typeof f === "string" ? fs.readFileSync(f) : Buffer.from(f.data))); return keys.get(k); };
const POINTS = ["A", "B", "C", "Z", "T1", "T2", "T3", "Wxi", "Wxiw"];
const EVALS = ["eval_a", "eval_b", "eval_c", "eval_s1", "eval_s2", "eval_zw", "eval_r"];
const proofToBytes = (p) => Buffer.concat([
  ...POINTS.map((k) => (String(p[k][2]) === "0" ? Buffer.alloc(64) : Buffer.concat([be32(p[k][0]), be32(p[k][1])]))),
  ...EVALS.map((k) => be32(p[k]))]);
const publicsOf = (buf) => { const out = []; for (let o = 0; o < buf.length; o += 32) out.push(fromLe(buf, o).toString()); return out; };

// circom_tester.wasm(circomFile): the circuit is resolved BY FILE NAME, as test/nzcp.js:104,118,347 pass it
// (`${__dirname}/../circuits/nzcp_live.circom`): <buildDir>/<basename>.{wprog,sym,json} written by
// `python -m nzcb_circom_b200.circom --all -o <buildDir>` (the role of circom's own --wasm --sym outputs).
// buildDir: options.buildDir, $NZCB_CIRCUITS or ./circuits_build next to this file.
const path = require("path");
const fnv1a64 = (str) => { let h = 0xCBF29CE484222325n; for (const c of Buffer.from(str, "utf8")) { h ^= BigInt(c); h = (h * 0x100000001B3n) & 0xFFFFFFFFFFFFFFFFn; } return h; };
exports.wasm = async (circomFile, options) => {
  const dir = (options && options.buildDir) || process.env.NZCB_CIRCUITS || path.join(__dirname, "circuits_build");
  const base = path.join(dir, path.basename(circomFile, ".circom"));
  const meta = JSON.parse(fs.readFileSync(base + ".json", "utf8"));
  const sym = fs.readFileSync(base + ".sym");
  const c = n.loadCircuit(fs.readFileSync(base + ".wprog"));
  // circom_runtime's _doCalculateWitness: every key of `input`, hashed, values flattened row-major; the errors
  // ("Signal not found", "Too many signals set", "Signal already set", "Input signal array access exceeds the size",
  // "Not all inputs have been set...") come from the library (nzcb_inputs_resolve)
  const marshal = (input) => {
    const names = Object.keys(input);
    const hashes = Buffer.alloc(8 * names.length), counts = Buffer.alloc(4 * names.length), vals = [];
    names.forEach((k, i) => { const f = flatten(input[k], []); hashes.writeBigUInt64LE(fnv1a64(k), 8 * i); counts.writeUInt32LE(f.length, 4 * i); vals.push(...f); });
    return n.resolveInputs(sym, hashes, counts, Buffer.concat(vals), meta.nInputs);
  };
  return {
    circuit: c, marshal, meta,
    async calculateWitness(input, sanityCheck) {
      const { witness, status } = n.calculateWitness(c, marshal(input), 1);
      if (sanityCheck !== false && status.readInt32LE(0) !== 0) throw new Error("Error: Assert Failed.");
      const w = []; for (let o = 0; o < witness.length; o += 32) w.push(fromLe(witness, o)); return w;
    },
    async calculateWTNSBin(input, sanityCheck) { // circom_runtime's name: the .wtns file bytes
      const { witness, status } = n.calculateWitness(c, marshal(input), 1);
      if (sanityCheck !== false && status.readInt32LE(0) !== 0) throw new Error("Error: Assert Failed.");
      return n.wtnsExport(witness);
    },
  };
};
exports.zKey = { async exportVerificationKey(zkeyFile) { return JSON.parse(n.vkeyToJson(typeof zkeyFile === "string" ? fs.readFileSync(zkeyFile) : Buffer.from(zkeyFile.data))); } };

exports.plonk = {
  async prove(zkeyFile, wtnsFile) { // snarkjs.plonk.prove(zkeyFileName, witnessFileName)
    const wt = wtnsFile.type === "mem" ? Buffer.from(wtnsFile.data) : fs.readFileSync(wtnsFile);
    const { proof, publicSignals } = n.prove(zkeyOf(zkeyFile), wt, null);
    return { proof: JSON.parse(n.proofToJson(proof)), publicSignals: publicsOf(publicSignals) };
  },
  async fullProve(input, tester, zkeyFile) { // snarkjs.plonk.fullProve(input, wasmFile, zkeyFileName)
    const r = n.fullProveBatch(tester.circuit, zkeyOf(zkeyFile), tester.marshal(input), 1, null);
    if (r.status.readInt32LE(0) !== 0) throw new Error(r.status.readInt32LE(0) === -6 ? "Error: Assert Failed." : "nzcb: prover error");
    return { proof: JSON.parse(n.proofToJson(r.proofs.subarray(0, 800))), publicSignals: publicsOf(r.publicSignals) };
  },
  async verify(vk, publicSignals, proof) { // snarkjs.plonk.verify(vk_verifier, publicSignals, proof)
    return n.verify(n.loadVkey(JSON.stringify(vk)), Buffer.concat(publicSignals.map(le32)), proofToBytes(proof));
  },
};

// pass URIs -> ToBeSigned (test/helpers/nzcp.js getCOSE + encodeToBeSigned) and -> proofs, on the device
const pack = (uris) => { const bufs = uris.map((u) => Buffer.from(u, "latin1")); const off = Buffer.alloc(4 * (uris.length + 1)); let o = 0; bufs.forEach((b, i) => { off.writeUInt32LE(o, 4 * i); o += b.length; }); off.writeUInt32LE(o, 4 * uris.length); return [Buffer.concat(bufs), off]; };
exports.toBeSigned = (passURIs, maxLen) => { const [u, off] = pack(passURIs); return n.toBeSigned(u, off, maxLen); };
exports.fullProveURIs = (tester, zkeyFile, passURIs, data, maxLen) => { const [u, off] = pack(passURIs); return n.fullProveURIs(tester.circuit, zkeyOf(zkeyFile), u, off, data || null, maxLen); };
