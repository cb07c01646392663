/* nzcb.h -- C ABI of libnzcb.so, the B200-native replacement for the proving /
 * witness path that noway/nzcb-circom reaches through third-party JS:
 *
 *   snarkjs 0.4.12 plonk.prove / fullProve / setup      (/root/reference/yarn.lock:7279,
 *                                                        recipe at /root/reference/Makefile:54-62)
 *   ffjavascript 0.2.48 G1.multiExpAffine, Fr.fft/ifft  (/root/reference/yarn.lock:3905)
 *   circom_tester.wasm(...).calculateWitness(input, sanityCheck)
 *                                                       (/root/reference/test/nzcp.js:3,42;
 *                                                        test/cbor.js, test/quinSelector.js call sites)
 *
 * Conventions
 *   - every function returns an int32 status: 0 = OK, < 0 = NZCB_E_*; nothing
 *     throws or aborts across the ABI; nzcb_last_error(ctx) returns a UTF-8
 *     message for the last failure on that ctx (snarkjs' own strings where it
 *     has one: "Invalid witness length", "Copy constraints does not match",
 *     "T Polynomial is not divisible", "Polinomial does not divide").
 *   - plain pointers and sizes only.  All *input* buffers are HOST memory,
 *     borrowed for the duration of the call; all *output* buffers are caller-
 *     allocated HOST memory.  Device-resident variants end in _dev.
 *   - byte layouts are exactly the iden3 file layouts so JS
 *     {type:"mem"} buffers pass straight through:
 *       "LEM" = 32-byte little-endian Montgomery (R = 2^256) field element,
 *       "LE"  = 32-byte little-endian canonical,  "BE" = big-endian canonical;
 *       G1 affine = x||y (64 B), infinity = all zero.
 *   - one ctx = one GPU = one CUDA stream set; a ctx is thread-compatible
 *     (one host thread at a time).  One process per GPU is the intended use.
 *   - there is no CPU fallback: every entry point fails with NZCB_E_CUDA if no
 *     sm_100-class device is usable.
 */
#ifndef NZCB_H
#define NZCB_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NZCB_OK 0
#define NZCB_E_INVALID (-1)   /* bad argument / malformed file */
#define NZCB_E_CUDA (-2)      /* CUDA runtime failure (message has the detail) */
#define NZCB_E_WITNESS (-3)   /* "Invalid witness length" and friends */
#define NZCB_E_COPY (-4)      /* "Copy constraints does not match" */
#define NZCB_E_DIVIDE (-5)    /* "T Polynomial is not divisible" / "Polinomial does not divide" */
#define NZCB_E_ASSERT (-6)    /* witness program: "Assert Failed" (circom_runtime error 4) */
#define NZCB_E_NOMEM (-7)
/* circom_runtime's input errors (witness_calculator.js exception codes): 1 "Signal not found",
 * 2 "Too many signals set", 3 "Signal already set", 6 "Input signal array access exceeds the size"; 7 = the JS
 * wrapper's "Not all inputs have been set" */
#define NZCB_E_SIGNAL(code) (-(100 + (code)))

typedef struct nzcb_ctx nzcb_ctx;
typedef struct nzcb_zkey nzcb_zkey;
typedef struct nzcb_circuit nzcb_circuit;

/* proof = what snarkjs puts in proof.json, binary:
 * nine G1 points as x||y BE canonical (toRprUncompressed) and seven Fr BE. */
typedef struct nzcb_proof {
    uint8_t A[64], B[64], C[64], Z[64], T1[64], T2[64], T3[64], Wxi[64], Wxiw[64];
    uint8_t eval_a[32], eval_b[32], eval_c[32], eval_s1[32], eval_s2[32], eval_zw[32], eval_r[32];
} nzcb_proof;

/* ---- context ---------------------------------------------------------- */
int32_t nzcb_ctx_create(int32_t device_id, nzcb_ctx** out);
void nzcb_ctx_free(nzcb_ctx* ctx);
const char* nzcb_last_error(const nzcb_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
uint64_t nzcb_launch_count(const nzcb_ctx* ctx);
/* milliseconds of device time of the last timed entry point (CUDA events on the ctx stream) */
float nzcb_last_device_ms(const nzcb_ctx* ctx);

/* Single-proof latency mode: `world` contexts (one per GPU / process) prove the SAME proof; each
 * commits only its contiguous slice of the point range of every fixed-base MSM, and the partial sums (one 128-byte
 * XYZZ point per commitment and rank) are exchanged through `allgather(user, send, recv, bytes)` -- recv receives
 * world x bytes in rank order; the caller implements it with NCCL / P2P.  world = 1 turns the mode off. */
int32_t nzcb_ctx_set_msm_split(nzcb_ctx* ctx, int32_t rank, int32_t world,
                               int (*allgather)(void* user, const void* send, void* recv, size_t bytes), void* user);

/* The same mode with the exchange ON THE DEVICE: the partial sums stay in HBM, `ncclAllGather` runs on the context's
 * stream over a communicator of the library's own, a kernel adds them up in rank order and only the summed commitment
 * is read back.  NCCL is not a link-time dependency: `libnccl_path` names the libnccl.so.2 to dlopen (the one torch
 * bundles; NULL = the loader's search path).  Rank 0 draws the 128-byte id with nzcb_nccl_unique_id and the caller
 * broadcasts it (any channel); world = 1 turns the mode off and destroys the communicator. */
int32_t nzcb_nccl_unique_id(const char* libnccl_path, uint8_t id[128]);
int32_t nzcb_ctx_set_msm_split_nccl(nzcb_ctx* ctx, int32_t rank, int32_t world, const char* libnccl_path,
                                    const uint8_t id[128]);

/* integer-pipe microbenchmark (roofline denominators): kind 0 = IMAD,
 * 1 = IMAD.WIDE.U32, 2 = Fr Montgomery multiply, 3 = Fq multiply, 4 = IMAD.HI.U32,
 * 5 = Fr multiply, portable CIOS variant; result in ops/s */
int32_t nzcb_microbench(nzcb_ctx* ctx, int32_t kind, uint32_t iters, uint32_t blocks_per_sm, double* ops_per_s);
/* mixed-addition (XYZZ += affine, the body of the MSM accumulation) loop variants; result in additions/s */
int32_t nzcb_microbench_madd(nzcb_ctx* ctx, int32_t variant, uint32_t iters, uint32_t log_table, double* madds_per_s);
/* store -> barrier -> dependent load round trip of one CTA through global (kind 0) or shared (kind 1) memory: the
 * floor of one level of the witness interpreter; result in ns per level */
int32_t nzcb_microbench_level(nzcb_ctx* ctx, int32_t kind, uint32_t iters, double* ns_per_level);
/* device self-test: carry-chain multiply vs portable CIOS on n random operand pairs x 16 */
int32_t nzcb_selftest_mul(nzcb_ctx* ctx, uint32_t n, uint64_t* mismatches);

/* ---- primitives: ffjavascript Fr.fft / Fr.ifft / G1.multiExpAffine ------ */
/* In-place NTT of 2^log_n LEM elements, natural order in and out; inverse
 * includes the 1/N factor. */
int32_t nzcb_ntt_fr(nzcb_ctx* ctx, uint8_t* data_lem, uint32_t log_n, int32_t inverse);
/* sum_i scalars[i] * bases[i];  bases n x 64 B affine LEM, scalars n x 32 B LE
 * canonical (what multiExpAffine takes after fromMontgomery); result affine LEM. */
int32_t nzcb_msm_g1(nzcb_ctx* ctx, const uint8_t* bases_affine_lem, const uint8_t* scalars_le, size_t n,
                    uint8_t out_affine_lem[64]);
/* device-resident primitives for the roofline measurement (buffers from nzcb_dev_alloc) */
int32_t nzcb_dev_alloc(nzcb_ctx* ctx, size_t bytes, void** dptr);
int32_t nzcb_dev_free(nzcb_ctx* ctx, void* dptr);
int32_t nzcb_dev_upload(nzcb_ctx* ctx, void* dptr, const void* host, size_t bytes);
int32_t nzcb_dev_download(nzcb_ctx* ctx, void* host, const void* dptr, size_t bytes);
int32_t nzcb_ntt_fr_dev(nzcb_ctx* ctx, void* d_data_lem, uint32_t log_n, int32_t inverse);
int32_t nzcb_msm_g1_dev(nzcb_ctx* ctx, const void* d_bases_affine_lem, const void* d_scalars_le, size_t n,
                        uint8_t out_affine_lem[64]);

/* Fixed bases (the way the prover holds the zkey's SRS): the window shifts 2^(c w) P_i are precomputed once,
 * every MSM over the table then uses one bucket set and 13 digits per scalar at 2^21 points.
 * bases: n x 64 B affine LEM (host).  The table owns ~(254/c + 1) x n x 64 B of HBM. */
typedef struct nzcb_g1_table nzcb_g1_table;
int32_t nzcb_g1_table_create(nzcb_ctx* ctx, const uint8_t* bases_affine_lem, size_t n, nzcb_g1_table** out);
void nzcb_g1_table_free(nzcb_g1_table* t);
/* sum_i scalars[i] * bases[i] over the first n bases of the table; scalars n x 32 B LE (host) */
int32_t nzcb_msm_g1_table(nzcb_ctx* ctx, const nzcb_g1_table* t, const uint8_t* scalars_le, size_t n,
                          uint8_t out_affine_lem[64]);
/* K <= 4 MSMs over the same table as one batch (one sort, one accumulation launch, one reduction);
 * d_scalars[k] = device buffer of n[k] x 32 B LE; out = K x 64 B affine LEM (host) */
int32_t nzcb_msm_g1_table_dev(nzcb_ctx* ctx, const nzcb_g1_table* t, const void* const* d_scalars, const size_t* n,
                              int32_t K, uint8_t* out_affine_lem);

/* Lagrange-basis SRS: [L_i(tau)]G1, i < 2^log_n, from the monomial points [tau^j]G1, j < 2^log_n -- the group
 * inverse DFT of `snarkjs powersoftau prepare phase2` (ptau sections 12-15).  Both buffers n x 64 B affine LEM. */
int32_t nzcb_g1_lagrange_basis(nzcb_ctx* ctx, const uint8_t* srs_affine_lem, uint32_t log_n, uint8_t* out_affine_lem);

/* ---- SRS + setup: `snarkjs powersoftau new` / `plonk setup` roles --------- */
/* [tau^i]G1, i < count, affine LEM (insecure known-trapdoor SRS, Makefile:64-67 role) */
int32_t nzcb_srs_g1(nzcb_ctx* ctx, const uint8_t tau_le[32], size_t count, uint8_t* out_affine_lem);
/* snarkjs `plonk setup`: r1cs file bytes + SRS -> PLONK zkey v1
 * bytes.  Call with zkey_out = NULL to get the size in *zkey_len. */
int32_t nzcb_plonk_setup(nzcb_ctx* ctx, const uint8_t* r1cs, size_t r1cs_len, const uint8_t* srs_g1_lem,
                         size_t srs_count, const uint8_t x2_g2_lem[128], uint8_t* zkey_out, size_t* zkey_len);

/* the same from the .ptau file itself: `snarkjs plonk setup circuit.r1cs pot.ptau circuit.zkey`
 * (/root/reference/Makefile:55,60; README.md:41).  Takes tauG1[0 .. 2^cirPower + 6) of section 2 and tauG2[1] of
 * section 3; fails with snarkjs' messages "Powers of tau is not prepared." (no section 12) and "circuit too big
 * for this power of tau ceremony." */
int32_t nzcb_plonk_setup_ptau(nzcb_ctx* ctx, const uint8_t* r1cs, size_t r1cs_len, const uint8_t* ptau, size_t ptau_len,
                              uint8_t* zkey_out, size_t* zkey_len);
/* header of a .ptau (no GPU needed): power, ceremonyPower, points in the tauG1 section, whether
 * `powersoftau prepare phase2` has run (section 12 present); any pointer may be NULL */
int32_t nzcb_ptau_info(const uint8_t* ptau, size_t len, uint32_t* power, uint32_t* ceremony_power, uint64_t* n_tau_g1,
                       int32_t* prepared);

/* what the R1CS -> PLONK expansion of `plonk setup` yields for this r1cs: gate count, additions,
 * PLONK signal count and log2 of the domain (so the caller can size the SRS: 2^power + 6 points) */
int32_t nzcb_plonk_setup_info(nzcb_ctx* ctx, const uint8_t* r1cs, size_t r1cs_len, uint32_t* n_gates,
                              uint32_t* n_additions, uint32_t* plonk_vars, uint32_t* power);

/* ---- prover: snarkjs plonk.prove(zkey, wtns) ---------------------------- */
/* parse a PLONK zkey (v1, 14 sections) and make it device resident once */
int32_t nzcb_zkey_load(nzcb_ctx* ctx, const uint8_t* zkey, size_t len, nzcb_zkey** out);
void nzcb_zkey_free(nzcb_zkey* zk);
int32_t nzcb_zkey_info(const nzcb_zkey* zk, uint32_t* n_vars, uint32_t* n_public, uint32_t* domain_size,
                       uint32_t* n_additions, uint32_t* n_constraints);
/* blinders: b1..b9 as 9 x 32 B LE canonical (what Fr.random() would have
 * returned, in call order); NULL -> drawn from the OS CSPRNG.
 * public_le: nPublic x 32 B LE canonical (publicSignals). */
int32_t nzcb_plonk_prove(nzcb_ctx* ctx, const nzcb_zkey* zk, const uint8_t* wtns, size_t wtns_len,
                         const uint8_t* blinders_le, nzcb_proof* out, uint8_t* public_le);
/* B independent proofs on this ctx's GPU (multi-GPU = one ctx/process per GPU,
 * proofs sharded by the caller; no collective).  wtns[i] / wtns_len[i] per proof. */
int32_t nzcb_plonk_prove_batch(nzcb_ctx* ctx, const nzcb_zkey* zk, const uint8_t* const* wtns,
                               const size_t* wtns_len, const uint8_t* blinders_le /* B x 9 x 32 or NULL */,
                               size_t B, nzcb_proof* out, uint8_t* public_le /* B x nPublic x 32 */,
                               int32_t* status /* B */);
/* proof.json / public.json text exactly as snarkjs prints them (key order of A.2 step 6) */
int32_t nzcb_proof_to_json(const nzcb_proof* proof, char* buf, size_t* len);

/* ---- witness: circom_tester calculateWitness / snarkjs wtns.calculate ---- */
/* load a compiled witness program (nzcb .wprog bytes emitted by the circuit builder) */
int32_t nzcb_circuit_load(nzcb_ctx* ctx, const uint8_t* wprog, size_t len, nzcb_circuit** out);
void nzcb_circuit_free(nzcb_circuit* c);
int32_t nzcb_circuit_info(const nzcb_circuit* c, uint32_t* n_witness, uint32_t* n_inputs, uint32_t* n_outputs);
/* B passes: inputs_le is B x nInputs x 32 B LE canonical, in main-input
 * declaration order (the order circom puts them in the witness);
 * wtns_out (may be NULL) receives B x nWitness x 32 B LE canonical -- the
 * payload of .wtns section 2; status[i] = 0 or NZCB_E_ASSERT per pass (a failed
 * pass never fails the batch, test/quinSelector.js:66 semantics). */
int32_t nzcb_witness_batch(nzcb_ctx* ctx, const nzcb_circuit* c, const uint8_t* inputs_le, size_t B,
                           uint8_t* wtns_out, int32_t* status);

/* circom_runtime's input contract in front of nzcb_witness_batch (what `cir.calculateWitness({name: value, ...})`
 * does before any constraint runs, test/nzcp.js:42; un-vendored circom_runtime, /root/reference/yarn.lock:2496):
 * signals are addressed by the FNV-1a-64 hash of their name, array values flattened row-major.
 * `sym` is the circuit's input table (<name>.sym written by `python -m nzcb_circom_b200.circom`): "NZSY", u32 1,
 * u32 nSignals, u32 nInputs, nSignals x {u64 hash, u32 offset, u32 size}.  hashes[k] / counts[k] name the k-th
 * provided signal and how many values follow for it in values_le (32 B LE each, taken mod r like Fr.e).
 * Returns 0 and the nInputs x 32 B buffer nzcb_witness_batch takes, or NZCB_E_SIGNAL(code) with circom_runtime's
 * message in err.  Host only: no context, no GPU. */
uint64_t nzcb_fnv1a64(const char* name, size_t len);
int32_t nzcb_inputs_resolve(const uint8_t* sym, size_t sym_len, uint32_t n_signals, const uint64_t* hashes,
                            const uint32_t* counts, const uint8_t* values_le, uint8_t* inputs_out, char* err,
                            size_t err_len);
/* witness values (nWitness x 32 B LE canonical, as nzcb_witness_batch returns them) -> .wtns v2 file bytes
 * (snarkjs wtns.calculate's output; *len in = capacity, out = size; out == NULL: query the size).  Host only. */
int32_t nzcb_wtns_export(const uint8_t* witness_le, uint32_t n_witness, uint8_t* out, size_t* len);
/* `snarkjs zkey export verificationkey circuit.zkey verification_key.json` (/root/reference/Makefile:56,61):
 * the JSON text, JSON.stringify(vk, null, 1) layout, from the zkey header alone.  Host only. */
int32_t nzcb_vkey_to_json(const uint8_t* zkey, size_t zkey_len, char* buf, size_t* len);

/* The same for batches too large to bring every wire back (BASELINE.json configs[3]: 65,536 passes x 26 MB):
 * outputs_le (may be NULL) receives w[1 .. nOutputs] of every pass (B x nOutputs x 32 B LE), i.e. what
 * circom_tester's getDecoratedOutput / test/nzcp.js:62-68 reads; digest (may be NULL) receives one u64 per pass,
 * sum_{i < nWitness, k < 8} (limb_k(w_i) + 1) * splitmix64(8 i + k) mod 2^64 over ALL wires of the pass, computed
 * on the device; every pass i with i % sample_stride == 0 has its whole witness copied to wtns_sample_out
 * (ceil(B / sample_stride) x nWitness x 32 B; NULL: none). */
int32_t nzcb_witness_batch_ex(nzcb_ctx* ctx, const nzcb_circuit* c, const uint8_t* inputs_le, size_t B,
                              uint8_t* outputs_le, uint64_t* digest, size_t sample_stride, uint8_t* wtns_sample_out,
                              int32_t* status);
/* the same with the marshalled inputs already resident in HBM (nzcb_dev_alloc / nzcb_pass_ingest_batch_dev): the
 * device-resident witness throughput of bench.py's roofline_witness */
int32_t nzcb_witness_batch_ex_dev(nzcb_ctx* ctx, const nzcb_circuit* c, const void* d_inputs_le, size_t B,
                                  uint8_t* outputs_le, uint64_t* digest, size_t sample_stride, uint8_t* wtns_sample_out,
                                  int32_t* status);

/* snarkjs plonk.fullProve for B passes: witness program then prover, wires never leave HBM.
 * status[i] = 0, NZCB_E_ASSERT (pass rejected by the circuit) or a prover error; a failed
 * pass zeroes out[i] and never fails the batch. */
int32_t nzcb_plonk_fullprove_batch(nzcb_ctx* ctx, const nzcb_circuit* c, const nzcb_zkey* zk, const uint8_t* inputs_le,
                                   size_t B, const uint8_t* blinders_le /* B x 9 x 32 or NULL */, nzcb_proof* out,
                                   uint8_t* public_le /* B x nPublic x 32 */, int32_t* status /* B */);

/* same with the marshalled inputs already resident in HBM (buffer from nzcb_dev_alloc / nzcb_dev_upload):
 * the device-resident throughput figure of bench.py */
int32_t nzcb_plonk_fullprove_batch_dev(nzcb_ctx* ctx, const nzcb_circuit* c, const nzcb_zkey* zk,
                                       const void* d_inputs_le, size_t B, const uint8_t* blinders_le, nzcb_proof* out,
                                       uint8_t* public_le, int32_t* status);

/* ---- verifier: snarkjs plonk.verify(vk, publicSignals, proof) -------------------------------------------
 * (the check the exported Solidity verifier performs on chain, /root/reference/Makefile:56-57,
 * 61-62.)  The verification key comes from a zkey file's header (section 2) or from verification_key.json as
 * `snarkjs zkey export verificationkey` writes it. */
typedef struct nzcb_vkey nzcb_vkey;
int32_t nzcb_vkey_from_zkey(nzcb_ctx* ctx, const uint8_t* zkey, size_t len, nzcb_vkey** out);
int32_t nzcb_vkey_from_json(nzcb_ctx* ctx, const char* json, size_t len, nzcb_vkey** out);
void nzcb_vkey_free(nzcb_vkey* vk);
/* B proofs against one key, one warp per proof.  public_le: B x n_public x 32 B LE (reduced mod r as
 * Fr.fromObject does); valid[i] = 1 (plonk.verify -> true) or 0: a point off the curve, a coordinate or
 * evaluation out of range, n_public != vk.nPublic, or the pairing check e(-A1, X_2) e(B1, [1]_2) != 1. */
int32_t nzcb_plonk_verify_batch(nzcb_ctx* ctx, const nzcb_vkey* vk, const nzcb_proof* proofs, const uint8_t* public_le,
                                uint32_t n_public, size_t B, int32_t* valid);
/* ffjavascript curve.pairingEq(P_0, Q_0, .., P_{n-1}, Q_{n-1}): *result = 1 iff prod e(P_i, Q_i) == 1, n <= 32.
 * g1: n x 64 B affine LEM; g2: n x 128 B affine LEM (x.c0 x.c1 y.c0 y.c1, the zkey's X_2 layout; infinity = zeros).
 * gt_out_lem (may be NULL): the product after the final exponentiation, 12 x 32 B LEM as the coefficients
 * (c0, c1) of 1, w, .., w^5 in Fq2[w]/(w^6 - (9 + u)). */
int32_t nzcb_pairing_eq(nzcb_ctx* ctx, const uint8_t* g1_affine_lem, const uint8_t* g2_affine_lem, uint32_t n,
                        int32_t* result, uint8_t* gt_out_lem);
/* [tau]G2 of the insecure known-trapdoor SRS (the tauG2 point a .ptau carries and `plonk setup` copies into the
 * zkey header as X_2); 128 B affine LEM */
int32_t nzcb_srs_g2(nzcb_ctx* ctx, const uint8_t tau_le[32], uint8_t out_affine_lem[128]);
/* `snarkjs zkey export soliditycalldata` (plonk): 0x<proof bytes>,["0x<pub>",..]; NULL buf -> size in *len */
int32_t nzcb_proof_to_calldata(const nzcb_proof* proof, const uint8_t* public_le, uint32_t n_public, char* buf,
                               size_t* len);

/* ---- pass ingest: the host helpers the reference's tests run before calculateWitness ---------------------
 * getCOSE + encodeToBeSigned (/root/reference/test/helpers/nzcp.js:9-24 base32ToBytes, :58-105 decodeCBORStream,
 * :141-172 decodeBytes / decodeCOSE, :180-206 encodeToBeSigned) and the input object of test/nzcp.js:36-41
 * (fitBytes, bufferToBitArray, evmRearrangeBytes: test/helpers/utils.js:49,2,87) for B passes on the device.
 * uris: the pass URIs ("NZCP:/1/<base32>", ASCII) back to back, pass i = uris[uri_off[i] .. uri_off[i+1]); the
 * first 8 characters are skipped unchecked, as decodeBytes does.  data20: B x 20 pass-through bytes (origData of
 * test/nzcp.js:36) or NULL = zeros.  max_len: 314 (nzcp_example) / 351 (nzcp_live), at most 1024.
 * tbs_out (may be NULL): B x max_len = fitBytes(ToBeSigned, max_len); tbs_len (may be NULL): the TRUE lengths;
 * inputs_le (may be NULL): B x (8 max_len + 161) x 32 B LE -- the main inputs in declaration order, what
 * nzcb_witness_batch / nzcb_plonk_fullprove_batch take.  status[i] = 0 or NZCB_E_INVALID ("invalid data": the JS
 * helpers throw); a rejected pass yields zeros, tbs_len 0 and toBeSignedLen = 0xFFFF in its inputs (which every
 * circuit rejects) and never fails the batch.  URIs longer than 4104 characters are rejected. */
int32_t nzcb_pass_ingest_batch(nzcb_ctx* ctx, const uint8_t* uris, const uint32_t* uri_off, size_t B,
                               const uint8_t* data20, uint32_t max_len, uint8_t* tbs_out, uint32_t* tbs_len,
                               uint8_t* inputs_le, int32_t* status);
/* the same with the marshalled inputs left in HBM (d_inputs_le from nzcb_dev_alloc, B x (8 max_len + 161) x 32 B),
 * ready for nzcb_plonk_fullprove_batch_dev */
int32_t nzcb_pass_ingest_batch_dev(nzcb_ctx* ctx, const uint8_t* uris, const uint32_t* uri_off, size_t B,
                                   const uint8_t* data20, uint32_t max_len, void* d_inputs_le, int32_t* status);
/* pass URIs -> proofs: ingest, witness program and prover fused on the device; ~0.6 KB per pass cross PCIe
 * instead of the 95 KB of marshalled inputs.  status[i] = 0, NZCB_E_INVALID (undecodable pass), NZCB_E_ASSERT
 * (rejected by the circuit) or a prover error. */
int32_t nzcb_plonk_fullprove_uri_batch(nzcb_ctx* ctx, const nzcb_circuit* c, const nzcb_zkey* zk, const uint8_t* uris,
                                       const uint32_t* uri_off, size_t B, const uint8_t* data20, uint32_t max_len,
                                       const uint8_t* blinders_le /* B x 9 x 32 or NULL */, nzcb_proof* out,
                                       uint8_t* public_le /* B x nPublic x 32 */, int32_t* status /* B */);

/* CUDA-event timing of the dominant kernel (MSM bucket accumulation) on the ctx stream.
 * nzcb_profile(ctx, 1) starts collecting; nzcb_profile_read returns launches, summed device ms and the
 * algorithmic modmul count of those launches (160 per MSM point) and resets. */
int32_t nzcb_profile(nzcb_ctx* ctx, int32_t enable);
int32_t nzcb_profile_read(nzcb_ctx* ctx, uint64_t* launches, double* total_ms, double* alg_modmul);
/* bucket additions (XYZZ += affine, 10 modmul each) the timed launches actually executed; read BEFORE
 * nzcb_profile_read, which resets.  Smaller than the algorithmic count when scalars are small (round 1). */
int32_t nzcb_profile_entries(nzcb_ctx* ctx, double* additions);

#ifdef __cplusplus
}
#endif
#endif /* NZCB_H */
