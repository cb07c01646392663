/* nzcb_napi.c -- N-API shim over include/nzcb.h: the binding a maintainer of noway/nzcb-circom adds so that
 * test/nzcp.js (circom_tester.wasm / calculateWitness) and the snarkjs calls of the Makefile recipes
 * (plonk setup / prove / fullProve / verify, /root/reference/Makefile:54-62) run on libnzcb.so.
 *
 * NOT BUILT OR RUN IN THIS REPOSITORY'S ENVIRONMENT: the image has no Node and no node_api.h.  It is
 * type-checked against a stub of the N-API declarations it uses (tests/hostcheck/node_api_stub.h,
 * tests/test_abi.py); the same C ABI is exercised for real through ctypes (nzcb_circom_b200/_lib.py).
 *
 * Conventions: every handle (ctx-bound zkey / circuit / vkey) is a napi external with a finalizer; buffers are
 * Node Buffers in exactly the layouts of include/nzcb.h; errors become JS exceptions carrying nzcb_last_error.
 */
#include <node_api.h>
#include <stdlib.h>
#include <string.h>
#include "nzcb.h"

static nzcb_ctx* g_ctx;

#define NAPI_OK(call)                                                   \
    do {                                                                \
        if ((call) != napi_ok) {                                        \
            napi_throw_error(env, NULL, "nzcb: N-API call failed");     \
            return NULL;                                                \
        }                                                               \
    } while (0)
#define NZCB_OK_OR_THROW(rc)                                            \
    do {                                                                \
        if ((rc) != 0) {                                                \
            napi_throw_error(env, NULL, nzcb_last_error(g_ctx));        \
            return NULL;                                                \
        }                                                               \
    } while (0)

static int is_nullish(napi_env env, napi_value v) {
    napi_valuetype t;
    if (napi_typeof(env, v, &t) != napi_ok) return 1;
    return t == napi_null || t == napi_undefined;
}

static void free_zkey(napi_env env, void* data, void* hint) { (void)env; (void)hint; nzcb_zkey_free((nzcb_zkey*)data); }
static void free_circuit(napi_env env, void* data, void* hint) { (void)env; (void)hint; nzcb_circuit_free((nzcb_circuit*)data); }
static void free_vkey(napi_env env, void* data, void* hint) { (void)env; (void)hint; nzcb_vkey_free((nzcb_vkey*)data); }

/* loadZkey(Buffer zkeyFileBytes) -> external: parsed and device resident once per process */
static napi_value LoadZkey(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1]; uint8_t* buf; size_t len; nzcb_zkey* zk; napi_value out;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    NAPI_OK(napi_get_buffer_info(env, argv[0], (void**)&buf, &len));
    NZCB_OK_OR_THROW(nzcb_zkey_load(g_ctx, buf, len, &zk));
    NAPI_OK(napi_create_external(env, zk, free_zkey, NULL, &out));
    return out;
}

/* loadCircuit(Buffer witnessProgram) -> external */
static napi_value LoadCircuit(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1]; uint8_t* buf; size_t len; nzcb_circuit* c; napi_value out;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    NAPI_OK(napi_get_buffer_info(env, argv[0], (void**)&buf, &len));
    NZCB_OK_OR_THROW(nzcb_circuit_load(g_ctx, buf, len, &c));
    NAPI_OK(napi_create_external(env, c, free_circuit, NULL, &out));
    return out;
}

/* loadVkey(string verificationKeyJson) -> external */
static napi_value LoadVkey(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1]; size_t len; char* js; nzcb_vkey* vk; napi_value out; int32_t rc;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    NAPI_OK(napi_get_value_string_utf8(env, argv[0], NULL, 0, &len));
    js = (char*)malloc(len + 1);
    if (!js) { napi_throw_error(env, NULL, "nzcb: out of memory"); return NULL; }
    NAPI_OK(napi_get_value_string_utf8(env, argv[0], js, len + 1, &len));
    rc = nzcb_vkey_from_json(g_ctx, js, len, &vk);
    free(js);
    NZCB_OK_OR_THROW(rc);
    NAPI_OK(napi_create_external(env, vk, free_vkey, NULL, &out));
    return out;
}

/* prove(zkey, Buffer wtnsFileBytes, Buffer|null blinders(9 x 32 LE)) -> { proof: Buffer(800), publicSignals: Buffer }
 * = snarkjs.plonk.prove(zkeyFile, wtnsFile) */
static napi_value Prove(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value argv[3]; nzcb_zkey* zk; uint8_t *wtns, *bl = NULL, *pub; size_t wlen, blen;
    uint32_t n_pub = 0; nzcb_proof proof; napi_value out, p, s; int32_t rc;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    NAPI_OK(napi_get_value_external(env, argv[0], (void**)&zk));
    NAPI_OK(napi_get_buffer_info(env, argv[1], (void**)&wtns, &wlen));
    if (argc > 2 && !is_nullish(env, argv[2])) {
        NAPI_OK(napi_get_buffer_info(env, argv[2], (void**)&bl, &blen));
        if (blen != 9 * 32) { napi_throw_error(env, NULL, "nzcb: blinders must be 9 x 32 bytes"); return NULL; }
    }
    NZCB_OK_OR_THROW(nzcb_zkey_info(zk, NULL, &n_pub, NULL, NULL, NULL));
    pub = (uint8_t*)malloc(32 * (size_t)(n_pub ? n_pub : 1));
    if (!pub) { napi_throw_error(env, NULL, "nzcb: out of memory"); return NULL; }
    rc = nzcb_plonk_prove(g_ctx, zk, wtns, wlen, bl, &proof, pub);
    if (rc != 0) { free(pub); NZCB_OK_OR_THROW(rc); }
    NAPI_OK(napi_create_object(env, &out));
    NAPI_OK(napi_create_buffer_copy(env, sizeof proof, &proof, NULL, &p));
    NAPI_OK(napi_create_buffer_copy(env, 32 * (size_t)n_pub, pub, NULL, &s));
    free(pub);
    NAPI_OK(napi_set_named_property(env, out, "proof", p));
    NAPI_OK(napi_set_named_property(env, out, "publicSignals", s));
    return out;
}

/* calculateWitness(circuit, Buffer inputs(B x nInputs x 32 LE), B) -> { witness: Buffer(B x nWitness x 32), status: Int32Array-like Buffer }
 * = circom_tester calculateWitness for B inputs; a failed assert sets status[i] = NZCB_E_ASSERT */
static napi_value CalculateWitness(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value argv[3]; nzcb_circuit* c; uint8_t* in; size_t ilen; uint32_t B, n_w, n_in, n_out;
    napi_value out, w, st; uint8_t* wbuf; int32_t* sbuf;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    NAPI_OK(napi_get_value_external(env, argv[0], (void**)&c));
    NAPI_OK(napi_get_buffer_info(env, argv[1], (void**)&in, &ilen));
    NAPI_OK(napi_get_value_uint32(env, argv[2], &B));
    NZCB_OK_OR_THROW(nzcb_circuit_info(c, &n_w, &n_in, &n_out));
    if (ilen != (size_t)B * n_in * 32) { napi_throw_error(env, NULL, "nzcb: inputs buffer has the wrong size"); return NULL; }
    NAPI_OK(napi_create_buffer(env, (size_t)B * n_w * 32, (void**)&wbuf, &w));
    NAPI_OK(napi_create_buffer(env, (size_t)B * sizeof(int32_t), (void**)&sbuf, &st));
    NZCB_OK_OR_THROW(nzcb_witness_batch(g_ctx, c, in, B, wbuf, sbuf));
    NAPI_OK(napi_create_object(env, &out));
    NAPI_OK(napi_set_named_property(env, out, "witness", w));
    NAPI_OK(napi_set_named_property(env, out, "status", st));
    return out;
}

static napi_value proofs_result(napi_env env, uint32_t B, uint32_t n_pub, nzcb_proof* proofs, uint8_t* pub, int32_t* status) {
    napi_value out, p, s, st;
    NAPI_OK(napi_create_object(env, &out));
    NAPI_OK(napi_create_buffer_copy(env, (size_t)B * sizeof(nzcb_proof), proofs, NULL, &p));
    NAPI_OK(napi_create_buffer_copy(env, (size_t)B * n_pub * 32, pub, NULL, &s));
    NAPI_OK(napi_create_buffer_copy(env, (size_t)B * sizeof(int32_t), status, NULL, &st));
    NAPI_OK(napi_set_named_property(env, out, "proofs", p));
    NAPI_OK(napi_set_named_property(env, out, "publicSignals", s));
    NAPI_OK(napi_set_named_property(env, out, "status", st));
    return out;
}

/* fullProveBatch(circuit, zkey, Buffer inputs, B, Buffer|null blinders) -> { proofs, publicSignals, status }
 * = snarkjs.plonk.fullProve for B inputs, witness and prover fused on the device */
static napi_value FullProveBatch(napi_env env, napi_callback_info info) {
    size_t argc = 5; napi_value argv[5]; nzcb_circuit* c; nzcb_zkey* zk; uint8_t *in, *bl = NULL; size_t ilen, blen;
    uint32_t B, n_pub = 0; napi_value out; int32_t rc;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    NAPI_OK(napi_get_value_external(env, argv[0], (void**)&c));
    NAPI_OK(napi_get_value_external(env, argv[1], (void**)&zk));
    NAPI_OK(napi_get_buffer_info(env, argv[2], (void**)&in, &ilen));
    NAPI_OK(napi_get_value_uint32(env, argv[3], &B));
    if (argc > 4 && !is_nullish(env, argv[4])) NAPI_OK(napi_get_buffer_info(env, argv[4], (void**)&bl, &blen));
    NZCB_OK_OR_THROW(nzcb_zkey_info(zk, NULL, &n_pub, NULL, NULL, NULL));
    {   /* the library reads B * nInputs * 32 and B * 9 * 32 bytes: a short Buffer must not become an out-of-bounds read */
        uint32_t n_in = 0;
        NZCB_OK_OR_THROW(nzcb_circuit_info(c, NULL, &n_in, NULL));
        if (ilen != (size_t)B * n_in * 32) { napi_throw_error(env, NULL, "nzcb: inputs must hold B * nInputs 32-byte values"); return NULL; }
        if (bl && blen != (size_t)B * 9 * 32) { napi_throw_error(env, NULL, "nzcb: blinders must hold B * 9 32-byte values"); return NULL; }
    }
    {
        nzcb_proof* proofs = (nzcb_proof*)malloc((size_t)(B ? B : 1) * sizeof(nzcb_proof));
        uint8_t* pub = (uint8_t*)malloc((size_t)(B ? B : 1) * (n_pub ? n_pub : 1) * 32);
        int32_t* status = (int32_t*)malloc((size_t)(B ? B : 1) * sizeof(int32_t));
        if (!proofs || !pub || !status) { free(proofs); free(pub); free(status); napi_throw_error(env, NULL, "nzcb: out of memory"); return NULL; }
        rc = nzcb_plonk_fullprove_batch(g_ctx, c, zk, in, B, bl, proofs, pub, status);
        out = rc == 0 ? proofs_result(env, B, n_pub, proofs, pub, status) : NULL;
        free(proofs); free(pub); free(status);
        NZCB_OK_OR_THROW(rc);
    }
    return out;
}

/* fullProveURIs(circuit, zkey, Buffer uris, Buffer uriOffsets(u32 x (B+1)), Buffer|null data(B x 20), maxLen)
 * -> { proofs, publicSignals, status }: pass URIs in, proofs out (test/helpers/nzcp.js getCOSE + encodeToBeSigned,
 * test/nzcp.js:36-41 and plonk.fullProve on the device) */
static napi_value FullProveURIs(napi_env env, napi_callback_info info) {
    size_t argc = 6; napi_value argv[6]; nzcb_circuit* c; nzcb_zkey* zk; uint8_t *uris, *data = NULL; uint32_t* off;
    size_t ulen, olen, dlen; uint32_t max_len, n_pub = 0, B; napi_value out; int32_t rc;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    NAPI_OK(napi_get_value_external(env, argv[0], (void**)&c));
    NAPI_OK(napi_get_value_external(env, argv[1], (void**)&zk));
    NAPI_OK(napi_get_buffer_info(env, argv[2], (void**)&uris, &ulen));
    NAPI_OK(napi_get_buffer_info(env, argv[3], (void**)&off, &olen));
    if (!is_nullish(env, argv[4])) NAPI_OK(napi_get_buffer_info(env, argv[4], (void**)&data, &dlen));
    NAPI_OK(napi_get_value_uint32(env, argv[5], &max_len));
    if (olen < 4 || olen % 4) { napi_throw_error(env, NULL, "nzcb: uriOffsets must hold B + 1 u32 values"); return NULL; }
    B = (uint32_t)(olen / 4 - 1);
    {   /* offsets ascending and inside the URI buffer; one 20-byte data value per pass */
        uint32_t i;
        for (i = 0; i < B; i++)
            if (off[i] > off[i + 1]) { napi_throw_error(env, NULL, "nzcb: uriOffsets must be ascending"); return NULL; }
        if (off[B] > ulen) { napi_throw_error(env, NULL, "nzcb: uriOffsets run past the end of uris"); return NULL; }
        if (data && dlen != (size_t)B * 20) { napi_throw_error(env, NULL, "nzcb: data must hold B * 20 bytes"); return NULL; }
    }
    NZCB_OK_OR_THROW(nzcb_zkey_info(zk, NULL, &n_pub, NULL, NULL, NULL));
    {
        nzcb_proof* proofs = (nzcb_proof*)malloc((size_t)(B ? B : 1) * sizeof(nzcb_proof));
        uint8_t* pub = (uint8_t*)malloc((size_t)(B ? B : 1) * (n_pub ? n_pub : 1) * 32);
        int32_t* status = (int32_t*)malloc((size_t)(B ? B : 1) * sizeof(int32_t));
        if (!proofs || !pub || !status) { free(proofs); free(pub); free(status); napi_throw_error(env, NULL, "nzcb: out of memory"); return NULL; }
        rc = nzcb_plonk_fullprove_uri_batch(g_ctx, c, zk, uris, off, B, data, max_len, NULL, proofs, pub, status);
        out = rc == 0 ? proofs_result(env, B, n_pub, proofs, pub, status) : NULL;
        free(proofs); free(pub); free(status);
        NZCB_OK_OR_THROW(rc);
    }
    return out;
}

/* verify(vkey, Buffer publicSignals(n x 32 LE), Buffer proof(800)) -> boolean = snarkjs.plonk.verify */
static napi_value Verify(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value argv[3]; nzcb_vkey* vk; uint8_t *pub, *proof; size_t plen, prlen; int32_t ok = 0; napi_value out;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    NAPI_OK(napi_get_value_external(env, argv[0], (void**)&vk));
    NAPI_OK(napi_get_buffer_info(env, argv[1], (void**)&pub, &plen));
    NAPI_OK(napi_get_buffer_info(env, argv[2], (void**)&proof, &prlen));
    if (prlen != sizeof(nzcb_proof) || plen % 32) { napi_throw_error(env, NULL, "nzcb: bad proof or publicSignals buffer"); return NULL; }
    NZCB_OK_OR_THROW(nzcb_plonk_verify_batch(g_ctx, vk, (const nzcb_proof*)proof, pub, (uint32_t)(plen / 32), 1, &ok));
    NAPI_OK(napi_get_boolean(env, ok == 1, &out));
    return out;
}

/* proofToJson(Buffer proof(800)) -> string, exactly JSON.stringify(proof, null, 1) of snarkjs */
static napi_value ProofToJson(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1]; uint8_t* proof; size_t prlen, n = 0; char* buf; napi_value out;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    NAPI_OK(napi_get_buffer_info(env, argv[0], (void**)&proof, &prlen));
    if (prlen != sizeof(nzcb_proof)) { napi_throw_error(env, NULL, "nzcb: a proof is 800 bytes"); return NULL; }
    nzcb_proof_to_json((const nzcb_proof*)proof, NULL, &n);
    buf = (char*)malloc(n);
    if (!buf || nzcb_proof_to_json((const nzcb_proof*)proof, buf, &n) != 0) { free(buf); napi_throw_error(env, NULL, "nzcb: proof_to_json failed"); return NULL; }
    NAPI_OK(napi_create_string_utf8(env, buf, strlen(buf), &out));
    free(buf);
    return out;
}

/* toBeSigned(Buffer uris, Buffer uriOffsets, maxLen) -> { toBeSigned: Buffer(B x maxLen), length: Buffer(u32 x B), status }
 * = encodeToBeSigned(getCOSE(uri)) of test/helpers/nzcp.js for B pass URIs */
static napi_value ToBeSigned(napi_env env, napi_callback_info info) {
    size_t argc = 3; napi_value argv[3]; uint8_t* uris; uint32_t* off; size_t ulen, olen; uint32_t max_len, B;
    napi_value out, t, l, st; uint8_t* tbuf; uint32_t* lbuf; int32_t* sbuf;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    NAPI_OK(napi_get_buffer_info(env, argv[0], (void**)&uris, &ulen));
    NAPI_OK(napi_get_buffer_info(env, argv[1], (void**)&off, &olen));
    NAPI_OK(napi_get_value_uint32(env, argv[2], &max_len));
    if (olen < 4 || olen % 4) { napi_throw_error(env, NULL, "nzcb: uriOffsets must hold B + 1 u32 values"); return NULL; }
    B = (uint32_t)(olen / 4 - 1);
    NAPI_OK(napi_create_buffer(env, (size_t)B * max_len, (void**)&tbuf, &t));
    NAPI_OK(napi_create_buffer(env, (size_t)B * 4, (void**)&lbuf, &l));
    NAPI_OK(napi_create_buffer(env, (size_t)B * 4, (void**)&sbuf, &st));
    NZCB_OK_OR_THROW(nzcb_pass_ingest_batch(g_ctx, uris, off, B, NULL, max_len, tbuf, lbuf, NULL, sbuf));
    NAPI_OK(napi_create_object(env, &out));
    NAPI_OK(napi_set_named_property(env, out, "toBeSigned", t));
    NAPI_OK(napi_set_named_property(env, out, "length", l));
    NAPI_OK(napi_set_named_property(env, out, "status", st));
    return out;
}

/* resolveInputs(Buffer sym, Buffer hashes(u64 x n), Buffer counts(u32 x n), Buffer values(32 B LE each), nInputs)
 * -> Buffer inputs (nInputs x 32): circom_runtime's setInputSignal loop; throws its messages ("Signal not found", ...) */
static napi_value ResolveInputs(napi_env env, napi_callback_info info) {
    size_t argc = 5; napi_value argv[5]; uint8_t *sym, *values, *outbuf; uint64_t* hashes; uint32_t* counts;
    size_t slen, hlen, clen, vlen, i, total = 0; uint32_t n_in; napi_value out; char err[128]; int32_t rc;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    NAPI_OK(napi_get_buffer_info(env, argv[0], (void**)&sym, &slen));
    NAPI_OK(napi_get_buffer_info(env, argv[1], (void**)&hashes, &hlen));
    NAPI_OK(napi_get_buffer_info(env, argv[2], (void**)&counts, &clen));
    NAPI_OK(napi_get_buffer_info(env, argv[3], (void**)&values, &vlen));
    NAPI_OK(napi_get_value_uint32(env, argv[4], &n_in));
    if (hlen % 8 || clen != hlen / 2) { napi_throw_error(env, NULL, "nzcb: hashes / counts length mismatch"); return NULL; }
    for (i = 0; i < clen / 4; i++) total += counts[i];
    if (vlen != total * 32) { napi_throw_error(env, NULL, "nzcb: values must hold 32 bytes per provided value"); return NULL; }
    NAPI_OK(napi_create_buffer(env, (size_t)n_in * 32, (void**)&outbuf, &out));
    rc = nzcb_inputs_resolve(sym, slen, (uint32_t)(hlen / 8), hashes, counts, values, outbuf, err, sizeof err);
    if (rc != 0) { napi_throw_error(env, NULL, err); return NULL; }
    return out;
}

/* wtnsExport(Buffer witness(n x 32 LE)) -> Buffer: the .wtns file snarkjs wtns.calculate writes */
static napi_value WtnsExport(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1]; uint8_t *w, *outbuf; size_t wlen, n = 0; napi_value out;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    NAPI_OK(napi_get_buffer_info(env, argv[0], (void**)&w, &wlen));
    if (wlen % 32) { napi_throw_error(env, NULL, "nzcb: a witness is n x 32 bytes"); return NULL; }
    nzcb_wtns_export(w, (uint32_t)(wlen / 32), NULL, &n);
    NAPI_OK(napi_create_buffer(env, n, (void**)&outbuf, &out));
    if (nzcb_wtns_export(w, (uint32_t)(wlen / 32), outbuf, &n) != 0) { napi_throw_error(env, NULL, "nzcb: wtns export failed"); return NULL; }
    return out;
}

/* vkeyToJson(Buffer zkey) -> string: `snarkjs zkey export verificationkey` (Makefile:56,61) */
static napi_value VkeyToJson(napi_env env, napi_callback_info info) {
    size_t argc = 1; napi_value argv[1]; uint8_t* zk; size_t zlen, n = 0; char* buf; napi_value out;
    NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    NAPI_OK(napi_get_buffer_info(env, argv[0], (void**)&zk, &zlen));
    if (nzcb_vkey_to_json(zk, zlen, NULL, &n) != 0) { napi_throw_error(env, NULL, "nzcb: not a PLONK zkey"); return NULL; }
    buf = (char*)malloc(n);
    if (!buf || nzcb_vkey_to_json(zk, zlen, buf, &n) != 0) { free(buf); napi_throw_error(env, NULL, "nzcb: vkey_to_json failed"); return NULL; }
    NAPI_OK(napi_create_string_utf8(env, buf, strlen(buf), &out));
    free(buf);
    return out;
}

static napi_value Init(napi_env env, napi_value exports) {
    static const struct { const char* name; napi_callback fn; } fns[] = {
        {"loadZkey", LoadZkey}, {"loadCircuit", LoadCircuit}, {"loadVkey", LoadVkey}, {"prove", Prove},
        {"calculateWitness", CalculateWitness}, {"fullProveBatch", FullProveBatch}, {"fullProveURIs", FullProveURIs},
        {"verify", Verify}, {"proofToJson", ProofToJson}, {"toBeSigned", ToBeSigned}, {"resolveInputs", ResolveInputs},
        {"wtnsExport", WtnsExport}, {"vkeyToJson", VkeyToJson}};
    size_t i;
    if (nzcb_ctx_create(0, &g_ctx) != 0) {
        napi_throw_error(env, NULL, nzcb_last_error(NULL));  /* "no CUDA device available ...; there is no CPU fallback" */
        return NULL;
    }
    for (i = 0; i < sizeof fns / sizeof fns[0]; i++) {
        napi_value f;
        NAPI_OK(napi_create_function(env, fns[i].name, NAPI_AUTO_LENGTH, fns[i].fn, NULL, &f));
        NAPI_OK(napi_set_named_property(env, exports, fns[i].name, f));
    }
    return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
