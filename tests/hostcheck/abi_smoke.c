/* Links libnzcb.so from plain C (no Python, no GPU needed for these entries): the host-only functions of
 * include/nzcb.h are called for real, the GPU entry points are only referenced so that the link resolves them.
 * Test infrastructure: tests/test_boundary_formats.py builds and runs it. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "nzcb.h"

int main(void) {
    /* FNV-1a 64 known answers */
    if (nzcb_fnv1a64("", 0) != 0xCBF29CE484222325ull || nzcb_fnv1a64("a", 1) != 0xAF63DC4C8601EC8Cull) return 1;
    /* a two-signal table: x[2] at 0, y at 2 */
    unsigned char sym[16 + 32];
    unsigned int hdr[3] = {1, 2, 3};
    unsigned long long hx = nzcb_fnv1a64("x", 1), hy = nzcb_fnv1a64("y", 1);
    unsigned int ox[2] = {0, 2}, oy[2] = {2, 1};
    memcpy(sym, "NZSY", 4); memcpy(sym + 4, hdr, 12);
    memcpy(sym + 16, &hx, 8); memcpy(sym + 24, ox, 8);
    memcpy(sym + 32, &hy, 8); memcpy(sym + 40, oy, 8);
    uint64_t hashes[2] = {hy, hx};
    uint32_t counts[2] = {1, 2};
    unsigned char vals[128] = {0}, out[96];
    char err[128];
    vals[0] = 7; vals[32] = 8; vals[64] = 9;
    if (nzcb_inputs_resolve(sym, sizeof sym, 2, hashes, counts, vals, out, err, sizeof err) != 0) return 2;
    if (out[0] != 8 || out[32] != 9 || out[64] != 7) return 3;
    hashes[0] = hx; hashes[1] = hy;   /* x first, with one value too many, while y is still unset */
    counts[0] = 3; counts[1] = 1;
    if (nzcb_inputs_resolve(sym, sizeof sym, 2, hashes, counts, vals, out, err, sizeof err) != NZCB_E_SIGNAL(6)) return 4;
    if (strcmp(err, "Input signal array access exceeds the size") != 0) return 5;
    /* .wtns writer: size query then fill */
    size_t n = 0;
    if (nzcb_wtns_export(vals, 3, NULL, &n) != 0 || n != 12 + 12 + 40 + 12 + 96) return 6;
    unsigned char* f = malloc(n);
    if (!f || nzcb_wtns_export(vals, 3, f, &n) != 0 || memcmp(f, "wtns", 4) != 0) return 7;
    free(f);
    /* proof.json of an all-zero proof: nine points at infinity */
    nzcb_proof p;
    memset(&p, 0, sizeof p);
    n = 0;
    if (nzcb_proof_to_json(&p, NULL, &n) != 0 || n < 100) return 8;
    /* the GPU entry points resolve at link time; without a device the context refuses loudly (no CPU fallback) */
    nzcb_ctx* ctx = NULL;
    int32_t rc = nzcb_ctx_create(0, &ctx);
    if (rc == 0) {
        printf("context on cuda:0 created\n");
        nzcb_ctx_free(ctx);
    } else {
        printf("no device: %s\n", nzcb_last_error(NULL));
    }
    void* keep[] = {(void*)nzcb_plonk_prove, (void*)nzcb_plonk_fullprove_batch, (void*)nzcb_witness_batch,
                    (void*)nzcb_plonk_verify_batch, (void*)nzcb_msm_g1, (void*)nzcb_ntt_fr};
    if (!keep[0]) return 9;
    printf("abi_smoke ok\n");
    return 0;
}
