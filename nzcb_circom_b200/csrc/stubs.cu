// Entry points still to be implemented this round; each fails loudly.
#include "common.cuh"
#define NZ_STUB(ctx, name) return (ctx) ? (ctx)->fail(NZCB_E_INVALID, name ": not implemented yet") : NZCB_E_INVALID
extern "C" {
int32_t nzcb_srs_g1(nzcb_ctx* ctx, const uint8_t*, size_t, uint8_t*) { NZ_STUB(ctx, "nzcb_srs_g1"); }
int32_t nzcb_plonk_setup(nzcb_ctx* ctx, const uint8_t*, size_t, const uint8_t*, size_t, const uint8_t*, uint8_t*, size_t*) { NZ_STUB(ctx, "nzcb_plonk_setup"); }
int32_t nzcb_circuit_load(nzcb_ctx* ctx, const uint8_t*, size_t, nzcb_circuit**) { NZ_STUB(ctx, "nzcb_circuit_load"); }
void nzcb_circuit_free(nzcb_circuit*) {}
int32_t nzcb_circuit_info(const nzcb_circuit*, uint32_t*, uint32_t*, uint32_t*) { return NZCB_E_INVALID; }
int32_t nzcb_witness_batch(nzcb_ctx* ctx, const nzcb_circuit*, const uint8_t*, size_t, uint8_t*, int32_t*) { NZ_STUB(ctx, "nzcb_witness_batch"); }
}
