// .ptau reader for `snarkjs plonk setup circuit.r1cs pot.ptau circuit.zkey` (/root/reference/Makefile:55,60; the
// ceremony file of /root/reference/README.md:41 or the local one of Makefile:64-67).  iden3 binfile "ptau" v1
// (SURVEY.md A.4): section 1 = n8, prime, power, ceremonyPower; section 2 = tauG1, 2^(power+1) - 1 points of 64 B
// (x || y, little-endian Montgomery); section 3 = tauG2, 2^power points of 128 B; section 12 = the Lagrange-basis
// points `powersoftau prepare phase2` appends.  plonk setup takes the first n + 6 tauG1 points (zkey section 14 and
// every key commitment) and tauG2[1] = [tau]_2 (X_2).  snarkjs commits with section 12's points; here the
// commitments are MSMs of the coefficient form over tauG1 -- the same group elements -- so section 12 is only
// required to be present, as snarkjs requires ("Powers of tau is not prepared.").
#include "common.cuh"

using namespace nzcb;

namespace {
struct PtauView {
    uint32_t power = 0, ceremony_power = 0;
    const uint8_t* tau_g1 = nullptr;
    uint64_t n_tau_g1 = 0;
    const uint8_t* tau_g2 = nullptr;
    uint64_t n_tau_g2 = 0;
    bool prepared = false;
};

// returns nullptr on success, else the message
const char* ptau_parse(const uint8_t* d, size_t len, PtauView* v) {
    if (!d || len < 12 || memcmp(d, "ptau", 4) != 0) return "ptau file: bad magic";
    uint32_t version, nsec;
    memcpy(&version, d + 4, 4);
    memcpy(&nsec, d + 8, 4);
    if (version > 1) return "ptau file: unsupported version";
    size_t pos = 12;
    bool have_hdr = false;
    for (uint32_t s = 0; s < nsec; s++) {
        if (pos + 12 > len) return "ptau file: truncated section table";
        uint32_t id;
        uint64_t size;
        memcpy(&id, d + pos, 4);
        memcpy(&size, d + pos + 4, 8);
        pos += 12;
        if (size > len - pos) return "ptau file: a section overruns the file";
        if (id == 1) {
            uint32_t n8;
            if (size < 4) return "ptau file: bad header";
            memcpy(&n8, d + pos, 4);
            if (n8 != 32 || size < 4 + 32 + 8) return "ptau file: not a 256-bit curve";
            const Fq q = Fq::modulus();
            if (memcmp(d + pos + 4, q.v, 32) != 0) return "ptau file: curve is not bn128";
            memcpy(&v->power, d + pos + 36, 4);
            memcpy(&v->ceremony_power, d + pos + 40, 4);
            have_hdr = true;
        } else if (id == 2) {
            v->tau_g1 = d + pos;
            v->n_tau_g1 = size / 64;
        } else if (id == 3) {
            v->tau_g2 = d + pos;
            v->n_tau_g2 = size / 128;
        } else if (id == 12) {
            v->prepared = true;
        }
        pos += size;
    }
    if (!have_hdr) return "ptau file: no header section";
    if (v->power > 28) return "ptau file: power out of range";
    if (!v->tau_g1 || v->n_tau_g1 < ((uint64_t)2 << v->power) - 1) return "ptau file: tauG1 section missing or short";
    if (!v->tau_g2 || v->n_tau_g2 < 2) return "ptau file: tauG2 section missing or short";
    return nullptr;
}
thread_local char g_ptau_err[128] = "";
}  // namespace

extern "C" int32_t nzcb_ptau_info(const uint8_t* ptau, size_t len, uint32_t* power, uint32_t* ceremony_power,
                                  uint64_t* n_tau_g1, int32_t* prepared) {
    PtauView v;
    const char* e = ptau_parse(ptau, len, &v);
    if (e) {
        snprintf(g_ptau_err, sizeof(g_ptau_err), "%s", e);
        return NZCB_E_INVALID;
    }
    if (power) *power = v.power;
    if (ceremony_power) *ceremony_power = v.ceremony_power;
    if (n_tau_g1) *n_tau_g1 = v.n_tau_g1;
    if (prepared) *prepared = v.prepared ? 1 : 0;
    return 0;
}

extern "C" int32_t nzcb_plonk_setup_ptau(nzcb_ctx* ctx, const uint8_t* r1cs, size_t r1cs_len, const uint8_t* ptau,
                                         size_t ptau_len, uint8_t* zkey_out, size_t* zkey_len) {
    if (!ctx || !r1cs || !ptau || !zkey_len) return NZCB_E_INVALID;
    PtauView v;
    const char* e = ptau_parse(ptau, ptau_len, &v);
    if (e) return ctx->fail(NZCB_E_INVALID, "%s", e);
    if (!v.prepared) return ctx->fail(NZCB_E_INVALID, "Powers of tau is not prepared.");
    uint32_t n_gates = 0, n_add = 0, n_vars = 0, cir_power = 0;
    NZ_TRY(nzcb_plonk_setup_info(ctx, r1cs, r1cs_len, &n_gates, &n_add, &n_vars, &cir_power));
    if (cir_power > v.power)
        return ctx->fail(NZCB_E_INVALID, "circuit too big for this power of tau ceremony. %u > 2**%u", n_gates, v.power);
    const size_t need = ((size_t)1 << cir_power) + 6;
    if (need > v.n_tau_g1) return ctx->fail(NZCB_E_INVALID, "ptau file: %zu tauG1 points needed, %llu present", need,
                                            (unsigned long long)v.n_tau_g1);
    return nzcb_plonk_setup(ctx, r1cs, r1cs_len, v.tau_g1, need, v.tau_g2 + 128, zkey_out, zkey_len);
}
