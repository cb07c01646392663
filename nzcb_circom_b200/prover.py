"""One object for the whole hot path of one circuit on one GPU: what a long-lived prover
process holds -- the compiled circuit's witness program, the SRS-derived zkey (device
resident, Montgomery form) and the fused witness+prove entry point.  Multi-GPU = one such
object per process / GPU, passes sharded by the caller, no collective."""
import time

from . import nzcp_helpers as H
from ._lib import default_context
from .circom_tester import wasm_tester
from .snarkjs import ZKey, plonk, powersoftau, zKey

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


def default_tau():
    """tau = Fr(keccak256("nzcb-b200-tau")): the known trapdoor of the synthetic SRS"""
    return 0xae4711c826850d09ad8857707a9efce27474fb4937e510dc529a1baf89b6f59


class CircuitProver:
    def __init__(self, circuit="nzcp_live", tau=None, ctx=None, verbose=False):
        self.ctx = ctx or default_context()
        self.timings = {}
        t = time.time()
        self.tester = wasm_tester(circuit, self.ctx)
        self.art = self.tester.compiled
        self.timings["circuit"] = time.time() - t
        self.tau = tau
        self.zk = None
        self.vk = None
        self.verbose = verbose

    def setup(self, srs_g1_lem=None, keep_zkey=False, x2_g2_lem=None):
        """powersoftau + plonk setup on the GPU, zkey made device resident; returns the vk object
        (or the zkey file bytes with keep_zkey=True, for callers that also want to write it out)"""
        from .snarkjs import R_MOD as _R
        art = self.art
        t = time.time()
        # the domain is only known after the R1CS -> PLONK expansion; ask the library for the zkey size
        # with a generous SRS first
        r1cs = art.r1cs_bytes()
        if srs_g1_lem is None:
            if self.tau is None:
                raise ValueError("need a trapdoor tau or an SRS")
            # nGates <= nConstraints + total LC terms; 2^21 + 6 covers nzcp_live (README.md:41)
            power = self._domain_power(r1cs)
            srs_g1_lem = powersoftau.new_g1(self.tau % _R, (1 << power) + 6, self.ctx)
        self.timings["srs"] = time.time() - t
        t = time.time()
        # X_2 = [tau]_2 when the trapdoor is known (the synthetic SRS); an external SRS must bring its own
        if x2_g2_lem is None:
            if self.tau is None:
                raise ValueError("an external SRS must come with its X_2 = [tau]_2 (x2_g2_lem)")
            x2_g2_lem = powersoftau.new_g2(self.tau % _R, self.ctx)
        x2 = x2_g2_lem
        zkey = plonk.setup(r1cs, srs_g1_lem, x2, self.ctx)
        self.timings["setup"] = time.time() - t
        t = time.time()
        self.vk = zKey.exportVerificationKey(zkey)
        self.zk = ZKey(zkey, self.ctx)
        self.timings["zkey_load"] = time.time() - t
        self.zkey_bytes_len = len(zkey)
        if keep_zkey:
            return zkey
        del zkey
        return self.vk

    def _domain_power(self, r1cs):
        """log2 of the PLONK domain (the R1CS -> PLONK gate expansion runs in the library)"""
        ng, na, nv, power = plonk.setup_info(r1cs, self.ctx)
        self.n_gates, self.n_additions, self.plonk_vars, self.power = ng, na, nv, power
        return power

    # ---- proving -----------------------------------------------------
    def prove_inputs(self, inputs, blinders_list=None):
        """inputs: list of circuit input dicts -> [(proof bytes | None, publicSignals, status)]"""
        return plonk.fullProveBatch(inputs, self.tester, self.zk, blinders_list, self.ctx)

    def prove_raw(self, inputs_le, B, blinders_list=None, device_inputs=None):
        return plonk.fullProveRaw(inputs_le, B, self.tester._handle(self.ctx), self.zk, blinders_list, self.ctx,
                                  device_inputs)

    def verify(self, publics_list, proofs):
        """plonk.verify of B proofs against this circuit's key, on the GPU -> [bool]"""
        from .snarkjs import VKey

        if getattr(self, "_vkey", None) is None:
            self._vkey = VKey(self.vk, self.ctx)
        return plonk.verify_batch(self._vkey, publics_list, proofs, self.ctx)

    def marshal(self, inputs):
        """host-side marshalling of input dicts to the B x nInputs x 32 B buffer the C ABI takes"""
        art = self.art
        out = bytearray()
        for inp in inputs:
            vals = art.flatten_input(inp)
            out += b"".join(int(v).to_bytes(32, "little") for v in vals)
        return bytes(out)


class NzcpProver(CircuitProver):
    """nzcp_live / nzcp_example: ToBeSigned bytes + 20 pass-through bytes in, proofs out"""

    def __init__(self, live=True, tau=None, ctx=None):
        super().__init__("nzcp_live" if live else "nzcp_example", tau, ctx)
        self.max_len = H.LIVE_TOBESIGNED_MAX if live else H.EXAMPLE_TOBESIGNED_MAX

    def marshal_passes(self, passes):
        """passes: list of (toBeSigned bytes, data 20 bytes) -> inputs buffer.  Bits become 32-byte
        field elements exactly as circom_runtime would receive them (test/nzcp.js:36-41)."""
        one = (1).to_bytes(32, "little")
        zero = bytes(32)
        out = bytearray()
        for tbs, data in passes:
            if len(tbs) > self.max_len:
                raise ValueError("ToBeSigned longer than the circuit supports")
            bits = H.bufferToBitArray(H.fitBytes(tbs, self.max_len))
            out += b"".join(one if b else zero for b in bits)
            out += len(tbs).to_bytes(32, "little")
            out += b"".join(one if b else zero for b in H.bufferToBitArray(H.evmRearrangeBytes(data)))
        return bytes(out)

    def prove_passes(self, passes, blinders_list=None):
        return self.prove_raw(self.marshal_passes(passes), len(passes), blinders_list)

    def prove_uris(self, passURIs, datas=None, blinders_list=None):
        """pass URIs ("NZCP:/1/...") -> proofs, ingest on the device (nzcb_plonk_fullprove_uri_batch)"""
        from .pass_ingest import fullProveURIs

        return fullProveURIs(passURIs, self.max_len, self.tester, self.zk, datas, blinders_list, self.ctx)
