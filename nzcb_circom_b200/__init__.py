"""nzcb_circom_b200 -- B200-native PLONK prover / batched witness generator for
the noway/nzcb-circom circuits.  The product is libnzcb.so (CUDA, sm_100a)
behind the C ABI in include/nzcb.h; this package is the host-side mirror of
the JS interfaces the reference drives it through (snarkjs, ffjavascript,
circom_tester)."""
from ._lib import Context, NzcbError, default_context, load  # noqa: F401
