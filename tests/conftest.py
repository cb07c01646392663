import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One nzcb context on cuda:0 for the whole GPU test session."""
    from nzcb_circom_b200 import Context

    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def nzcp_live_prover(ctx):
    """nzcp_live at full size (domain 2^21): circuit, synthetic SRS, GPU plonk setup, resident zkey."""
    from nzcb_circom_b200.prover import NzcpProver, default_tau

    pr = NzcpProver(live=True, tau=default_tau(), ctx=ctx)
    pr.zkey_bytes = pr.setup(keep_zkey=True)  # kept for the full-size byte comparison against the C oracle
    return pr
