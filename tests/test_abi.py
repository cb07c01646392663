"""The drop-in boundary without a GPU: libnzcb.so loads, exports every function include/nzcb.h declares, the ctypes
binding covers them all, and the entry points fail loudly (no CPU fallback) when no sm_100 device is usable."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "nzcb.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nzcb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from nzcb_circom_b200 import _lib

    lib = _lib.load()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/nzcb.h but not exported by libnzcb.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in nzcb_circom_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and header disagree"


def test_no_cpu_fallback_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from nzcb_circom_b200 import Context, NzcbError

    with pytest.raises(NzcbError) as e:
        Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under nzcb_circom_b200/ may import it"""
    pkg = os.path.join(ROOT, "nzcb_circom_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                with open(os.path.join(d, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_proof_to_json_needs_no_gpu():
    from nzcb_circom_b200 import _lib

    lib = _lib.load()
    p = _lib.Proof()
    n = ctypes.c_size_t(0)
    assert lib.nzcb_proof_to_json(ctypes.byref(p), None, ctypes.byref(n)) == 0 and n.value > 100
    buf = ctypes.create_string_buffer(n.value)
    assert lib.nzcb_proof_to_json(ctypes.byref(p), buf, ctypes.byref(n)) == 0
    assert b'"protocol": "plonk"' in buf.value and b'"curve": "bn128"' in buf.value


def test_napi_shim_type_checks_against_the_header():
    """integration/nzcb_napi.c (the binding a maintainer of the reference adds; Node is absent here) compiles against
    include/nzcb.h and a stub of the N-API declarations it uses, and calls only functions the header declares"""
    import shutil
    import subprocess
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        shutil.copy(os.path.join(ROOT, "tests", "hostcheck", "node_api_stub.h"), os.path.join(d, "node_api.h"))
        subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", d, "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "integration", "nzcb_napi.c")], check=True)
    with open(os.path.join(ROOT, "integration", "nzcb_napi.c")) as f:
        used = set(re.findall(r"\b(nzcb_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)))
    assert used and used <= set(_declared()) | {"nzcb_napi_register_stub"}, used - set(_declared())
