"""Port of /root/reference/test/quinSelector.js: select every index of QuinSelector(0..5),
out-of-range index is rejected (:63-87); plus log2 / pow compile-time suites
(circuits/log2_test.circom:6-22, circuits/pow_test.circom:5-29; test/log2.js, test/pow.js)."""
import pytest

from nzcb_circom_b200.circom.circomlib import log2, pow_
from tests.wbackend import BACKENDS, calc


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
def test_quin_selector(backend, n):
    arr = list(range(1, n + 1))
    for index in range(n):
        w = calc(f"quinSelector{n}_test", {"in": arr, "index": index}, backend, check_r1cs=True)
        assert w[1] == arr[index]
    assert calc(f"quinSelector{n}_test", {"in": arr, "index": n}, backend) is None          # :63-87
    assert calc(f"quinSelector{n}_test", {"in": arr, "index": n + 1000}, backend) is None
    # a slightly negative index passes LessThan by wrap-around and selects nothing -- the behaviour
    # ConstructNullifier relies on (nzcptpl.circom:413,417)
    assert calc(f"quinSelector{n}_test", {"in": arr, "index": -1}, backend)[1] == 0


@pytest.mark.parametrize("backend", BACKENDS)
def test_quin_selector_0(backend):  # quinSelector.circom:41 -- no choices: out = 0, no bounds check
    assert calc("quinSelector0_test", {"index": 0}, backend)[1] == 0
    assert calc("quinSelector0_test", {"index": 7}, backend)[1] == 0


def test_log2():
    for i in range(1, 100):
        assert log2(2 ** i) == i and log2(2 ** i + 1) == i
    assert [log2(x) + 1 for x in range(1, 9)] == [1, 2, 2, 3, 3, 3, 3, 4]
    assert log2(0) == -1


def test_pow():
    assert [pow_(1, k) for k in range(3)] == [1, 1, 1]
    assert [pow_(2, k) for k in range(9)] == [1, 2, 4, 8, 16, 32, 64, 128, 256]
    assert [pow_(3, k) for k in range(9)] == [1, 3, 9, 27, 81, 243, 729, 2187, 6561]
