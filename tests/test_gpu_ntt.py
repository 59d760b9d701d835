"""-m gpu: Fr NTT / iNTT / coset variants through the C ABI against the oracle's restatement of the
reference's recursive fft (/root/reference/zkp/plonk/polynomial.py:292-378,
/root/reference/zkp/plonk/utils.py:145-205): bit-exact at sizes the oracle finishes in seconds
(n = 1 .. 2^12, the reference tests use n = 1, 4, 8), Horner spot checks + round trips at 2^20."""
import random

import pytest

from oracle import ref_path

pytestmark = pytest.mark.gpu
R = ref_path.R


def _run(native, vals, omega, inverse=False, shift=None):
    n = len(vals)
    out = native.fr_ntt(native.fr_vec_bytes(vals), n.bit_length() - 1, omega, inverse, shift)
    return native.fr_vec_from_bytes(out)


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 9, 10, 11, 12])
def test_fft_matches_reference_recursion(native, log_n):
    n = 1 << log_n
    rng = random.Random(log_n)
    vals = [rng.randrange(R) for _ in range(n)]
    w = ref_path.get_root_of_unity(n)
    assert _run(native, vals, w) == ref_path.fft(vals, w)
    assert _run(native, vals, w, inverse=True) == ref_path.ifft(vals, w)


@pytest.mark.parametrize("log_n", [2, 3, 11])
def test_coset_fft_matches_reference(native, log_n):
    n = 1 << log_n
    rng = random.Random(50 + log_n)
    vals = [rng.randrange(R) for _ in range(n)]
    w = ref_path.get_root_of_unity(n)
    assert _run(native, vals, w, shift=5) == ref_path.coset_fft(vals, w, 5)
    assert _run(native, vals, w, inverse=True, shift=5) == ref_path.coset_ifft(vals, w, 5)
    assert _run(native, vals, w, shift=7) == ref_path.coset_fft(vals, w, 7)


def test_fft_with_inverse_root_and_edge_values(native):
    n = 8
    w = ref_path.get_root_of_unity(n)
    vals = [0, 1, R - 1, 0, 5, R - 2, 0, 0]
    assert _run(native, vals, ref_path.inv(w)) == ref_path.fft(vals, ref_path.inv(w))
    assert _run(native, [0] * n, w) == [0] * n
    assert _run(native, [1] + [0] * (n - 1), w) == [1] * n


@pytest.mark.parametrize("log_n", [14, 17, 20])
def test_ntt_large_horner_and_round_trip(native, log_n):
    n = 1 << log_n
    rng = random.Random(900 + log_n)
    vals = [rng.randrange(R) for _ in range(n)]
    w = ref_path.get_root_of_unity(n)
    data = native.fr_vec_bytes(vals)
    ev = native.fr_ntt(data, log_n, w)
    for k in [0, 1, 2, n // 2 - 1, n // 2, n - 1] + [rng.randrange(n) for _ in range(3)]:
        got = int.from_bytes(ev[32 * k:32 * k + 32], "little")
        assert got == ref_path.poly_eval(vals, pow(w, k, R))
    assert native.fr_ntt(ev, log_n, w, inverse=True) == data
    cev = native.fr_ntt(data, log_n, w, False, 5)
    k = rng.randrange(n)
    assert int.from_bytes(cev[32 * k:32 * k + 32], "little") == ref_path.poly_eval(vals, 5 * pow(w, k, R) % R)
    assert native.fr_ntt(cev, log_n, w, True, 5) == data


@pytest.mark.parametrize("log_n", [13, 15, 16, 18, 19, 21, 22, 23, 24])
def test_ntt_every_pass_shape_device_resident(native, log_n):
    """Every decomposition of log_n into shared-memory passes (10 + up to two passes of <= 7 stages),
    on device-resident data: NTT output element k == p(w^k) by the (oracle-checked) device Horner, the
    coset variant == p(5 w^k), and inverse(forward(x)) == x bit for bit."""
    n = 1 << log_n
    w = ref_path.get_root_of_unity(n)
    h = native.scalars_generate(0x5EED0004 + log_n, n)
    orig = native.scalars_alloc(n)
    native.scalars_copy(orig, 0, h, 0, n)
    native.ntt_dev(h, 0, log_n, w)
    rng = random.Random(log_n)
    for k in [0, 1, n // 2, n - 1, rng.randrange(n), rng.randrange(n)]:
        got = int.from_bytes(native.scalars_download(h, k, 1), "little")
        assert got == native.fr_poly_eval_dev(orig, 0, n, pow(w, k, R)), (log_n, k)
    native.ntt_dev(h, 0, log_n, w, inverse=True)
    diff = native.scalars_alloc(n)
    native.vec_op_dev(1, diff, 0, h, 0, orig, 0, n)
    assert native.scalars_is_zero(diff, 0, n)
    native.ntt_dev(h, 0, log_n, w, coset_shift=5)
    k = rng.randrange(n)
    assert int.from_bytes(native.scalars_download(h, k, 1), "little") == \
        native.fr_poly_eval_dev(orig, 0, n, 5 * pow(w, k, R) % R)
    native.ntt_dev(h, 0, log_n, w, inverse=True, coset_shift=5)
    native.vec_op_dev(1, diff, 0, h, 0, orig, 0, n)
    assert native.scalars_is_zero(diff, 0, n)
