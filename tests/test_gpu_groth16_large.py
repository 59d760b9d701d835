"""-m gpu: the coefficient-level Groth16 prover on device-resident keys (BASELINE config 3 shape):
CRS with known toxic waste generated on the device, so every proof element has a closed-form
discrete log and is checked bit-for-bit against one oracle scalar multiplication
(zkp/groth16/test.py:331 checks proof_A the same way); the quotient is checked against the oracle's
long division (small k) and by the polynomial identity at a random point (large k)."""
import random

import pytest

from oracle import bn254, ref_path

pytestmark = pytest.mark.gpu
R = bn254.R


def _case(native, k, mp, seed, precompute):
    from interactive_zkp_study_b200.zkp.groth16 import device_prover as dp
    rng = random.Random(seed)
    alpha, beta, delta, x = (rng.randrange(1, R) for _ in range(4))
    z_host = [rng.randrange(R) for _ in range(k)] + [1]           # monic degree-k divisor
    zx = ref_path.poly_eval(z_host, x)
    priv = [rng.randrange(R) for _ in range(mp)]
    key = dp.setup_from_toxic(k, alpha, beta, delta, x, zx, priv, precompute=precompute)
    uA, uB, uC = (native.scalars_generate(0x5EED0100 + i, k) for i in range(3))
    Z = native.scalars_load(native.fr_vec_bytes(z_host), k + 1)
    rx = native.scalars_generate(0x5EED0200, max(mp, 1))
    r, s = rng.randrange(R), rng.randrange(R)
    A, B, C, hq, hr = dp.prove(key, uA, uB, uC, Z, rx, r, s, keep_quotient=True)
    ev = lambda h, n, at: native.fr_poly_eval_dev(h, 0, n, at)
    a = (alpha + ev(uA, k, x) + r * delta) % R
    b = (beta + ev(uB, k, x) + s * delta) % R
    rx_host = native.fr_vec_from_bytes(native.scalars_download(rx, 0, mp)) if mp else []
    wires = sum(p * q for p, q in zip(rx_host, priv)) % R
    hx = ev(hq, k - 1, x) if k > 1 else 0
    c = (s * a + r * (beta + ev(uB, k, x)) + wires + hx * zx % R * pow(delta, -1, R)) % R
    assert (int(A[0]), int(A[1])) == bn254.g1_mul(bn254.G1, a)
    got_b = ((int(B[0].coeffs[0]), int(B[0].coeffs[1])), (int(B[1].coeffs[0]), int(B[1].coeffs[1])))
    assert got_b == bn254.g2_mul(bn254.G2, b)
    assert (int(C[0]), int(C[1])) == bn254.g1_mul(bn254.G1, c)
    return uA, uB, uC, z_host, hq, hr


@pytest.mark.parametrize("k,mp,precompute", [(8, 5, False), (257, 100, False), (4096, 4000, True)])
def test_device_prover_small(native, k, mp, precompute):
    uA, uB, uC, z_host, hq, hr = _case(native, k, mp, 1000 + k, precompute)
    dl = lambda h, n: native.fr_vec_from_bytes(native.scalars_download(h, 0, n))
    a, b, c = dl(uA, k), dl(uB, k), dl(uC, k)
    if k <= 300:
        P = ref_path.g16_subtract_polys(ref_path.g16_multiply_polys(a, b), c)
        H, rem = ref_path.g16_div_polys(P, z_host)
        assert dl(hq, k - 1) == H and dl(hr, k) == rem
    t = 0x1234567890abcdef
    lhs = (ref_path.poly_eval(a, t) * ref_path.poly_eval(b, t) - ref_path.poly_eval(c, t)) % R
    rhs = (ref_path.poly_eval(dl(hq, k - 1), t) * ref_path.poly_eval(z_host, t) + ref_path.poly_eval(dl(hr, k), t)) % R
    assert lhs == rhs


def test_device_prover_2_16(native):
    """Same checks at 2^16 constraints with window-precomputed tables (the 2^20 run is in bench.py)."""
    k = 1 << 16
    uA, uB, uC, z_host, hq, hr = _case(native, k, k - 2, 77, True)
    t = 0xfeedface12345
    ev = lambda h, n: native.fr_poly_eval_dev(h, 0, n, t)
    lhs = (ev(uA, k) * ev(uB, k) - ev(uC, k)) % R
    rhs = (ev(hq, k - 1) * ref_path.poly_eval(z_host, t) + ev(hr, k)) % R
    assert lhs == rhs


def test_device_prover_2_20_config3(native):
    """BASELINE config 3 inside the GPU suite: Groth16 prove at 2^20 constraints, every proof element equal to its
    closed-form discrete log times the generator and the quotient identity at a random point (the same
    checks bench.py's groth16_prove sub-record makes)."""
    k = 1 << 20
    uA, uB, uC, z_host, hq, hr = _case(native, k, k - 2, 2026, True)
    t = 0xfeedface12345
    ev = lambda h, n: native.fr_poly_eval_dev(h, 0, n, t)
    zh = native.scalars_load(native.fr_vec_bytes(z_host), k + 1)
    lhs = (ev(uA, k) * ev(uB, k) - ev(uC, k)) % R
    rhs = (ev(hq, k - 1) * ev(zh, k + 1) + ev(hr, k)) % R
    assert lhs == rhs
