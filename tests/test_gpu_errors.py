"""-m gpu: error behaviour of the C ABI (negative return code + zkp_last_error message, surfaced as
ZkpB200Error) and of the table cache; nothing here may crash the process or fall back to the CPU."""
import pytest

from oracle import bn254

pytestmark = pytest.mark.gpu
R = bn254.R


def test_bad_handles_and_ranges(native):
    from interactive_zkp_study_b200.native import DeviceHandle, ZkpB200Error
    pts = [bn254.g1_mul(bn254.G1, k) for k in (2, 3, 5, 7)]
    table = native.g1_table_load(native.g1_vec_bytes(pts), 4)
    sc = native.scalars_load(native.fr_vec_bytes([1, 2, 3, 4]), 4)
    with pytest.raises(ZkpB200Error, match="range|exceeds"):
        native.g1_msm_table(table, 2, native.fr_vec_bytes([1, 2, 3]), 3)        # 2 + 3 > 4 points
    with pytest.raises(ZkpB200Error, match="bounds"):
        native.g1_msm_dev(table, 0, sc, 2, 3)                                     # 2 + 3 > 4 scalars
    bogus = DeviceHandle(987654321, 4, "g1")
    with pytest.raises(ZkpB200Error, match="handle"):
        native.g1_msm_dev(bogus, 0, sc, 0, 4)
    bogus.handle = 0
    with pytest.raises(ZkpB200Error, match="handle"):
        native.g1_msm_dev(sc, 0, sc, 0, 4)                                        # a scalar vector is not a table
    g2t = native.g2_table_load(native.g2_bytes(bn254.G2), 1)
    with pytest.raises(ZkpB200Error, match="handle"):
        native.g1_msm_dev(g2t, 0, sc, 0, 1)                                       # G2 table on the G1 entry point
    native.table_precompute(table, 6)
    with pytest.raises(ZkpB200Error, match="already"):
        native.table_precompute(table, 6)
    with pytest.raises(ZkpB200Error, match="window"):
        native.set_window_bits(33)
    # the library is still healthy afterwards
    assert native.g1_msm_dev(table, 0, sc, 0, 4) == bn254.g1_msm(pts, [1, 2, 3, 4])


def test_polynomial_argument_errors(native):
    from interactive_zkp_study_b200.native import ZkpB200Error
    enc = native.fr_vec_bytes
    with pytest.raises(ZkpB200Error, match="a_len >= b_len"):
        native.fr_poly_divmod(enc([1, 2]), 2, enc([1, 2, 3]), 3)
    with pytest.raises(ZkpB200Error, match="leading coefficient"):
        native.fr_poly_divmod(enc([1, 2, 3]), 3, enc([1, 0]), 2)
    with pytest.raises(ZkpB200Error, match="log_n"):
        native.fr_ntt(enc([0]), 29, 1)
    with pytest.raises(ZkpB200Error, match="empty|null"):
        native.fr_poly_mul(b"", 0, enc([1]), 1)


def test_table_cache_follows_list_contents(native):
    """tables.py caches by identity but must notice in-place edits (static tables are
    re-passed on every call)."""
    from interactive_zkp_study_b200 import tables
    from interactive_zkp_study_b200.compat import g1_from_ints
    from interactive_zkp_study_b200.zkp.plonk.kzg import commit
    from interactive_zkp_study_b200.zkp.plonk.polynomial import Polynomial
    from interactive_zkp_study_b200.zkp.plonk.srs import SRS
    pts = [g1_from_ints(bn254.g1_mul(bn254.G1, k)) for k in (1, 2, 3, 4)]
    srs = SRS(pts, [None, None], 3)
    p = Polynomial([5, 6, 7, 8])
    want = bn254.g1_msm([(int(q[0]), int(q[1])) for q in pts], [5, 6, 7, 8])
    got = commit(p, srs)
    assert (int(got[0]), int(got[1])) == want
    h1 = tables.g1_table(srs.g1_powers)
    assert tables.g1_table(srs.g1_powers) is h1                    # cache hit
    srs.g1_powers[3] = g1_from_ints(bn254.g1_mul(bn254.G1, 9))   # in-place edit of the SRS list
    want2 = bn254.g1_msm([(int(q[0]), int(q[1])) for q in srs.g1_powers], [5, 6, 7, 8])
    got2 = commit(p, srs)
    assert (int(got2[0]), int(got2[1])) == want2 != want
    # a longer table edited at an index no sampling scheme would look at, then a coordinate mutated in place
    pts = [g1_from_ints(bn254.g1_add(bn254.g1_mul(bn254.G1, 77), bn254.g1_mul(bn254.G1, k))) for k in range(1, 41)]
    srs = SRS(pts, [None, None], 39)
    coeffs = list(range(3, 43))
    ref = lambda: bn254.g1_msm([(int(q[0]), int(q[1])) for q in srs.g1_powers], coeffs)
    big = Polynomial(coeffs)
    first = commit(big, srs)
    assert (int(first[0]), int(first[1])) == ref()
    srs.g1_powers[5] = g1_from_ints(bn254.g1_mul(bn254.G1, 123456))
    second = commit(big, srs)
    assert (int(second[0]), int(second[1])) == ref() != (int(first[0]), int(first[1]))
    other = bn254.g1_mul(bn254.G1, 424242)
    srs.g1_powers[17][0].n, srs.g1_powers[17][1].n = other       # same objects, new coordinates
    third = commit(big, srs)
    assert (int(third[0]), int(third[1])) == ref() != (int(second[0]), int(second[1]))
    # a same-length list that reuses the id() of a dead one cannot hit: entries hold their list alive
    assert all(e.points is not None for e in tables._cache.values())
