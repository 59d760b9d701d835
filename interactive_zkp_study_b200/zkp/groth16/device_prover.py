"""Groth16 proving at scale: coefficient-vector entry points over device-resident keys.

The reference's `proof_a/b/c(sigma..., Ax, Rx, ...)` signatures take the dense numWires x numGates QAP
matrices, which cannot exist at 2^20 constraints (2^40 entries).  What the prover needs
from them is only the three coefficient vectors uA = R.Ax, uB = R.Bx, uC = R.Cx (numGates entries
each) -- this module is the same algebra as zkp/groth16/proving.py + poly_utils.hxr
(/root/reference/zkp/groth16/proving.py:23-75, poly_utils.py:116-125) on those vectors, with the CRS
resident on the GPU in three concatenated, window-precomputed tables so that each proof element is
ONE multi-scalar multiplication:

    TA  (G1) = [x^j]_1 (j < k) | alpha_1 | delta_1                  scalars  uA | 1 | r            -> A
    TB2 (G2) = [x^j]_2 (j < k) | beta_2  | delta_2                  scalars  uB | 1 | s            -> B
    TC  (G1) = [x^j]_1 (j < k) | beta_1 | sigma1_4[private] | sigma1_5 (k-1) | alpha_1 | delta_1
                                                       scalars  r*uB + s*uA | r | Rx_priv | H | s | s*r -> C

C = s*A + r*B1 - s*r*delta_1 + (wires) + (H) of proving.py:64-73 with A and B1 expanded over the SAME points [x^j]_1:
the s*A term becomes s*uA on the first k rows plus s*alpha_1 + s*r*delta_1, the -s*r*delta_1 cancels against
r*B1's s*r*delta_1, so C needs neither A's value nor a scalar multiplication -- the three elements are three
independent multi-scalar multiplications, and B (G2, the longest) starts before the quotient it does not need.
"""
from ... import native
from ...compat import G1, G2, curve_order, g1_from_ints, g2_from_ints

R = curve_order


class DeviceKey:
    def __init__(self, k, m_priv, TA, TB2, TC):
        self.k, self.m_priv, self.TA, self.TB2, self.TC = k, m_priv, TA, TB2, TC


def tc_rows(k, m_priv):
    """Rows of the C table: [x^j]_1 (k) | beta_1 | sigma1_4[private] (m_priv) | sigma1_5 (k-1) | alpha_1 | delta_1."""
    return k + 1 + m_priv + (k - 1) + 2


def _powers(x, count):
    return native.fr_prefix_product(native.fr_vec_bytes([x % R] * count), count)


def setup_from_toxic(k, alpha, beta, delta, x_val, zx_val, priv_vals, precompute=True):
    """CRS generation on the device (batched fixed-base multiplications; reference setup.py:15-69).
    priv_vals[i] = (beta*A_i(x) + alpha*B_i(x) + C_i(x)) / delta for the private wires, in order."""
    m_priv = len(priv_vals)
    pw = _powers(x_val, k)
    inv_delta = pow(delta % R, -1, R)
    s15 = native.fr_vec_op(3, pw[:32 * (k - 1)], native.fe_bytes(zx_val % R * inv_delta % R), k - 1) if k > 1 else b""
    enc = native.fr_vec_bytes
    scA = pw + enc([alpha % R, delta % R])
    scC = pw + enc([beta % R]) + enc([v % R for v in priv_vals]) + s15 + enc([alpha % R, delta % R])
    TA = native.g1_fixed_base_mul(native.g1_bytes(G1), scA, k + 2)
    TB2 = native.g2_fixed_base_mul(native.g2_bytes(G2), pw + enc([beta % R, delta % R]), k + 2)
    TC = native.g1_fixed_base_mul(native.g1_bytes(G1), scC, tc_rows(k, m_priv))
    if precompute:
        for t in (TA, TB2, TC):
            if t.n >= 2:
                native.table_precompute(t)
    return DeviceKey(k, m_priv, TA, TB2, TC)


def _scalars_ab(key, uA, uB, r, s):
    """Scalar vectors of the A and B elements (they need only uA, uB): uA | 1 | r and uB | 1 | s."""
    k = key.k
    one = native.fr_vec_bytes([1])
    scA = native.scalars_alloc(k + 2)
    native.scalars_copy(scA, 0, uA, 0, k)
    native.scalars_upload(scA, k, one + native.fe_bytes(r), 2)
    scB = native.scalars_alloc(k + 2)
    native.scalars_copy(scB, 0, uB, 0, k)
    native.scalars_upload(scB, k, one + native.fe_bytes(s), 2)
    return scA, scB


def _scalars_c(key, uA, uB, hq, rx_priv, r, s):
    """Scalar vector of C: r*uB + s*uA | r | Rx_priv | H | s | s*r  (see the table layout above)."""
    k, mp = key.k, key.m_priv
    nC = tc_rows(k, mp)
    scC = native.scalars_alloc(nC)
    native.scalars_copy(scC, 0, uB, 0, k)
    native.scalars_scale(scC, 0, k, r)
    native.axpy_dev(scC, 0, s, uA, 0, k)
    native.scalars_upload(scC, k, native.fe_bytes(r), 1)
    if mp:
        native.scalars_copy(scC, k + 1, rx_priv, 0, mp)
    if k > 1:
        native.scalars_copy(scC, k + 1 + mp, hq, 0, k - 1)
    native.scalars_upload(scC, nC - 2, native.fe_bytes(s) + native.fe_bytes(s * r % R), 2)
    return scC, nC


def prove(key, uA, uB, uC, Z, rx_priv, r, s, keep_quotient=False):
    """uA, uB, uC: device scalar handles with k coefficients; Z: handle with k+1 coefficients (monic
    divisor); rx_priv: handle with the private wires' witness values.  Returns (A, B, C) as the
    reference's point types (and the quotient / remainder handles when keep_quotient)."""
    k = key.k
    r, s = int(r) % R, int(s) % R
    # H: k-1 coefficients; the remainder (k coefficients, zero for a satisfied instance) only on request
    # B (G2, the longest of the three) needs only uB: it starts on the library's second stream and runs beside the
    # quotient; A and C follow on the first stream
    scA, scB = _scalars_ab(key, uA, uB, r, s)
    native.msm_dev_begin(key.TB2, 0, scB, 0, k + 2)
    hq, hr = native.groth16_quotient_dev(uA, uB, uC, k, Z, k + 1, want_remainder=keep_quotient)
    scC, nC = _scalars_c(key, uA, uB, hq, rx_priv, r, s)
    A = native.g1_msm_dev(key.TA, 0, scA, 0, k + 2)
    C = native.g1_msm_dev(key.TC, 0, scC, 0, nC)
    B = native.msm_dev_end("g2")
    for h in (scA, scB, scC):
        h.free()
    out = (g1_from_ints(A), g2_from_ints(B), g1_from_ints(C))
    if keep_quotient:
        return out + (hq, hr)
    hq.free()
    return out


# ------------------------------------------------------------------ the same proof over the GPUs of one box
class ShardedDeviceKey:
    """This rank's contiguous row range of each of the three CRS tables (sharded.shard_range of the
    concatenated table), resident and window-precomputed on this rank's GPU."""

    def __init__(self, k, m_priv, TA, TB2, TC, rA, rB, rC):
        self.k, self.m_priv, self.TA, self.TB2, self.TC = k, m_priv, TA, TB2, TC
        self.rA, self.rB, self.rC = rA, rB, rC        # (start, count) of this rank in each table


def setup_from_toxic_sharded(comm, k, alpha, beta, delta, x_val, zx_val, priv_vals, precompute=True):
    """setup_from_toxic with every table cut by point range over the ranks of `comm`: a rank
    generates and keeps only its rows.  Same arguments on every rank."""
    from ... import sharded
    m_priv = len(priv_vals)
    pw = _powers(x_val, k)
    inv_delta = pow(delta % R, -1, R)
    s15 = native.fr_vec_op(3, pw[:32 * (k - 1)], native.fe_bytes(zx_val % R * inv_delta % R), k - 1) if k > 1 else b""
    enc = native.fr_vec_bytes
    scA = pw + enc([alpha % R, delta % R])
    scB = pw + enc([beta % R, delta % R])
    scC = pw + enc([beta % R]) + enc([v % R for v in priv_vals]) + s15 + enc([alpha % R, delta % R])
    out = []
    for sc, g2 in ((scA, False), (scB, True), (scC, False)):
        start, count = sharded.shard_range(len(sc) // 32, comm.rank, comm.world)
        part = sc[32 * start:32 * (start + count)]
        t = (native.g2_fixed_base_mul(native.g2_bytes(G2), part, count) if g2
             else native.g1_fixed_base_mul(native.g1_bytes(G1), part, count))
        if precompute and count >= 2:
            native.table_precompute(t)
        out.append((t, (start, count)))
    (TA, rA), (TB2, rB), (TC, rC) = out
    return ShardedDeviceKey(k, m_priv, TA, TB2, TC, rA, rB, rC)


def prove_sharded(comm, key, uA, uB, uC, Z, rx_priv, r, s):
    """prove() with each of the three multi-scalar multiplications sharded by point range over the ranks
    (zkp_g1_msm_multi / zkp_g2_msm_multi: local Pippenger -> one all-gather of the partial sums -> fold, inside
    the library).  Collective: every rank passes the same coefficient vectors and gets the same proof.  The
    quotient -- NTT work, which stays on one GPU -- is computed by every rank for itself: that
    costs no wall time and needs no 32 MiB broadcast of H; the A and B collectives run beside it on the second
    stream (zkp_g1/g2_msm_multi_begin / _end)."""
    k = key.k
    r, s = int(r) % R, int(s) % R
    # A and B need only uA, uB: their collectives start on the library's second stream and run beside the quotient
    scA, scB = _scalars_ab(key, uA, uB, r, s)
    native.msm_multi_begin(key.TA, 0, scA, key.rA[0], key.rA[1])
    native.msm_multi_begin(key.TB2, 0, scB, key.rB[0], key.rB[1])
    hq, _ = native.groth16_quotient_dev(uA, uB, uC, k, Z, k + 1, want_remainder=False)
    scC, _ = _scalars_c(key, uA, uB, hq, rx_priv, r, s)
    C = native.g1_msm_multi(key.TC, 0, scC, key.rC[0], key.rC[1])
    A = native.msm_multi_end("g1")
    B = native.msm_multi_end("g2")
    for h in (scA, scB, scC, hq):
        h.free()
    return (g1_from_ints(A), g2_from_ints(B), g1_from_ints(C))
