"""PLONK proving at scale: the five rounds of /root/reference/zkp/plonk/prover/round1..5.py on
device-resident vectors (BASELINE config 4).

The list-of-FR `Polynomial` objects of the reference cost O(n) Python conversions per operation and its
round 3 is O(n^2); at 2^20 gates the rounds have to stay in HBM.  Same protocol, same transcript, same
proof elements (bit-identical to the reference-minted golden proofs for equal blinding scalars):

  round 1  3 iNTT, blinding, 3 MSMs                          (round1.py:38-108)
  round 2  grand product = fused num/den kernel + batch inversion + product scan, iNTT, MSM (on the second stream,
           beside the coset transforms of round 3, which do not need alpha)                  (round2.py, permutation.py:89-137)
  round 3  quotient on the coset {g w_N^i}, N = 4n (the smallest power of two >= 3n+6): 4 coset NTTs of the witness polynomials (the 8 circuit
           polynomials' coset evaluations are part of the device key), one pointwise kernel
           t = [gate + alpha perm]/Z_H + alpha^2 (z-1)/(n(x-1)), one inverse coset NTT, divisibility =
           "coefficients above 3n+5 vanish", 3 MSMs                                   (round3.py:56-184)
  round 4  6 two-level Horner evaluations                     (round4.py:39-81)
  round 5  linearisation r(x) by axpy, the two openings as (P(x)-P(z))/(x-z) = powers + sum scan, 2 MSMs (round5.py:42-175)
"""
import secrets

from ... import native
from ...compat import curve_order, g1_from_ints
from .field import FR, get_root_of_unity
from .transcript import Transcript
from .prover import Proof

R = curve_order
K1, K2 = 2, 3
COSET_SHIFT = 5          # 5^(8n) != 1 (5 is a non-residue), so Z_H never vanishes on g*<w8>
CIRCUIT_POLYS = ("q_l", "q_r", "q_o", "q_m", "q_c", "s_sigma1", "s_sigma2", "s_sigma3")


class DeviceKey:
    """Preprocessed circuit on the device: coefficient forms (canonical), their 8n-coset evaluations
    (Montgomery), sigma evaluations on H, the static coset data and the SRS table."""

    def __init__(self):
        self.coeffs, self.coset, self.comm = {}, {}, {}
        self.ranks, self.srs_range = None, None     # sharded SRS: the communicator and this rank's row range


def _alloc_from(src, n, total):
    h = native.scalars_alloc(total)
    native.scalars_copy(h, 0, src, 0, n)
    return h


def _commit(key, h, off, length):
    """kzg.commit (kzg.py:32-67) of coefficients h[off : off + length] against SRS powers 0 .. length - 1."""
    if key.ranks is None:
        return native.g1_msm_dev(key.srs_table, 0, h, off, length)
    # sharded SRS: this rank holds powers [start, start + count); its share of the commitment is the
    # MSM over the powers of its range that the polynomial reaches (possibly none), folded inside the library
    start, count = key.srs_range
    m = max(0, min(count, length - start))
    return native.g1_msm_multi(key.srs_table, 0, h, off + start if m else 0, m)


def _commit_begin(key, h, off, length):
    """_commit in split form: the MSM (or this rank's share and the collective) is enqueued on the library's second
    stream; _commit_end fetches the point.  What is called in between runs beside it."""
    if key.ranks is None:
        native.msm_dev_begin(key.srs_table, 0, h, off, length)
    else:
        start, count = key.srs_range
        m = max(0, min(count, length - start))
        native.msm_multi_begin(key.srs_table, 0, h, off + start if m else 0, m)


def _commit_end(key):
    return native.msm_dev_end("g1") if key.ranks is None else native.msm_multi_end("g1")


def _commit_many(key, items):
    """Several commitments against the SRS: items = [(handle, offset, length)].  On one GPU they go out as one
    pipelined call (the tail of one MSM overlaps the accumulation of the next); sharded, each is a collective."""
    if key.ranks is not None:
        return [_commit(key, h, off, length) for h, off, length in items]
    return native.g1_msm_dev_batch(key.srs_table, [(h, off, 0, length) for h, off, length in items])


def preprocess(n, selector_evals, sigma_evals, srs_table, srs_size, comm=None, srs_range=None):
    """selector_evals: 5 handles (q_l, q_r, q_o, q_m, q_c evaluations on H, n each); sigma_evals: 3
    handles with the evaluations of S_sigma1..3 (permutation.py:44-86).  Mirrors
    preprocessor.py:59-130: 8 iNTTs + 8 commitments, plus the static round-3 data.
    comm / srs_range: a sharded.Communicator and the (start, count) row range of the SRS that `srs_table`
    holds on this rank; every commitment of preprocess() and prove() is then a collective over the ranks,
    while the NTT / quotient work is done by every rank for itself (it stays single-GPU work)."""
    key = DeviceKey()
    key.ranks, key.srs_range = comm, srs_range
    key.n, key.log_n = n, n.bit_length() - 1
    key.omega = int(get_root_of_unity(n))
    # quotient coset: the smallest power-of-two multiple of n with at least 3n+6 points (deg t = 3n+5)
    key.ext = 4 if n >= 8 else (8 if n >= 2 else 16)
    key.log_ext = key.ext.bit_length() - 1
    key.N8 = key.ext * n
    key.omega8 = int(get_root_of_unity(key.N8))
    key.srs_table, key.srs_size = srs_table, srs_size
    key.sigma_evals = list(sigma_evals)
    for name, ev in zip(CIRCUIT_POLYS, list(selector_evals) + list(sigma_evals)):
        c = _alloc_from(ev, n, n)
        native.ntt_dev(c, 0, key.log_n, key.omega, inverse=True)
        key.coeffs[name] = c
        e = _alloc_from(c, n, key.N8)
        native.scalars_convert(e, 0, n, True)
        native.ntt_dev(e, 0, key.log_n + key.log_ext, key.omega8, coset_shift=COSET_SHIFT)
        key.coset[name] = e
    # the eight commitments as one pipelined launch group
    for name, pt in zip(CIRCUIT_POLYS, _commit_many(key, [(key.coeffs[name], 0, n) for name in CIRCUIT_POLYS])):
        key.comm[name] = pt
    key.x = native.scalars_alloc(key.N8)
    key.l1f = native.scalars_alloc(key.N8)
    key.zh8 = native.scalars_alloc(key.ext)
    native.plonk_coset_setup_dev(n, key.ext, COSET_SHIFT, key.omega8, key.x, key.l1f, key.zh8)
    return key


def _blind(h, n, blinds):
    """p(x) += (b0 + b1 x + ...)(x^n - 1): subtract the b's at the bottom, add them at x^n."""
    k = len(blinds)
    up = native.scalars_load(native.fr_vec_bytes([b % R for b in blinds]), k)
    native.vec_op_dev(1, h, 0, h, 0, up, 0, k)
    native.vec_op_dev(0, h, n, h, n, up, 0, k)
    up.free()


def prove(key, a_vals, b_vals, c_vals, blinds=None, keep=False):
    """a_vals, b_vals, c_vals: device handles with the n wire values.  blinds: optional 9 scalars
    (round 1: 2 per wire polynomial, round 2: 3) to reproduce a golden proof; random otherwise.
    Returns a Proof whose fields are the reference's types."""
    n, log_n, w = key.n, key.log_n, key.omega
    N8 = key.N8
    if blinds is None:
        blinds = [secrets.randbelow(R) for _ in range(9)]
    tr = Transcript()
    proof = Proof()
    # ---- round 1
    wires = {}
    for i, (name, vals) in enumerate((("a", a_vals), ("b", b_vals), ("c", c_vals))):
        h = _alloc_from(vals, n, n + 2)
        native.ntt_dev(h, 0, log_n, w, inverse=True)
        _blind(h, n, blinds[2 * i:2 * i + 2])
        wires[name] = h
    for name, pt in zip("abc", _commit_many(key, [(wires[k], 0, n + 2) for k in "abc"])):
        setattr(proof, name + "_comm", g1_from_ints(pt))
    for name in "abc":
        tr.append_point(name.encode() + b"_comm", getattr(proof, name + "_comm"))
    # ---- round 2
    beta, gamma = int(tr.challenge_scalar(b"beta")), int(tr.challenge_scalar(b"gamma"))
    z = native.scalars_alloc(n + 3)
    if n > 1:
        num, den = native.scalars_alloc(n), native.scalars_alloc(n)
        native.plonk_perm_terms_dev(a_vals, b_vals, c_vals, *key.sigma_evals, n, w, beta, gamma, num, den)
        native.batch_inverse_dev(den, 0, n)
        native.vec_op_dev(2, num, 0, num, 0, den, 0, n)
        native.scan_dev(0, z, 0, num, 0, n)               # z_0 = 1, z_{i+1} = z_i * num_i / den_i
        num.free()
        den.free()
    else:
        native.scalars_upload(z, 0, native.fe_bytes(1), 1)
    native.ntt_dev(z, 0, log_n, w, inverse=True)
    _blind(z, n, blinds[6:9])
    # the z commitment starts on the library's second stream; the four coset transforms of round 3 need a, b, c, z
    # but not alpha, so they run beside it and the challenge is drawn once the commitment is back
    _commit_begin(key, z, 0, n + 3)
    ext = []
    for h, length in ((wires["a"], n + 2), (wires["b"], n + 2), (wires["c"], n + 2), (z, n + 3)):
        e = _alloc_from(h, length, N8)
        native.scalars_convert(e, 0, length, True)
        native.ntt_dev(e, 0, log_n + key.log_ext, key.omega8, coset_shift=COSET_SHIFT)
        ext.append(e)
    proof.z_comm = g1_from_ints(_commit_end(key))
    tr.append_point(b"z_comm", proof.z_comm)
    # ---- round 3
    alpha = int(tr.challenge_scalar(b"alpha"))
    t = native.scalars_alloc(N8)
    native.plonk_quotient_dev(ext + [key.coset[k] for k in CIRCUIT_POLYS], n, key.ext, key.x, key.l1f, key.zh8, beta,
                              gamma, alpha, t)
    for e in ext:
        e.free()
    native.ntt_dev(t, 0, log_n + key.log_ext, key.omega8, inverse=True, coset_shift=COSET_SHIFT)
    native.scalars_convert(t, 0, N8, False)
    if not native.scalars_is_zero(t, 3 * n + 6, N8 - (3 * n + 6)):
        raise ValueError(
            "제약 다항식이 Z_H(x)로 나누어 떨어지지 않습니다. "
            "회로 또는 witness에 오류가 있습니다."
        )
    t_pts = _commit_many(key, [(t, 0, n), (t, n, n), (t, 2 * n, n + 6)])
    proof.t_lo_comm, proof.t_mid_comm, proof.t_hi_comm = (g1_from_ints(p) for p in t_pts)
    for part in ("t_lo", "t_mid", "t_hi"):
        tr.append_point(part.encode() + b"_comm", getattr(proof, part + "_comm"))
    # ---- round 4
    zeta = int(tr.challenge_scalar(b"zeta"))
    ev = native.fr_poly_eval_dev
    # the six openings in one batched set of Horner launches
    a_e, b_e, c_e, s1_e, s2_e, zw_e = native.fr_poly_eval_multi_dev(
        [(wires[k], 0, n + 2, zeta) for k in "abc"]
        + [(key.coeffs["s_sigma1"], 0, n, zeta), (key.coeffs["s_sigma2"], 0, n, zeta), (z, 0, n + 3, zeta * w % R)])
    for name, val in (("a_eval", a_e), ("b_eval", b_e), ("c_eval", c_e), ("s_sigma1_eval", s1_e),
                      ("s_sigma2_eval", s2_e), ("z_omega_eval", zw_e)):
        setattr(proof, name, FR(val))
        tr.append_scalar(name.encode(), val)
    # ---- round 5
    v = int(tr.challenge_scalar(b"v"))
    zh_zeta = (pow(zeta, n, R) - 1) % R
    l1_zeta = 1 if zeta == 1 else pow(n, -1, R) * zh_zeta % R * pow((zeta - 1) % R, -1, R) % R
    perm_z = alpha * (a_e + beta * zeta + gamma) % R * (b_e + beta * K1 * zeta + gamma) % R * (c_e + beta * K2 * zeta + gamma) % R
    ab = (a_e + beta * s1_e + gamma) * (b_e + beta * s2_e + gamma) % R
    perm_s3 = alpha * ab % R * beta % R * zw_e % R
    const = (-alpha * ab % R * zw_e % R * (c_e + gamma) - alpha * alpha % R * l1_zeta) % R   # pi(zeta) = 0
    # linearisation polynomial: seven scalar-times-polynomial terms in one pass
    r = native.scalars_alloc(n + 3)
    native.lincomb_dev(r, 0, n + 3,
                       [(k, key.coeffs[name], 0, n) for name, k in
                        (("q_m", a_e * b_e % R), ("q_l", a_e), ("q_r", b_e), ("q_o", c_e), ("q_c", 1),
                         ("s_sigma3", (-perm_s3) % R))]
                       + [((perm_z + alpha * alpha % R * l1_zeta) % R, z, 0, n + 3)])
    native.scalars_add_const(r, 0, 1, const)
    r_eval = ev(r, 0, n + 3, zeta)
    proof.r_eval = FR(r_eval)
    zeta_n = pow(zeta, n, R)
    P = native.scalars_alloc(n + 6)
    terms = [(1, t, 0, n), (zeta_n, t, n, n), (zeta_n * zeta_n % R, t, 2 * n, n + 6), (v, r, 0, n + 3)]
    vp = v
    for h, length in ((wires["a"], n + 2), (wires["b"], n + 2), (wires["c"], n + 2),
                      (key.coeffs["s_sigma1"], n), (key.coeffs["s_sigma2"], n)):
        vp = vp * v % R
        terms.append((vp, h, 0, length))
    native.lincomb_dev(P, 0, n + 6, terms)                # the batched opening polynomial in one pass
    W = native.scalars_alloc(n + 5)
    native.div_linear_dev(P, 0, n + 6, zeta, W, 0)
    Ww = native.scalars_alloc(n + 2)
    native.div_linear_dev(z, 0, n + 3, zeta * w % R, Ww, 0)
    w_pts = _commit_many(key, [(W, 0, n + 5), (Ww, 0, n + 2)])
    proof.W_zeta_comm, proof.W_zeta_omega_comm = (g1_from_ints(p) for p in w_pts)
    state = {"a": wires["a"], "b": wires["b"], "c": wires["c"], "z": z, "t": t, "r": r, "W": W, "Ww": Ww, "P": P,
             "challenges": {"beta": beta, "gamma": gamma, "alpha": alpha, "zeta": zeta, "v": v}}
    if keep:
        return proof, state
    for h in (wires["a"], wires["b"], wires["c"], z, t, r, W, Ww, P):
        h.free()
    return proof
