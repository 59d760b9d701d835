"""TEST INFRASTRUCTURE (oracle) -- restatement of the reference's PLONK verifier
(/root/reference/zkp/plonk/verifier.py:42-208) over plain ints, the plain-int group law of
oracle/bn254.py and the pairing of oracle/shim/py_ecc.  It is the acceptance check for proofs
produced on the GPU at sizes where no golden proof exists.  Pinned in tests/test_oracle.py: it accepts
every reference-minted golden proof and rejects tampered ones (as the reference's test_e2e.py does).
"""
import hashlib

from . import bn254

R = bn254.R
K1, K2 = 2, 3


def _pairing():
    """The oracle's py_ecc restatement, imported under its own package path: it is never put on
    sys.path as `py_ecc`, so the product's `import py_ecc` probe (compat.py) cannot pick it up."""
    from .shim.py_ecc import bn128
    return bn128


class _Transcript:
    """/root/reference/zkp/plonk/transcript.py:47-123."""

    def __init__(self):
        self.state = bytearray(b"plonk")

    def point(self, label, p):
        self.state += label
        self.state += bytes(64) if p is None else p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")

    def scalar(self, label, s):
        self.state += label + (s % R).to_bytes(32, "big")

    def challenge(self, label):
        self.state += label
        h = hashlib.sha256(bytes(self.state)).digest()
        self.state += h
        return int.from_bytes(h, "big") % R


def verify(proof, pre_comm, n, omega, g2_powers):
    """proof: dict of the 16 proof fields (points as int pairs / None, evaluations as ints);
    pre_comm: dict q_l,q_r,q_o,q_m,q_c,s_sigma1,s_sigma2,s_sigma3 -> G1 int pairs;
    g2_powers: [G2, tau*G2] as nested int pairs."""
    add, mul, neg, inv = bn254.g1_add, bn254.g1_mul, bn254.g1_neg, lambda a: bn254.inv(a, R)
    mulr = lambda p, k: mul(p, k % R)
    t = _Transcript()
    for k in ("a_comm", "b_comm", "c_comm"):
        t.point(k.encode(), proof[k])
    beta, gamma = t.challenge(b"beta"), t.challenge(b"gamma")
    t.point(b"z_comm", proof["z_comm"])
    alpha = t.challenge(b"alpha")
    for k in ("t_lo_comm", "t_mid_comm", "t_hi_comm"):
        t.point(k.encode(), proof[k])
    zeta = t.challenge(b"zeta")
    for k in ("a_eval", "b_eval", "c_eval", "s_sigma1_eval", "s_sigma2_eval", "z_omega_eval"):
        t.scalar(k.encode(), proof[k])
    v, u = t.challenge(b"v"), t.challenge(b"u")
    a_e, b_e, c_e = proof["a_eval"], proof["b_eval"], proof["c_eval"]
    s1_e, s2_e, zw_e = proof["s_sigma1_eval"], proof["s_sigma2_eval"], proof["z_omega_eval"]
    zh = (pow(zeta, n, R) - 1) % R
    den = (zeta - 1) % R
    l1 = 1 if den == 0 else inv(n) * zh % R * inv(den) % R   # lagrange_basis_eval(0, ...), utils.py:45-81
    D = mulr(pre_comm["q_m"], a_e * b_e)
    D = add(D, mulr(pre_comm["q_l"], a_e))
    D = add(D, mulr(pre_comm["q_r"], b_e))
    D = add(D, mulr(pre_comm["q_o"], c_e))
    D = add(D, pre_comm["q_c"])
    perm_z = alpha * (a_e + beta * zeta + gamma) % R * (b_e + beta * K1 * zeta + gamma) % R * (c_e + beta * K2 * zeta + gamma) % R
    D = add(D, mulr(proof["z_comm"], perm_z))
    ab = (a_e + beta * s1_e + gamma) * (b_e + beta * s2_e + gamma) % R
    D = add(D, neg(mulr(pre_comm["s_sigma3"], alpha * ab % R * beta % R * zw_e)))
    D = add(D, mulr(proof["z_comm"], alpha * alpha % R * l1))
    r0 = (-alpha * ab % R * zw_e % R * (c_e + gamma) - alpha * alpha % R * l1) % R
    zn = pow(zeta, n, R)
    F = add(proof["t_lo_comm"], add(mulr(proof["t_mid_comm"], zn), mulr(proof["t_hi_comm"], zn * zn)))
    F = add(F, mulr(D, v))
    F = add(F, mulr(bn254.G1, v * r0))
    vp = v * v % R
    for key, src in (("a_comm", proof), ("b_comm", proof), ("c_comm", proof), ("s_sigma1", pre_comm), ("s_sigma2", pre_comm)):
        F = add(F, mulr(src[key], vp))
        vp = vp * v % R
    r_eval = proof["r_eval"]
    e = (r_eval * inv(zh) + v * r_eval) % R
    vp = v * v % R
    for ev in (a_e, b_e, c_e, s1_e, s2_e):
        e = (e + vp * ev) % R
        vp = vp * v % R
    e = (e + u * zw_e) % R
    A = add(proof["W_zeta_comm"], mulr(proof["W_zeta_omega_comm"], u))
    B = mulr(proof["W_zeta_comm"], zeta)
    B = add(B, mulr(proof["W_zeta_omega_comm"], u * zeta % R * omega))
    B = add(B, F)
    B = add(B, mulr(proof["z_comm"], u))
    B = add(B, neg(mulr(bn254.G1, e)))
    bn = _pairing()
    to1 = lambda p: None if p is None else (bn.FQ(p[0]), bn.FQ(p[1]))
    to2 = lambda p: (bn.FQ2([p[0][0], p[0][1]]), bn.FQ2([p[1][0], p[1][1]]))
    return bn.pairing(to2(g2_powers[1]), to1(A)) == bn.pairing(to2(g2_powers[0]), to1(B))
