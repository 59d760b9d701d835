"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's prover hot path over plain
Python ints (Fr elements are ints in [0, r); G1/G2 points as in oracle/bn254.py).

Each function follows the reference function cited in its docstring, statement for statement where
the algorithm matters (recursion shape of fft, long division, the numWires-length quirk of hxr).
It is pinned against the reference's own modules (imported from /root/reference on top of
oracle/shim) by tests/golden/make_golden.py -> tests/golden/*.json, checked in tests/test_oracle.py.
Never imported by the product package.
"""
from . import bn254

R = bn254.R
P = bn254.P


def inv(a):
    """FR division semantics: py_ecc prime_field_inv, inv(0) = 0."""
    return bn254.inv(a, R)


# ------------------------------------------------------------------ zkp/plonk/field.py
def get_root_of_unity(n):
    """/root/reference/zkp/plonk/field.py:145-182: omega = 5^((r-1)/n), n a power of two <= 2^28."""
    if n < 1 or (n & (n - 1)) != 0:
        raise ValueError("n must be a power of two: %d" % n)
    if n > (1 << 28):
        raise ValueError("n must be <= 2^28: %d" % n)
    if n == 1:
        return 1
    return pow(5, (R - 1) // n, R)


def get_roots_of_unity(n):
    """/root/reference/zkp/plonk/field.py:185-209."""
    w = get_root_of_unity(n)
    out, cur = [], 1
    for _ in range(n):
        out.append(cur)
        cur = cur * w % R
    return out


# ------------------------------------------------------------------ zkp/plonk/polynomial.py
def fft(coeffs, omega):
    """/root/reference/zkp/plonk/polynomial.py:292-341: recursive radix-2 DIT, natural order."""
    n = len(coeffs)
    if n == 1:
        return [coeffs[0] % R]
    even_vals = fft(coeffs[0::2], omega * omega % R)
    odd_vals = fft(coeffs[1::2], omega * omega % R)
    result = [0] * n
    omega_k = 1
    half = n // 2
    for k in range(half):
        t = omega_k * odd_vals[k] % R
        result[k] = (even_vals[k] + t) % R
        result[k + half] = (even_vals[k] - t) % R
        omega_k = omega_k * omega % R
    return result


def ifft(evals, omega):
    """/root/reference/zkp/plonk/polynomial.py:344-378: fft with omega^-1, then times n^-1."""
    n = len(evals)
    coeffs = fft(evals, inv(omega))
    n_inv = inv(n)
    return [c * n_inv % R for c in coeffs]


def trim(coeffs):
    """Polynomial._trim, /root/reference/zkp/plonk/polynomial.py:66-72 (zero polynomial = [0])."""
    c = list(coeffs) if coeffs else [0]
    while len(c) > 1 and c[-1] == 0:
        c.pop()
    return c


def poly_eval(coeffs, x):
    """Polynomial.evaluate (Horner), /root/reference/zkp/plonk/polynomial.py:85-106."""
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


def poly_add(a, b):
    """Polynomial.__add__, polynomial.py:108-118 (trimmed)."""
    m = max(len(a), len(b))
    return trim([((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % R for i in range(m)])


def poly_sub(a, b):
    """Polynomial.__sub__, polynomial.py:123-133 (trimmed)."""
    m = max(len(a), len(b))
    return trim([((a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0)) % R for i in range(m)])


def poly_scale(a, k):
    return trim([c * k % R for c in a])


def poly_mul(a, b):
    """Polynomial.__mul__ (schoolbook), /root/reference/zkp/plonk/polynomial.py:144-159 (trimmed)."""
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                out[i + j] = (out[i + j] + x * y) % R
    return trim(out)


def poly_div(a, b):
    """poly_div (long division), /root/reference/zkp/plonk/polynomial.py:385-435 -> (q, r) trimmed."""
    a, b = trim(a), trim(b)
    if b == [0]:
        raise ValueError("division by the zero polynomial")
    rem = list(a)
    deg_b, deg_a = len(b) - 1, len(rem) - 1
    if deg_a < deg_b:
        return [0], trim(rem)
    quot = [0] * (deg_a - deg_b + 1)
    lead_inv = inv(b[-1])
    for i in range(deg_a - deg_b, -1, -1):
        coeff = rem[i + deg_b] * lead_inv % R
        quot[i] = coeff
        for j in range(deg_b + 1):
            rem[i + j] = (rem[i + j] - coeff * b[j]) % R
    return trim(quot), trim(rem)


# ------------------------------------------------------------------ zkp/plonk/utils.py
def coset_fft(coeffs, omega, k=5):
    """/root/reference/zkp/plonk/utils.py:145-176: c_i <- c_i * k^i, then fft."""
    shifted, kp = [], 1
    for c in coeffs:
        shifted.append(c * kp % R)
        kp = kp * k % R
    return fft(shifted, omega)


def coset_ifft(evals, omega, k=5):
    """/root/reference/zkp/plonk/utils.py:179-205: ifft, then c_i <- c_i * k^-i."""
    coeffs = ifft(evals, omega)
    k_inv, kp, out = inv(k), 1, []
    for c in coeffs:
        out.append(c * kp % R)
        kp = kp * k_inv % R
    return out


# ------------------------------------------------------------------ zkp/plonk/kzg.py
def commit(coeffs, g1_powers, max_degree=None):
    """commit, /root/reference/zkp/plonk/kzg.py:32-67 on trimmed coefficients."""
    coeffs = trim(coeffs)
    degree = 0 if coeffs == [0] else len(coeffs) - 1
    if max_degree is None:
        max_degree = len(g1_powers) - 1
    if degree > max_degree:
        raise ValueError("polynomial degree %d exceeds the SRS degree %d" % (degree, max_degree))
    acc = None
    for i, c in enumerate(coeffs):
        if c == 0:
            continue
        acc = bn254.g1_add(acc, bn254.g1_mul(g1_powers[i], c % R))
    return acc


def srs_generate(max_degree, seed):
    """SRS.generate, /root/reference/zkp/plonk/srs.py:50-87 (seeded branch)."""
    import hashlib
    tau = int.from_bytes(hashlib.sha256(str(seed).encode()).digest(), "big") % R
    g1_powers, tp = [], 1
    for _ in range(max_degree + 1):
        g1_powers.append(bn254.g1_mul(bn254.G1, tp % R))
        tp = tp * tau % R
    g2_powers = [bn254.G2, bn254.g2_mul(bn254.G2, tau)]
    return g1_powers, g2_powers, tau


# ------------------------------------------------------------------ zkp/plonk/permutation.py
K1, K2 = 2, 3


def compute_accumulator(a_vals, b_vals, c_vals, s1, s2, s3, n, domain, beta, gamma):
    """compute_accumulator, /root/reference/zkp/plonk/permutation.py:89-137 (s1..s3 = evaluations of
    the permutation polynomials, build_permutation_polynomials :44-86)."""
    z = [1]
    for i in range(n - 1):
        num = ((a_vals[i] + beta * domain[i] + gamma)
               * (b_vals[i] + beta * K1 * domain[i] + gamma)
               * (c_vals[i] + beta * K2 * domain[i] + gamma)) % R
        den = ((a_vals[i] + beta * s1[i] + gamma)
               * (b_vals[i] + beta * s2[i] + gamma)
               * (c_vals[i] + beta * s3[i] + gamma)) % R
        z.append(z[-1] * num % R * inv(den) % R)
    return z


# ------------------------------------------------------------------ zkp/groth16/poly_utils.py
def g16_multiply_polys(a, b):
    """_multiply_polys, /root/reference/zkp/groth16/poly_utils.py:17-22 (no trimming)."""
    o = [0] * (len(a) + len(b) - 1)
    for i in range(len(a)):
        for j in range(len(b)):
            o[i + j] = (o[i + j] + a[i] * b[j]) % R
    return o


def g16_subtract_polys(a, b):
    """_subtract_polys / _add_polys, poly_utils.py:25-34."""
    o = [0] * max(len(a), len(b))
    for i in range(len(a)):
        o[i] = (o[i] + a[i]) % R
    for i in range(len(b)):
        o[i] = (o[i] - b[i]) % R
    return o


def g16_div_polys(a, b):
    """_div_polys, poly_utils.py:37-45: quotient of len(a)-len(b)+1 entries, remainder len(b)-1."""
    o = [0] * (len(a) - len(b) + 1)
    remainder = list(a)
    while len(remainder) >= len(b):
        leading_fac = remainder[-1] * inv(b[-1]) % R
        pos = len(remainder) - len(b)
        o[pos] = leading_fac
        sub = g16_multiply_polys(b, [0] * pos + [leading_fac])
        remainder = g16_subtract_polys(remainder, sub)[:-1]
    return o, remainder


def g16_multiply_vec_matrix(vec, matrix):
    """_multiply_vec_matrix, poly_utils.py:52-59: result has len(vec) entries (numWires), of which
    the first numGates are filled (the reference's length quirk)."""
    assert not len(vec) == len(matrix[0])
    target = [0] * len(vec)
    for i in range(len(matrix)):
        for j in range(len(matrix[0])):
            target[j] = (target[j] + vec[i] * matrix[i][j]) % R
    return target


def hxr(Ax, Bx, Cx, Zx, Rvec):
    """hxr, /root/reference/zkp/groth16/poly_utils.py:116-125 -> (Hx, remainder)."""
    Rax = g16_multiply_vec_matrix(Rvec, Ax)
    Rbx = g16_multiply_vec_matrix(Rvec, Bx)
    Rcx = g16_multiply_vec_matrix(Rvec, Cx)
    Px = g16_subtract_polys(g16_multiply_polys(Rax, Rbx), Rcx)
    return g16_div_polys(Px, Zx)


# ------------------------------------------------------------------ zkp/groth16/proving.py
def _nested_msm(first, table, M, Rx, add, mul):
    """The numWires x numGates loop shared by proof_a / proof_b / proof_c, proving.py:27-31."""
    acc = first
    for i in range(len(M)):
        temp = None
        for j in range(len(M[0])):
            temp = add(temp, mul(table[j], int(M[i][j])))
        acc = add(acc, mul(temp, int(Rx[i])))
    return acc


def proof_a(sigma1_1, sigma1_2, Ax, Rx, r):
    """proof_a, /root/reference/zkp/groth16/proving.py:23-33."""
    acc = _nested_msm(sigma1_1[0], sigma1_2, Ax, Rx, bn254.g1_add, bn254.g1_mul)
    return bn254.g1_add(acc, bn254.g1_mul(sigma1_1[2], int(r)))


def proof_b(sigma2_1, sigma2_2, Bx, Rx, s):
    """proof_b, /root/reference/zkp/groth16/proving.py:35-45 (G2)."""
    acc = _nested_msm(sigma2_1[0], sigma2_2, Bx, Rx, bn254.g2_add, bn254.g2_mul)
    return bn254.g2_add(acc, bn254.g2_mul(sigma2_1[2], int(s)))


def proof_c(sigma1_1, sigma1_2, sigma1_4, sigma1_5, Bx, Rx, Hx, s, r, prf_A, pub_r_indexs=None):
    """proof_c, /root/reference/zkp/groth16/proving.py:47-75."""
    if pub_r_indexs is None:
        pub_r_indexs = [0, 1]
    add, mul, neg = bn254.g1_add, bn254.g1_mul, bn254.g1_neg
    num_gates, num_wires = len(Bx[0]), len(Bx)
    temp_b = _nested_msm(sigma1_1[1], sigma1_2, Bx, Rx, add, mul)
    temp_b = add(temp_b, mul(sigma1_1[2], int(s)))
    c = add(add(mul(prf_A, int(s)), mul(temp_b, int(r))), neg(mul(mul(sigma1_1[2], int(s)), int(r))))
    for i in range(num_wires):
        if i in pub_r_indexs:
            continue
        c = add(c, mul(sigma1_4[i], int(Rx[i])))
    for i in range(num_gates - 1):
        c = add(c, mul(sigma1_5[i], int(Hx[i])))
    return c


# ------------------------------------------------------------------ zkp/groth16/qap_creator_lcm.py (exact restatement)
def qap_mk_singleton(point_loc, height, total_pts):
    """/root/reference/zkp/groth16/qap_creator_lcm.py:50-66: the polynomial that is `height` at x = point_loc
    and zero at the other points of {1..total_pts}.  The reference computes it in floats (height / fac);
    here the division is the Fr inverse, which agrees with its rounded, determinant-scaled result wherever
    the floats are exact (golden: tests/golden/groth16_qap.json)."""
    fac = 1
    for i in range(1, total_pts + 1):
        if i != point_loc:
            fac = fac * (point_loc - i) % R
    o = [height % R * inv(fac) % R]
    for i in range(1, total_pts + 1):
        if i != point_loc:
            o = g16_multiply_polys(o, [(-i) % R, 1])
    return o


def qap_lagrange_interp(vec):
    """qap_creator_lcm.py:70-78: sum of the singletons; vec[i] = p(i + 1)."""
    o = [0] * len(vec)
    for i, v in enumerate(vec):
        if v % R:
            term = qap_mk_singleton(i + 1, v, len(vec))
            o = [(a + b) % R for a, b in zip(o, term)]
    return o


def qap_vandermonde_det(k):
    """determinant_fast(k_matrix(k)) of qap_creator_lcm.py:97-121: the Vandermonde determinant of the points
    1..k, prod_{i<j} (j - i)."""
    d = 1
    for j in range(1, k + 1):
        for i in range(1, j):
            d = d * (j - i) % R
    return d


def r1cs_to_qap_times_lcm(A, B, C):
    """qap_creator_lcm.py:114-135: per-wire interpolation of the transposed R1CS, A and B times det, C times
    det^2; Z = (x-1)...(x-k).  Returns (Ax, Bx, Cx, Zx) over Fr (what getFRPoly2D/1D make of the floats)."""
    k = len(A)
    det = qap_vandermonde_det(k)
    cols = lambda M: [list(c) for c in zip(*M)]
    new_a = [[c * det % R for c in qap_lagrange_interp(col)] for col in cols(A)]
    new_b = [[c * det % R for c in qap_lagrange_interp(col)] for col in cols(B)]
    new_c = [[c * det % R * det % R for c in qap_lagrange_interp(col)] for col in cols(C)]
    Z = [1]
    for i in range(1, k + 1):
        Z = g16_multiply_polys(Z, [(-i) % R, 1])
    return new_a, new_b, new_c, Z


def qap_eval_rows(M, x):
    """poly_utils.py:86-106 ax_val / bx_val / cx_val: every wire polynomial at x."""
    return [poly_eval(row, x) for row in M]
