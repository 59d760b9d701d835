"""Drop-in for /root/reference/zkp/plonk/polynomial.py: Polynomial, fft, ifft, poly_div,
lagrange_basis -- same API, coefficient lists of FR, auto-trimming; the arithmetic runs on the GPU.

  Polynomial.evaluate (:85-106, Horner)          -> zkp_fr_poly_eval
  Polynomial.__add__/__sub__/__neg__/scalar mul   -> zkp_fr_vec_op
  Polynomial.__mul__ (:144-159, O(n^2))           -> zkp_fr_poly_mul   (three NTTs)
  fft (:292-341, recursive) / ifft (:344-378)     -> zkp_fr_ntt        (natural order in and out)
  poly_div (:385-435, long division)              -> zkp_fr_poly_divmod (Newton inverse, or the
                                                     stride recurrence when the divisor is x^n - 1)
Products and quotients/remainders are unique, so every returned coefficient equals the reference's.
"""
from ... import native
from .field import FR, CURVE_ORDER

_enc = native.fr_vec_bytes
_dec = native.fr_vec_from_bytes


def _ints(coeffs):
    return [int(c) % CURVE_ORDER for c in coeffs]


class Polynomial:
    """Polynomial over FR in coefficient form: coeffs = [c0, c1, ...] (reference :34-72)."""

    def __init__(self, coeffs=None):
        if coeffs is None:
            self.coeffs = [FR(0)]
        else:
            self.coeffs = [c if isinstance(c, FR) else FR(c) for c in coeffs]
        self._trim()

    @classmethod
    def _from_ints(cls, ints):
        p = cls.__new__(cls)
        p.coeffs = [FR(v) for v in ints] or [FR(0)]
        p._trim()
        return p

    def _trim(self):
        c = self.coeffs
        while len(c) > 1 and c[-1].n == 0:
            c.pop()

    @property
    def degree(self):
        if len(self.coeffs) == 1 and self.coeffs[0] == FR(0):
            return 0
        return len(self.coeffs) - 1

    def is_zero(self):
        return len(self.coeffs) == 1 and self.coeffs[0] == FR(0)

    def evaluate(self, point):
        if not isinstance(point, FR):
            point = FR(point)
        c = _ints(self.coeffs)
        return FR(native.fr_poly_eval(_enc(c), len(c), int(point)))

    def _binary(self, other, op):
        if isinstance(other, (int, FR)):
            other = Polynomial([other])
        n = max(len(self.coeffs), len(other.coeffs))
        a = _ints(self.coeffs) + [0] * (n - len(self.coeffs))
        b = _ints(other.coeffs) + [0] * (n - len(other.coeffs))
        return Polynomial._from_ints(_dec(native.fr_vec_op(op, _enc(a), _enc(b), n)))

    def __add__(self, other):
        return self._binary(other, 0)

    def __radd__(self, other):
        return self.__add__(other)

    def __sub__(self, other):
        return self._binary(other, 1)

    def __rsub__(self, other):
        if isinstance(other, (int, FR)):
            other = Polynomial([other])
        return other.__sub__(self)

    def __neg__(self):
        return Polynomial([FR(0)] * len(self.coeffs))._binary(self, 1)

    def __mul__(self, other):
        if isinstance(other, (int, FR)):
            k = int(other) % CURVE_ORDER
            a = _ints(self.coeffs)
            return Polynomial._from_ints(_dec(native.fr_vec_op(3, _enc(a), native.fe_bytes(k), len(a))))
        a, b = _ints(self.coeffs), _ints(other.coeffs)
        return Polynomial._from_ints(_dec(native.fr_poly_mul(_enc(a), len(a), _enc(b), len(b))))

    def __rmul__(self, other):
        return self.__mul__(other)

    def __eq__(self, other):
        if isinstance(other, (int, FR)):
            other = Polynomial([other])
        if not isinstance(other, Polynomial):
            return False
        return self.coeffs == other.coeffs

    def __repr__(self):
        terms = []
        for i, c in enumerate(self.coeffs):
            if c == FR(0):
                continue
            terms.append(str(int(c)) if i == 0 else (f"{int(c)}*x" if i == 1 else f"{int(c)}*x^{i}"))
        return "Poly(" + " + ".join(terms) + ")" if terms else "Poly(0)"

    def __len__(self):
        return len(self.coeffs)

    def scale(self, scalar):
        return self * scalar

    def divide_by_vanishing(self, n):
        q, r = poly_div(self, Polynomial._vanishing_coeffs(n))
        for c in r.coeffs:
            if c != FR(0):
                raise ValueError("소거 다항식으로 나누어 떨어지지 않습니다 (제약 불만족)")
        return q

    @staticmethod
    def _vanishing_coeffs(n):
        coeffs = [FR(0)] * (n + 1)
        coeffs[0] = FR(FR.field_modulus - 1)
        coeffs[n] = FR(1)
        return Polynomial(coeffs)

    @classmethod
    def zero(cls):
        return cls([FR(0)])

    @classmethod
    def one(cls):
        return cls([FR(1)])

    @classmethod
    def vanishing(cls, n):
        return cls._vanishing_coeffs(n)

    @classmethod
    def from_evaluations(cls, evals, omega):
        return cls(ifft(evals, omega))


def _ntt(values, omega, inverse, shift=None):
    n = len(values)
    if n < 1 or n & (n - 1):
        raise ValueError("fft input length must be a power of two")
    out = native.fr_ntt(_enc(_ints(values)), n.bit_length() - 1, int(omega) % CURVE_ORDER, inverse,
                        None if shift is None else int(shift) % CURVE_ORDER)
    return [FR(v) for v in _dec(out)]


def fft(coeffs, omega):
    """Coefficients -> evaluations at 1, w, ..., w^(n-1) (reference :292-341)."""
    return _ntt(coeffs, omega, False)


def ifft(evals, omega):
    """Evaluations -> coefficients: transform with w^-1, times n^-1 (reference :344-378)."""
    return _ntt(evals, omega, True)


def poly_div(a, b):
    """a = b*q + r (reference :385-435)."""
    if b.is_zero():
        raise ValueError("0으로 나눌 수 없습니다")
    if len(a.coeffs) < len(b.coeffs):
        return Polynomial.zero(), Polynomial(list(a.coeffs))
    ai, bi = _ints(a.coeffs), _ints(b.coeffs)
    q, r = native.fr_poly_divmod(_enc(ai), len(ai), _enc(bi), len(bi))
    return Polynomial._from_ints(_dec(q)), Polynomial._from_ints(_dec(r))


def lagrange_basis(domain, i):
    """L_i(x) = prod_{j != i} (x - d_j) / (d_i - d_j) in coefficient form (reference :438-475).

    The reference multiplies the n-1 linear factors one after another (O(n^2)); here they are folded
    pairwise (a product tree of GPU NTT products) and the denominator is one device inversion."""
    n = len(domain)
    d = _ints(domain)
    leaves = [Polynomial._from_ints([(-d[j]) % CURVE_ORDER, 1]) for j in range(n) if j != i]
    if not leaves:
        return Polynomial([FR(1)])
    while len(leaves) > 1:
        nxt = [leaves[k] * leaves[k + 1] for k in range(0, len(leaves) - 1, 2)]
        if len(leaves) & 1:
            nxt.append(leaves[-1])
        leaves = nxt
    # denominator prod_{j != i} (d_i - d_j) = N(d_i) with N = prod_{j != i} (x - d_j), already built
    den = leaves[0].evaluate(FR(d[i]))
    inv = native.fr_vec_from_bytes(native.fr_batch_inverse(native.fe_bytes(int(den)), 1))[0]
    return leaves[0] * FR(inv)
