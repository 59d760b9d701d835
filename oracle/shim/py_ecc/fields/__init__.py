"""TEST INFRASTRUCTURE (oracle) -- restatement of ``py_ecc.fields`` (py-ecc==7.0.1).

py-ecc is a third-party dependency of the reference
(/root/reference/requirements.txt:14), is not vendored under /root/reference
and cannot be installed in this image (no network, no wheel).  This module
restates the published semantics of its non-optimised field classes as used
by the reference:

  * ``FQ``: call sites /root/reference/zkp/plonk/field.py:36-51 (``FR(FQ)``
    with ``field_modulus = curve_order``), zkp/groth16/poly_utils.py:12-13,
    zkp/groth16/setup.py:39 (``(FQ(0), FQ(0))`` placeholders).
  * ``FQ2``: /root/reference/plonk_serializers.py:56-67 (``.coeffs[i]``).
  * ``FQ12``: pairing target group (zkp/groth16/verifying.py:29-40).

Semantics restated: ``FQ(int)`` reduces mod
``field_modulus``; binary ops accept ``int`` or ``FQ`` and return
``type(self)``; ``/`` multiplies by the modular inverse with ``inv(0) == 0``;
``**`` is square-and-multiply; ``==`` compares ``.n`` (ints allowed, anything
else raises ``TypeError``); extension fields are polynomial quotient rings
with ``coeffs`` a tuple of base-field elements.
"""

field_modulus = 21888242871839275222246405745257275088696311157297823662689037894645226208583

# w^2 + 1 = 0   and   w^12 - 18 w^6 + 82 = 0
FQ2_MODULUS_COEFFS = (1, 0)
FQ12_MODULUS_COEFFS = (82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0)


def prime_field_inv(a, n):
    """Extended-Euclid inverse with inv(0) = 0 (py_ecc.utils.prime_field_inv)."""
    a %= n
    if a == 0:
        return 0
    lm, hm = 1, 0
    low, high = a, n
    while low > 1:
        r = high // low
        nm, new = hm - lm * r, high - low * r
        lm, low, hm, high = nm, new, lm, low
    return lm % n


def _as_int(other, cls_name):
    if isinstance(other, FQ):
        return other.n
    if isinstance(other, int):
        return other
    raise TypeError(
        "Expected an int or FQ object, but got object of type {}".format(type(other))
    )


class FQ(object):
    """Element of a prime field; subclasses set ``field_modulus``."""

    n = None
    field_modulus = None

    def __init__(self, val):
        if self.field_modulus is None:
            raise AttributeError("Field Modulus hasn't been specified")
        if isinstance(val, FQ):
            self.n = val.n
        elif isinstance(val, int):
            self.n = val % self.field_modulus
        else:
            raise TypeError(
                "Expected an int or FQ object, but got object of type {}".format(type(val))
            )

    def __add__(self, other):
        on = _as_int(other, "FQ")
        return type(self)((self.n + on) % self.field_modulus)

    def __mul__(self, other):
        on = _as_int(other, "FQ")
        return type(self)((self.n * on) % self.field_modulus)

    def __rmul__(self, other):
        return self * other

    def __radd__(self, other):
        return self + other

    def __rsub__(self, other):
        on = _as_int(other, "FQ")
        return type(self)((on - self.n) % self.field_modulus)

    def __sub__(self, other):
        on = _as_int(other, "FQ")
        return type(self)((self.n - on) % self.field_modulus)

    def __mod__(self, other):
        raise NotImplementedError("Modulo Operation not yet supported by fields")

    def __div__(self, other):
        on = _as_int(other, "FQ")
        return type(self)(
            self.n * prime_field_inv(on, self.field_modulus) % self.field_modulus
        )

    def __truediv__(self, other):
        return self.__div__(other)

    def __rdiv__(self, other):
        on = _as_int(other, "FQ")
        return type(self)(
            prime_field_inv(self.n, self.field_modulus) * on % self.field_modulus
        )

    def __rtruediv__(self, other):
        return self.__rdiv__(other)

    def __pow__(self, other):
        if other == 0:
            return type(self)(1)
        elif other == 1:
            return type(self)(self.n)
        elif other % 2 == 0:
            return (self * self) ** (other // 2)
        else:
            return ((self * self) ** int(other // 2)) * self

    def __eq__(self, other):
        if isinstance(other, FQ):
            return self.n == other.n
        elif isinstance(other, int):
            return self.n == other
        else:
            raise TypeError(
                "Expected an int or FQ object, but got object of type {}".format(type(other))
            )

    def __ne__(self, other):
        return not self == other

    def __neg__(self):
        return type(self)(-self.n)

    def __repr__(self):
        return repr(self.n)

    def __int__(self):
        return self.n

    @classmethod
    def one(cls):
        return cls(1)

    @classmethod
    def zero(cls):
        return cls(0)


def _deg(p):
    d = len(p) - 1
    while d and p[d] == 0:
        d -= 1
    return d


def _poly_rounded_div(a, b, mod):
    dega = _deg(a)
    degb = _deg(b)
    temp = [x for x in a]
    o = [0 for _ in a]
    for i in range(dega - degb, -1, -1):
        o[i] = (o[i] + temp[degb + i] * prime_field_inv(b[degb], mod)) % mod
        for c in range(degb + 1):
            temp[c + i] = (temp[c + i] - o[i] * b[c]) % mod
    return [x % mod for x in o[: _deg(o) + 1]]


class FQP(object):
    """Polynomial extension field F_q[w]/(w^deg + modulus_coeffs)."""

    degree = 0
    field_modulus = None
    FQP_corresponding_FQ_class = None
    modulus_coeffs = None

    def __init__(self, coeffs, modulus_coeffs=()):
        if self.field_modulus is None:
            raise AttributeError("Field Modulus hasn't been specified")
        if len(coeffs) != len(modulus_coeffs):
            raise Exception("coeffs and modulus_coeffs aren't of the same length")
        if self.FQP_corresponding_FQ_class is None:
            type(self).FQP_corresponding_FQ_class = type(
                "FQP_corresponding_FQ_class", (FQ,), {"field_modulus": self.field_modulus}
            )
        fq = self.FQP_corresponding_FQ_class
        self.coeffs = tuple(fq(c) for c in coeffs)
        self.modulus_coeffs = tuple(modulus_coeffs)
        self.degree = len(self.modulus_coeffs)

    # internal: build from already-reduced ints without re-validation
    def _from_ints(self, ints):
        return type(self)(ints)

    def _ints(self):
        return [c.n for c in self.coeffs]

    def __add__(self, other):
        if not isinstance(other, type(self)):
            raise TypeError("Expected an FQP object")
        return self._from_ints([x.n + y.n for x, y in zip(self.coeffs, other.coeffs)])

    def __sub__(self, other):
        if not isinstance(other, type(self)):
            raise TypeError("Expected an FQP object")
        return self._from_ints([x.n - y.n for x, y in zip(self.coeffs, other.coeffs)])

    def __mod__(self, other):
        raise NotImplementedError("Modulo Operation not yet supported by fields")

    def __mul__(self, other):
        if isinstance(other, (int, FQ)):
            on = other.n if isinstance(other, FQ) else other
            return self._from_ints([c.n * on for c in self.coeffs])
        elif isinstance(other, FQP):
            deg = self.degree
            p = self.field_modulus
            a = self._ints()
            bb = other._ints()
            b = [0] * (deg * 2 - 1)
            for i, ai in enumerate(a):
                if ai:
                    for j, bj in enumerate(bb):
                        b[i + j] += ai * bj
            mc = self.modulus_coeffs
            nz = [(i, c) for i, c in enumerate(mc) if c]
            while len(b) > deg:
                exp, top = len(b) - deg - 1, b.pop() % p
                if top:
                    for i, c in nz:
                        b[exp + i] -= top * c
            return self._from_ints(b)
        else:
            raise TypeError("Expected an int or FQ object or FQP object")

    def __rmul__(self, other):
        return self * other

    def __div__(self, other):
        if isinstance(other, (int, FQ)):
            on = other.n if isinstance(other, FQ) else other
            inv = prime_field_inv(on, self.field_modulus)
            return self._from_ints([c.n * inv for c in self.coeffs])
        elif isinstance(other, FQP):
            return self * other.inv()
        else:
            raise TypeError("Expected an int or FQ object or FQP object")

    def __truediv__(self, other):
        return self.__div__(other)

    def __pow__(self, other):
        o = type(self).one()
        t = self
        while other > 0:
            if other & 1:
                o = o * t
            other >>= 1
            if other:
                t = t * t
        return o

    def inv(self):
        """Inverse in F_q[w]/(modulus): extended Euclid over F_q[w]; the result is
        the unique inverse, so any correct algorithm is bit-identical to py_ecc's."""
        p = self.field_modulus
        deg = self.degree
        if deg == 2 and self.modulus_coeffs == (1, 0):
            a, b = self.coeffs[0].n, self.coeffs[1].n
            d = prime_field_inv((a * a + b * b) % p, p)
            return self._from_ints([a * d, -b * d])
        lm, hm = [1] + [0] * deg, [0] * (deg + 1)
        low = self._ints() + [0]
        high = [c % p for c in self.modulus_coeffs] + [1]
        while _deg(low):
            r = _poly_rounded_div(high, low, p)
            r += [0] * (deg + 1 - len(r))
            nm = [x for x in hm]
            new = [x for x in high]
            for i in range(deg + 1):
                for j in range(deg + 1 - i):
                    nm[i + j] -= lm[i] * r[j]
                    new[i + j] -= low[i] * r[j]
            nm = [x % p for x in nm]
            new = [x % p for x in new]
            lm, low, hm, high = nm, new, lm, low
        inv0 = prime_field_inv(low[0], p)
        return self._from_ints([x * inv0 for x in lm[:deg]])

    def __repr__(self):
        return repr(self.coeffs)

    def __eq__(self, other):
        if not isinstance(other, type(self)):
            raise TypeError("Expected an FQP object")
        for c1, c2 in zip(self.coeffs, other.coeffs):
            if c1 != c2:
                return False
        return True

    def __ne__(self, other):
        return not self == other

    def __neg__(self):
        return self._from_ints([-c.n for c in self.coeffs])

    @classmethod
    def one(cls):
        return cls([1] + [0] * (cls.degree - 1))

    @classmethod
    def zero(cls):
        return cls([0] * cls.degree)


class FQ2(FQP):
    degree = 2
    FQ2_MODULUS_COEFFS = None

    def __init__(self, coeffs):
        if self.FQ2_MODULUS_COEFFS is None:
            raise AttributeError("FQ2 Modulus Coeffs haven't been specified")
        super().__init__(coeffs, self.FQ2_MODULUS_COEFFS)


class FQ12(FQP):
    degree = 12
    FQ12_MODULUS_COEFFS = None

    def __init__(self, coeffs):
        if self.FQ12_MODULUS_COEFFS is None:
            raise AttributeError("FQ12 Modulus Coeffs haven't been specified")
        super().__init__(coeffs, self.FQ12_MODULUS_COEFFS)


class bn128_FQ(FQ):
    field_modulus = field_modulus


class bn128_FQP(FQP):
    field_modulus = field_modulus


class bn128_FQ2(FQ2):
    field_modulus = field_modulus
    FQ2_MODULUS_COEFFS = FQ2_MODULUS_COEFFS


class bn128_FQ12(FQ12):
    field_modulus = field_modulus
    FQ12_MODULUS_COEFFS = FQ12_MODULUS_COEFFS
