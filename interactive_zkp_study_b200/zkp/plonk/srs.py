"""Drop-in for /root/reference/zkp/plonk/srs.py: SRS container and generate().

generate (:50-87) is max_degree+1 sequential fixed-base scalar multiplications in the reference;
here the powers of tau are one device prefix product and [tau^i]_1 one batched fixed-base
multiplication kernel.  The resulting g1_powers table stays resident on the device
and is registered with the table cache, so the first commit() does not upload it again.
"""
import hashlib

from ... import native, tables
from ...compat import g1_from_ints, g2_from_ints
from .field import FR, G1, G2, CURVE_ORDER


class SRS:
    def __init__(self, g1_powers, g2_powers, max_degree):
        self.g1_powers = g1_powers
        self.g2_powers = g2_powers
        self.max_degree = max_degree

    @classmethod
    def generate(cls, max_degree, seed=None):
        if seed is not None:
            h = hashlib.sha256(str(seed).encode()).digest()
            tau_int = int.from_bytes(h, "big") % CURVE_ORDER
        else:
            import secrets
            tau_int = secrets.randbelow(CURVE_ORDER - 1) + 1
        n = max_degree + 1
        powers = native.fr_prefix_product(native.fr_vec_bytes([tau_int] * n), n)  # [1, tau, ..., tau^d]
        handle = native.g1_fixed_base_mul(native.g1_bytes(G1), powers, n)
        raw = native.table_download(handle, 0, n)
        g1_powers = [g1_from_ints(native.g1_from_bytes(raw[64 * i:64 * i + 64])) for i in range(n)]
        h2 = native.g2_fixed_base_mul(native.g2_bytes(G2), native.fe_bytes(tau_int), 1)
        tau_g2 = g2_from_ints(native.g2_from_bytes(native.table_download(h2, 0, 1)))
        h2.free()
        tables.adopt_g1(g1_powers, handle)
        return cls(g1_powers, [G2, tau_g2], max_degree)
