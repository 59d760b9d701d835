"""Binary wire format for the PLONK objects.

The reference moves every prover object between its Flask routes and TinyDB as JSON with one DECIMAL
STRING per field element (/root/reference/plonk_serializers.py:23-250; round trips per round in
plonk_routes.py:298-373).  At 2^20 gates that is ~80 bytes and two bignum <-> decimal conversions per
coefficient per round.  This module is the same set of functions -- same names, same objects in and
out -- over a binary encoding that is the C ABI's own layout, so a polynomial that lives on the GPU is
written and read without ever becoming Python integers:

    Fr      32 bytes little-endian, canonical
    G1      x | y, 64 bytes; the point at infinity (None) is 64 zero bytes
    G2      x.c0 | x.c1 | y.c0 | y.c1, 128 bytes; None is 128 zero bytes
    vectors the elements back to back

Composite objects (SRS, PreprocessedData, Proof) are a flat container: magic, then (name, tag, length,
payload) records.  `to_text` / `from_text` wrap any payload as base64 for stores that only take strings
(TinyDB, JSON responses).
"""
import base64
import struct

from ... import native
from ...compat import FQ, FQ2, curve_order
from .field import FR
from .polynomial import Polynomial
from .preprocessor import PreprocessedData
from .prover import Proof
from .srs import SRS
from .transcript import Transcript

MAGIC = b"ZKPB\x01"
T_NONE, T_FR, T_G1, T_G2, T_FRVEC, T_INT, T_U32VEC, T_G1VEC, T_G2VEC, T_BYTES = range(10)


# ------------------------------------------------------------------ scalars and points
def serialize_fr(val):
    return (int(val) % curve_order).to_bytes(32, "little")


def deserialize_fr(b):
    return FR(int.from_bytes(b, "little"))


def serialize_g1(point):
    return native.g1_bytes(point)


def deserialize_g1(b):
    p = native.g1_from_bytes(bytes(b))
    return None if p is None else (FQ(p[0]), FQ(p[1]))


def serialize_g2(point):
    return native.g2_bytes(point)


def deserialize_g2(b):
    p = native.g2_from_bytes(bytes(b))
    return None if p is None else (FQ2([p[0][0], p[0][1]]), FQ2([p[1][0], p[1][1]]))


def serialize_fr_list(lst):
    return native.fr_vec_bytes(lst)


def deserialize_fr_list(b):
    return [FR(v) for v in native.fr_vec_from_bytes(bytes(b))]


def serialize_poly(poly):
    """Polynomial -> coefficient bytes (None stays None, as in the reference)."""
    return None if poly is None else native.fr_vec_bytes(poly.coeffs)


def deserialize_poly(b):
    return None if b is None else Polynomial(deserialize_fr_list(b))


# device-resident vectors: no Python integers on the way
def serialize_handle(handle, n=None, offset=0):
    """Device scalar vector (canonical form) -> the same bytes serialize_fr_list would give."""
    return native.scalars_download(handle, offset, handle.n - offset if n is None else n)


def deserialize_to_handle(b):
    return native.scalars_load(bytes(b), len(b) // 32)


def serialize_transcript(transcript):
    return bytes(transcript.state)


def deserialize_transcript(b):
    t = Transcript.__new__(Transcript)
    t.state = bytearray(b)
    return t


# ------------------------------------------------------------------ container
def dumps(fields):
    """fields: iterable of (name, tag, payload) -> one blob."""
    out = [MAGIC]
    for name, tag, payload in fields:
        nb = name.encode("ascii")
        payload = b"" if payload is None else payload
        out.append(struct.pack("<BBQ", len(nb), tag, len(payload)))
        out.append(nb)
        out.append(payload)
    return b"".join(out)


def loads(blob):
    """-> {name: (tag, payload)}; raises ValueError on a foreign or truncated blob."""
    blob = bytes(blob)
    if blob[:len(MAGIC)] != MAGIC:
        raise ValueError("not a zkp-b200 wire blob")
    pos, out = len(MAGIC), {}
    while pos < len(blob):
        if pos + 10 > len(blob):
            raise ValueError("truncated wire blob")
        nlen, tag, plen = struct.unpack_from("<BBQ", blob, pos)
        pos += 10
        if pos + nlen + plen > len(blob):
            raise ValueError("truncated wire blob")
        name = blob[pos:pos + nlen].decode("ascii")
        pos += nlen
        out[name] = (tag, blob[pos:pos + plen])
        pos += plen
    return out


def _g1f(name, p):
    return (name, T_NONE, None) if p is None else (name, T_G1, serialize_g1(p))


def _frf(name, v):
    return (name, T_NONE, None) if v is None else (name, T_FR, serialize_fr(v))


def _get_g1(rec, name):
    tag, payload = rec.get(name, (T_NONE, b""))
    return None if tag == T_NONE else deserialize_g1(payload)


def _get_fr(rec, name):
    tag, payload = rec.get(name, (T_NONE, b""))
    return None if tag == T_NONE else deserialize_fr(payload)


def to_text(blob):
    return base64.b64encode(blob).decode("ascii")


def from_text(text):
    return base64.b64decode(text.encode("ascii"))


# ------------------------------------------------------------------ SRS
def serialize_srs(srs):
    return dumps([("g1_powers", T_G1VEC, native.g1_vec_bytes(srs.g1_powers)),
                  ("g2_powers", T_G2VEC, native.g2_vec_bytes(srs.g2_powers)),
                  ("max_degree", T_INT, struct.pack("<q", srs.max_degree))])


def deserialize_srs(blob):
    rec = loads(blob)
    g1b, g2b = rec["g1_powers"][1], rec["g2_powers"][1]
    g1 = [deserialize_g1(g1b[i:i + 64]) for i in range(0, len(g1b), 64)]
    g2 = [deserialize_g2(g2b[i:i + 128]) for i in range(0, len(g2b), 128)]
    return SRS(g1, g2, struct.unpack("<q", rec["max_degree"][1])[0])


# ------------------------------------------------------------------ PreprocessedData
_PP_POLYS = ("q_l", "q_r", "q_o", "q_m", "q_c", "s_sigma1", "s_sigma2", "s_sigma3")


def serialize_preprocessed(pp):
    fields = [("n", T_INT, struct.pack("<q", pp.n)), ("omega", T_FR, serialize_fr(pp.omega)),
              ("domain", T_FRVEC, serialize_fr_list(pp.domain))]
    for name in _PP_POLYS:
        fields.append((name + "_poly", T_FRVEC, serialize_poly(getattr(pp, name + "_poly"))))
        fields.append(_g1f(name + "_comm", getattr(pp, name + "_comm")))
    fields.append(("sigma", T_U32VEC, struct.pack("<%dI" % len(pp.sigma), *pp.sigma)))
    fields.append(("num_public_inputs", T_INT, struct.pack("<q", pp.num_public_inputs)))
    return dumps(fields)


def deserialize_preprocessed(blob):
    rec = loads(blob)
    pp = PreprocessedData()
    pp.n = struct.unpack("<q", rec["n"][1])[0]
    pp.omega = deserialize_fr(rec["omega"][1])
    pp.domain = deserialize_fr_list(rec["domain"][1])
    for name in _PP_POLYS:
        setattr(pp, name + "_poly", deserialize_poly(rec[name + "_poly"][1]))
        setattr(pp, name + "_comm", _get_g1(rec, name + "_comm"))
    sb = rec["sigma"][1]
    pp.sigma = list(struct.unpack("<%dI" % (len(sb) // 4), sb))
    pp.num_public_inputs = struct.unpack("<q", rec["num_public_inputs"][1])[0]
    return pp


# ------------------------------------------------------------------ Proof
_PROOF_G1 = ("a_comm", "b_comm", "c_comm", "z_comm", "t_lo_comm", "t_mid_comm", "t_hi_comm", "W_zeta_comm",
             "W_zeta_omega_comm")
_PROOF_FR = ("a_eval", "b_eval", "c_eval", "s_sigma1_eval", "s_sigma2_eval", "z_omega_eval", "r_eval")


def serialize_proof(proof):
    """9 G1 points + 7 field elements: 800 bytes of payload for a complete proof."""
    return dumps([_g1f(n, getattr(proof, n)) for n in _PROOF_G1] + [_frf(n, getattr(proof, n)) for n in _PROOF_FR])


def deserialize_proof(blob):
    rec = loads(blob)
    proof = Proof()
    for n in _PROOF_G1:
        setattr(proof, n, _get_g1(rec, n))
    for n in _PROOF_FR:
        setattr(proof, n, _get_fr(rec, n))
    return proof
