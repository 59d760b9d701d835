"""Fiat-Shamir transcript with the byte layout of /root/reference/zkp/plonk/transcript.py:47-123
(SHA-256 over an append-only state; scalars and coordinates as 32-byte big-endian; infinity as 64
zero bytes; every challenge digest is appended back).  Sequential hashing is outside the GPU hot
path; it lives here only so the prover mirror is self-contained -- in a deployment
the reference's own module is used unchanged and produces the same bytes."""
import hashlib

from .field import FR, CURVE_ORDER


class Transcript:
    def __init__(self, label=b"plonk"):
        self.state = bytearray(label)

    def append_scalar(self, label, scalar):
        self.state += label
        self.state += (int(scalar) % CURVE_ORDER).to_bytes(32, "big")

    def append_point(self, label, point):
        self.state += label
        if point is None:
            self.state += bytes(64)
        else:
            self.state += int(point[0]).to_bytes(32, "big") + int(point[1]).to_bytes(32, "big")

    def challenge_scalar(self, label):
        self.state += label
        digest = hashlib.sha256(bytes(self.state)).digest()
        self.state += digest
        return FR(int.from_bytes(digest, "big") % CURVE_ORDER)
