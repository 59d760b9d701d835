"""Model of a Montgomery product on the FP64 pipe (the one unit of the SM this path leaves idle).
Exact-integer emulation of what the double-precision instructions would compute -- fma with round-toward-zero on
operands that are integers below 2^52 -- so the ALGORITHM (52-bit limbs, Montgomery radix 2^260, high / low halves of
every limb product from two FMAs, integer accumulation of the IEEE bit patterns) is checked against a b R^-1 mod p
on the CPU before any kernel is written.  Nothing in the product uses this; it prints the instruction budget the
design note quotes.   usage: python tools/dp_montgomery_model.py [trials]"""
import random
import sys

P = 21888242871839275222246405745257275088696311157297823662689037894645226208583   # BN254 base field
LIMBS, W = 5, 52
MASK = (1 << W) - 1
R = 1 << (LIMBS * W)
N_PRIME = (-pow(P, -1, 1 << W)) % (1 << W)        # -p^-1 mod 2^52
C1 = 1 << (2 * W)                                  # 2^104: forces the exponent, ulp = 2^52
C2 = C1 + (1 << W)
counts = {"dfma": 0, "dadd": 0, "iadd64": 0, "logic": 0}


def rz53(x):
    """x rounded toward zero to 53 significant bits (what a round-toward-zero FP64 result holds)."""
    if x == 0:
        return 0
    s, m = (1, x) if x > 0 else (-1, -x)
    shift = max(0, m.bit_length() - 53)
    return s * ((m >> shift) << shift)


def fma_rz(a, b, c):
    counts["dfma"] += 1
    return rz53(a * b + c)


def bits_of(x, exponent):
    """IEEE-754 pattern of the double x = 2^exponent + k 2^(exponent-52), as the integer pipe sees it."""
    k = (x - (1 << exponent)) >> (exponent - 52)
    assert 0 <= k <= MASK and x == (1 << exponent) + (k << (exponent - 52))
    return ((1023 + exponent) << 52) | k


def limb_product(a, b):
    """(high, low) 52-bit halves of a * b as the two bit patterns an FMA pair leaves in registers."""
    hi = fma_rz(a, b, C1)                  # 2^104 + floor(a b / 2^52) 2^52: the low half is truncated away
    counts["dadd"] += 1
    lo = fma_rz(a, b, C2 - hi)             # 2^52 + (a b mod 2^52), exact
    return bits_of(hi, 104), bits_of(lo, 52)


BIAS_HI, BIAS_LO = (1023 + 104) << 52, (1023 + 52) << 52


def to_limbs(x):
    return [(x >> (W * i)) & MASK for i in range(LIMBS)]


def mont_mul(a, b):
    """a b 2^-260 mod p on limbs; columns are 64-bit integer sums of bit patterns, biases removed per column."""
    al, bl = to_limbs(a), to_limbs(b)
    pl = to_limbs(P)
    col = [0] * (2 * LIMBS + 1)
    terms_hi = [0] * (2 * LIMBS + 1)
    terms_lo = [0] * (2 * LIMBS + 1)

    def mac(x, y, k):
        h, l = limb_product(x, y)
        col[k + 1] += h
        col[k] += l
        terms_hi[k + 1] += 1
        terms_lo[k] += 1
        counts["iadd64"] += 2

    for i in range(LIMBS):
        for j in range(LIMBS):
            mac(al[i], bl[j], i + j)
    for i in range(LIMBS):
        # settle column i: remove the biases of the terms it has received, take its low 52 bits
        col[i] -= terms_hi[i] * BIAS_HI + terms_lo[i] * BIAS_LO
        terms_hi[i] = terms_lo[i] = 0
        counts["iadd64"] += 1
        t = col[i] & MASK
        counts["logic"] += 2                # mask, or-into-exponent (the int -> double conversion is one more dadd)
        counts["dadd"] += 1
        _, ql = limb_product(t, N_PRIME)    # q = t n' mod 2^52
        q = ql - BIAS_LO
        counts["iadd64"] += 1
        for j in range(LIMBS):
            mac(q, pl[j], i + j)
        col[i] -= terms_lo[i] * BIAS_LO     # only low halves land in column i at this point
        terms_lo[i] = 0
        assert col[i] & MASK == 0
        col[i + 1] += col[i] >> W           # carry into the next column
        counts["iadd64"] += 2
        counts["logic"] += 1
    out = 0
    for k in range(LIMBS, 2 * LIMBS + 1):
        col[k] -= terms_hi[k] * BIAS_HI + terms_lo[k] * BIAS_LO
        counts["iadd64"] += 1
        out += col[k] << (W * (k - LIMBS))  # carry propagation + conditional subtraction: ~25 integer instructions
    assert out < 2 * P
    return out - P if out >= P else out


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    rng = random.Random(5)
    edge = [0, 1, 2, P - 1, P - 2, (1 << 52) - 1, 1 << 52, (1 << 253), P >> 1, MASK << 208]
    cases = [(a % P, b % P) for a in edge for b in edge] + [(rng.randrange(P), rng.randrange(P)) for _ in range(trials)]
    r_inv = pow(R, -1, P)
    for k in counts:
        counts[k] = 0
    for a, b in cases:
        assert mont_mul(a, b) == a * b * r_inv % P, (a, b)
    n = len(cases)
    print("%d products agree with a b 2^-260 mod p" % n)
    print("per product: %d DFMA + %d DADD on the FP64 pipe, %d 64-bit integer additions, %d logic ops (+ ~25 for the final "
          "carry / conditional subtraction)" % tuple(counts[k] // n for k in ("dfma", "dadd", "iadd64", "logic")))
    print("the integer product it would run beside: 120 IMAD.WIDE + 8 IMAD.HI + 8 IMAD, 31 IADD3 (profiles/r2_sass_instruction_mix.txt)")


if __name__ == "__main__":
    main()
