"""TEST INFRASTRUCTURE (oracle) -- restatement of the device-side synthetic scalar stream
(zkp_scalars_generate, include/zkp_b200.h): SplitMix64-finaliser counter PRNG,
254-bit draws, first of 16 that is < r, else the 16th minus r.  numpy-vectorised."""
import numpy as np

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_M = (1 << 64) - 1


def mix64(z):
    z = (z + 0x9E3779B97F4A7C15) & _M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
    return z ^ (z >> 31)


def _mix64_np(z):
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def scalar(seed, i):
    v = 0
    for t in range(16):
        w = [mix64((seed + 64 * i + 4 * t + j) & _M) for j in range(4)]
        v = w[0] | (w[1] << 64) | (w[2] << 128) | ((w[3] & 0x3FFFFFFFFFFFFFFF) << 192)
        if v < R:
            return v
    return v - R


def scalars(seed, n):
    """list of n Python ints, identical to the device stream."""
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        out = [None] * n
        pending = np.arange(n)
        last = None
        for t in range(16):
            if len(pending) == 0:
                break
            base = np.uint64(seed) + np.uint64(64) * idx[pending] + np.uint64(4 * t)
            w = [_mix64_np(base + np.uint64(j)) for j in range(4)]
            w[3] = w[3] & np.uint64(0x3FFFFFFFFFFFFFFF)
            vals = [int(a) | (int(b) << 64) | (int(c) << 128) | (int(d) << 192)
                    for a, b, c, d in zip(w[0].tolist(), w[1].tolist(), w[2].tolist(), w[3].tolist())]
            keep = []
            for k, v in zip(pending.tolist(), vals):
                if v < R:
                    out[k] = v
                else:
                    keep.append(k)
                    last = (k, v)
                    if t == 15:
                        out[k] = v - R
            pending = np.array(keep, dtype=np.int64)
        return out
