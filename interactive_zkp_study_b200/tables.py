"""Device-resident copies of the static point tables the reference re-passes on every call
(``srs.g1_powers``, ``sigma1_2``, ``sigma1_5``, ``sigma2_2``; SURVEY.md 8b "Ownership").

Cached by object identity, guarded by a content fingerprint (length + sampled points) so that a
recycled ``id()`` or an in-place edit can never serve stale points.
"""
from collections import OrderedDict

from . import native

_MAX_ENTRIES = 32
_cache = OrderedDict()


def _fingerprint(points, g2):
    n = len(points)
    idx = sorted(set([0, 1, n // 3, n // 2, (2 * n) // 3, n - 2, n - 1]) & set(range(n)))
    enc = native.g2_bytes if g2 else native.g1_bytes
    return (n, tuple(enc(points[i]) for i in idx))


def _get(points, g2):
    key = (id(points), bool(g2))
    fp = _fingerprint(points, g2)
    hit = _cache.get(key)
    if hit is not None and hit[0] == fp:
        _cache.move_to_end(key)
        return hit[1]
    n = len(points)
    if g2:
        handle = native.g2_table_load(native.g2_vec_bytes(points), n)
    else:
        handle = native.g1_table_load(native.g1_vec_bytes(points), n)
    _maybe_precompute(handle, n)
    _cache[key] = (fp, handle)
    while len(_cache) > _MAX_ENTRIES:
        _, (_, old) = _cache.popitem(last=False)
        old.free()
    return handle


PRECOMPUTE_MIN_POINTS = 2


def _maybe_precompute(handle, n):
    """Cached tables are static (SRS / CRS), so they all get the window-precomputed layout: an MSM on
    it has no 254-step doubling chain, which is most of the latency of a small commit (a toy-size
    commit drops from ~1.5 ms to ~0.3 ms) and ~25 % of a 2^20 one.  Cost: ceil(255/c) x the memory."""
    if n >= PRECOMPUTE_MIN_POINTS:
        native.table_precompute(handle)


def g1_table(points):
    return _get(points, False)


def g2_table(points):
    return _get(points, True)


def clear():
    while _cache:
        _, (_, h) = _cache.popitem()
        h.free()


def adopt_g1(points, handle):
    """Register an already device-resident table (e.g. fresh from SRS.generate) for `points`."""
    _maybe_precompute(handle, len(points))
    _cache[(id(points), False)] = (_fingerprint(points, False), handle)
