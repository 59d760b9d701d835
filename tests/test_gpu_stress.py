"""-m gpu: the library under concurrent callers (the mutex that serialises entry points: Flask's dev
server is threaded, ctypes drops the GIL -- SURVEY H9) and under repetition (no device-memory growth
across many proofs: pooled handle buffers, arena reuse, bounded caches)."""
import os
import random
import sys
import threading

import pytest

from oracle import bn254, ref_path

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu
R = bn254.R


def test_concurrent_callers_get_their_own_results(native):
    rng = random.Random(1)
    base = bn254.g1_mul(bn254.G1, 12345)
    pts, acc = [], base
    for _ in range(64):
        pts.append(acc)
        acc = bn254.g1_add(acc, base)
    pts_b = native.g1_vec_bytes(pts)
    jobs = []
    for k in range(8):
        sc = [rng.randrange(R) for _ in range(64)]
        vals = [rng.randrange(R) for _ in range(256)]
        jobs.append((sc, bn254.g1_msm(pts, sc), vals, ref_path.fft(vals, ref_path.get_root_of_unity(256))))
    errors = []

    def worker(job):
        sc, want_pt, vals, want_ev = job
        try:
            for _ in range(10):
                assert native.g1_msm(pts_b, native.fr_vec_bytes(sc), 64) == want_pt
                got = native.fr_vec_from_bytes(native.fr_ntt(native.fr_vec_bytes(vals), 8, ref_path.get_root_of_unity(256)))
                assert got == want_ev
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(j,)) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:2]


def test_repeated_proofs_do_not_grow_device_memory(native):
    from interactive_zkp_study_b200.zkp.plonk import device_prover as dp
    from interactive_zkp_study_b200.zkp.groth16 import device_prover as g16
    import plonk_synth
    key, wit, _ = plonk_synth.device_setup(plonk_synth.chain_circuit(1 << 10, seed=3))
    rng = random.Random(5)
    k = 1 << 10
    gkey = g16.setup_from_toxic(k, 11, 22, 33, 44, 55, [rng.randrange(R) for _ in range(k - 2)], precompute=False)
    uA, uB, uC = (native.scalars_generate(100 + i, k) for i in range(3))
    Z = native.scalars_load(native.fr_vec_bytes([rng.randrange(R) for _ in range(k)] + [1]), k + 1)
    rx = native.scalars_generate(200, k - 2)
    for _ in range(5):                                   # warm every pool / cache
        dp.prove(key, *wit)
        g16.prove(gkey, uA, uB, uC, Z, rx, 3, 4)
    native.sync()
    free0, _ = native.device_mem_info()
    for _ in range(60):
        dp.prove(key, *wit)
        g16.prove(gkey, uA, uB, uC, Z, rx, 3, 4)
    native.sync()
    free1, _ = native.device_mem_info()
    assert free0 - free1 < (64 << 20), "device memory shrank by %d MiB over 60 proofs" % ((free0 - free1) >> 20)
