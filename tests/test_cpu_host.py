"""CPU: host-side logic of the drop-in layer that needs no device -- byte encodings, the compat
field types, Polynomial trimming/degree, error behaviour, the multi-rank shard arithmetic (gloo)."""
import os
import socket

import pytest

from interactive_zkp_study_b200 import native
from interactive_zkp_study_b200.compat import FQ, FQ2, FR, curve_order, field_modulus


def test_encodings_round_trip():
    assert native.fe_bytes(1) == b"\x01" + bytes(31)
    v = [0, 1, curve_order - 1, 1 << 200]
    assert native.fr_vec_from_bytes(native.fr_vec_bytes(v)) == v
    p = (FQ(1), FQ(2))
    assert native.g1_from_bytes(native.g1_bytes(p)) == (1, 2)
    assert native.g1_bytes(None) == bytes(64) and native.g1_from_bytes(bytes(64)) is None
    q = (FQ2([1, 2]), FQ2([3, 4]))
    assert native.g2_from_bytes(native.g2_bytes(q)) == ((1, 2), (3, 4))
    assert native.g2_bytes(((1, 2), (3, 4))) == native.g2_bytes(q)


def test_compat_field_semantics_match_py_ecc():
    a, b = FR(5), FR(curve_order - 2)
    assert int(a + b) == 3 and int(a - b) == 7 and int(a * 3) == 15 and int(3 * a) == 15
    assert int(FR(-1)) == curve_order - 1 and int(-a) == curve_order - 5
    assert a / a == FR(1) and int(FR(0) / FR(0)) == 0       # inv(0) == 0
    assert int(FR(1) / a * a) == 1 and int(2 / FR(2)) == 1
    assert a ** 0 == FR(1) and int(a ** 3) == 125 and a == 5 and a != 6
    assert int(FQ(field_modulus + 4)) == 4 and isinstance(a + 1, FR) and not isinstance(FQ(1) + 1, FR)
    with pytest.raises(TypeError):
        a == "5"
    with pytest.raises(TypeError):
        FR(1.5)


def test_polynomial_host_behaviour():
    from interactive_zkp_study_b200.zkp.plonk.polynomial import Polynomial
    from interactive_zkp_study_b200.zkp.plonk.field import get_root_of_unity
    p = Polynomial([1, 2, 0, 0])
    assert [int(c) for c in p.coeffs] == [1, 2] and p.degree == 1 and len(p) == 2 and not p.is_zero()
    assert Polynomial().is_zero() and Polynomial([0, 0, 0]).degree == 0 and Polynomial.zero() == Polynomial([0])
    assert Polynomial([3]) == 3 and Polynomial([3]) == FR(3) and (Polynomial([1]) == "x") is False
    assert repr(Polynomial([1, 0, 3])) == "Poly(1 + 3*x^2)" and repr(Polynomial.zero()) == "Poly(0)"
    assert [int(c) for c in Polynomial.vanishing(4).coeffs] == [curve_order - 1, 0, 0, 0, 1]
    assert int(get_root_of_unity(1)) == 1
    assert int(get_root_of_unity(4)) == 21888242871839275217838484774961031246007050428528088939761107053157389710902
    for bad in (0, 3, 12, (1 << 28) + 1):
        with pytest.raises(ValueError):
            get_root_of_unity(bad)


def test_transcript_bytes_follow_the_reference_layout():
    import hashlib
    from interactive_zkp_study_b200.zkp.plonk.transcript import Transcript
    t = Transcript()
    t.append_point(b"a_comm", (FQ(1), FQ(2)))
    t.append_point(b"b_comm", None)
    t.append_scalar(b"x", FR(7))
    state = b"plonk" + b"a_comm" + (1).to_bytes(32, "big") + (2).to_bytes(32, "big") + b"b_comm" + bytes(64) + \
        b"x" + (7).to_bytes(32, "big") + b"beta"
    want = int.from_bytes(hashlib.sha256(state).digest(), "big") % curve_order
    assert int(t.challenge_scalar(b"beta")) == want
    assert bytes(t.state) == state + hashlib.sha256(state).digest()


def test_commit_degree_check_precedes_any_device_work():
    from interactive_zkp_study_b200.zkp.plonk.polynomial import Polynomial
    from interactive_zkp_study_b200.zkp.plonk.kzg import commit
    from interactive_zkp_study_b200.zkp.plonk.srs import SRS
    srs = SRS([(FQ(1), FQ(2))] * 3, [None, None], 2)
    with pytest.raises(ValueError, match="SRS"):
        commit(Polynomial([1, 2, 3, 4]), srs)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


_R256 = 1 << 256


def _encode_partial(pt, scale=1):
    """Oracle-side encoder of the 128-byte shard partial (sharded.py wire format): XYZZ Montgomery,
    un-normalised on purpose (zz = scale^2, zzz = scale^3) like a real GPU partial."""
    from oracle import bn254
    P = bn254.P
    if pt is None:
        return bytes(128)
    zz, zzz = pow(scale, 2, P), pow(scale, 3, P)
    coords = (pt[0] * zz % P, pt[1] * zzz % P, zz, zzz)
    return b"".join((c * _R256 % P).to_bytes(32, "little") for c in coords)


def _decode_partial(b):
    from oracle import bn254
    P = bn254.P
    rinv = pow(_R256, -1, P)
    x, y, zz, zzz = (int.from_bytes(b[32 * i:32 * i + 32], "little") * rinv % P for i in range(4))
    if zz == 0:
        return None
    return (x * pow(zz, -1, P) % P, y * pow(zzz, -1, P) % P)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle import bn254, synthetic
    from interactive_zkp_study_b200 import sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the product's bootstrap channel: rank 0's 128-byte communicator id reaches every rank
    ident = bytes(range(128)) if rank == 0 else None
    got_id = sharded.tcp_broadcast(ident, rank, world, "127.0.0.1", port + 1)
    total = 49                               # odd on purpose: ranks own 25 and 24 points
    start, count = sharded.shard_range(total, rank, world)
    s = synthetic.scalars(0x5EED0002, total)[start:start + count]
    k = synthetic.scalars(0x5EED0001, total)[start:start + count]
    pts = [bn254.g1_mul(bn254.G1, x) for x in s]
    part = bn254.g1_msm(pts, k)             # the per-rank partial sum (CPU stand-in for the GPU shard)
    # the exchange step of the sharded MSM (ncclAllGather inside the library on a GPU box), here over gloo
    # with oracle-made partials in the library's 128-byte wire format
    send = torch.frombuffer(bytearray(_encode_partial(part, scale=7 + rank)), dtype=torch.uint8)
    recv = torch.zeros(sharded.PARTIAL_BYTES * world, dtype=torch.uint8)
    dist.all_gather_into_tensor(recv, send)
    blob = bytes(recv.numpy().tobytes())
    if rank == 0:
        acc = None
        for r in range(world):
            acc = bn254.g1_add(acc, _decode_partial(blob[128 * r:128 * r + 128]))
        s_all = synthetic.scalars(0x5EED0002, total)
        k_all = synthetic.scalars(0x5EED0001, total)
        want = bn254.g1_mul(bn254.G1, sum(a * b for a, b in zip(s_all, k_all)) % bn254.R)
        q.put(acc == want and got_id == bytes(range(128)))
    else:
        assert got_id == bytes(range(128))
    dist.barrier()
    dist.destroy_process_group()


def test_tcp_rendezvous_world1_and_bad_arguments():
    from interactive_zkp_study_b200 import sharded
    assert sharded.tcp_broadcast(b"abc", 0, 1) == b"abc"          # a single rank needs no socket
    with pytest.raises(ValueError):
        sharded.Communicator(2, 2)
    with pytest.raises(ValueError):
        sharded.g1_msm_sharded(None, None, None, 0)


def test_shard_range_tiles_the_index_space():
    from interactive_zkp_study_b200.sharded import shard_range
    for total in (0, 1, 7, 8, 49, 1 << 20, (1 << 26) + 3):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            for r in range(world):
                start, count = shard_range(total, r, world)
                assert start == nxt and count in (total // world, total // world + 1)
                nxt = start + count
            assert nxt == total
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_sharded_msm_partition_and_combine_world2_gloo():
    """N>1 path on CPU (gloo, world_size 2): the product's shard_range and communicator-id rendezvous,
    and the one exchange step of the sharded MSM (an all-gather of 128-byte partials; inside the library
    it is ncclAllGather) with oracle-made partials in the wire format; the fold == the unsharded MSM."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert ok is True


def test_plonk_scalar_helpers_match_reference_formulas():
    """utils.py scalar helpers (reference zkp/plonk/utils.py:25-81,208-246): closed forms on the host."""
    from interactive_zkp_study_b200.zkp.plonk import utils
    from interactive_zkp_study_b200.zkp.plonk.field import get_root_of_unity
    from oracle import ref_path
    R = curve_order
    assert [utils.next_power_of_2(n) for n in (0, 1, 2, 3, 4, 5, 8, 9, 1000)] == [1, 1, 2, 4, 4, 8, 8, 16, 1024]
    assert utils.pad_to_power_of_2([FR(1), FR(2), FR(3)]) == [FR(1), FR(2), FR(3), FR(0)]
    assert utils.pad_to_power_of_2([1, 2, 3, 4, 5], fill=7) == [1, 2, 3, 4, 5, 7, 7, 7]
    n = 8
    w = get_root_of_unity(n)
    zeta = FR(123456789)
    assert int(utils.vanishing_poly_eval(n, zeta)) == (pow(123456789, n, R) - 1) % R
    dom = ref_path.get_roots_of_unity(n)
    for i in range(n):
        # L_i(zeta) by direct Lagrange product
        num = den = 1
        for j in range(n):
            if j != i:
                num = num * (123456789 - dom[j]) % R
                den = den * (dom[i] - dom[j]) % R
        assert int(utils.lagrange_basis_eval(i, n, w, zeta)) == num * pow(den, -1, R) % R
        assert utils.lagrange_basis_eval(i, n, w, FR(dom[i])) == FR(1)
    pub = [FR(5), FR(7)]
    want = sum(int(v) * int(utils.lagrange_basis_eval(i, n, w, zeta)) for i, v in enumerate(pub)) % R
    assert int(utils.public_input_poly_eval(pub, n, w, zeta)) == want
    assert utils.public_input_polynomial([], n, w).is_zero()


def test_binary_wire_format_round_trips_the_reference_objects():
    """zkp/plonk/wire.py: the reference's serializer surface (plonk_serializers.py:23-250) over
    the binary ABI layout -- round trips of the reference-minted proof / preprocessed data / SRS carry exactly
    the values its decimal-string JSON carries."""
    from interactive_zkp_study_b200.zkp.plonk import wire
    from interactive_zkp_study_b200.zkp.plonk.polynomial import Polynomial
    from interactive_zkp_study_b200.zkp.plonk.preprocessor import PreprocessedData
    from interactive_zkp_study_b200.zkp.plonk.prover import Proof
    from interactive_zkp_study_b200.zkp.plonk.srs import SRS
    from interactive_zkp_study_b200.zkp.plonk.transcript import Transcript
    from tests.util import load
    f = load("plonk_chain16.json")
    pt = lambda p: None if p is None else (FQ(int(p[0])), FQ(int(p[1])))
    pt2 = lambda p: (FQ2([int(p[0][0]), int(p[0][1])]), FQ2([int(p[1][0]), int(p[1][1])]))
    # scalars, points, infinity
    assert int(wire.deserialize_fr(wire.serialize_fr(FR(curve_order - 1)))) == curve_order - 1
    assert wire.deserialize_g1(wire.serialize_g1(None)) is None and len(wire.serialize_g1(None)) == 64
    assert wire.deserialize_g2(wire.serialize_g2(None)) is None
    g = pt(f["proof"]["a_comm"])
    assert wire.deserialize_g1(wire.serialize_g1(g)) == g
    # proof: 9 points + 7 scalars
    proof = Proof()
    for k, v in f["proof"].items():
        setattr(proof, k, pt(v) if k.endswith("_comm") else FR(int(v)))
    blob = wire.serialize_proof(proof)
    back = wire.deserialize_proof(wire.from_text(wire.to_text(blob)))
    for k, v in f["proof"].items():
        got = getattr(back, k)
        assert (got == pt(v)) if k.endswith("_comm") else (int(got) == int(v)), k
    assert len(blob) < 1200                                  # vs ~2.6 KB of decimal JSON
    partial = Proof()
    partial.a_comm = g                                       # rounds not run yet stay None (plonk_routes stores partial proofs)
    back = wire.deserialize_proof(wire.serialize_proof(partial))
    assert back.a_comm == g and back.z_comm is None and back.r_eval is None
    # preprocessed data
    pp = PreprocessedData()
    pp.n, pp.omega, pp.domain = f["n"], FR(int(f["omega"])), [FR(int(v)) for v in f["domain"]]
    for k in f["pre"]:
        setattr(pp, k + "_poly", Polynomial([FR(int(v)) for v in f["pre"][k]]))
        setattr(pp, k + "_comm", pt(f["pre_comm"][k]))
    pp.sigma, pp.num_public_inputs = f["sigma"], f["num_public_inputs"]
    back = wire.deserialize_preprocessed(wire.serialize_preprocessed(pp))
    assert (back.n, int(back.omega), back.sigma, back.num_public_inputs) == (pp.n, int(pp.omega), pp.sigma, pp.num_public_inputs)
    assert [int(v) for v in back.domain] == [int(v) for v in pp.domain]
    for k in f["pre"]:
        assert [int(c) for c in getattr(back, k + "_poly").coeffs] == [int(v) for v in f["pre"][k]]
        assert getattr(back, k + "_comm") == pt(f["pre_comm"][k])
    # SRS and transcript
    srs = SRS([pt(p) for p in f["g1_powers"]], [pt2(p) for p in f["g2_powers"]], f["srs_max_degree"])
    back = wire.deserialize_srs(wire.serialize_srs(srs))
    assert back.max_degree == srs.max_degree and back.g1_powers == srs.g1_powers and back.g2_powers == srs.g2_powers
    t = Transcript()
    t.append_scalar(b"x", FR(5)) if hasattr(t, "append_scalar") else None
    t2 = wire.deserialize_transcript(wire.serialize_transcript(t))
    assert bytes(t2.state) == bytes(t.state)
    assert wire.deserialize_poly(wire.serialize_poly(None)) is None
    with pytest.raises(ValueError):
        wire.loads(b"nope")
    with pytest.raises(ValueError):
        wire.loads(blob[:-5])


def test_sparse_r1cs_host_logic():
    """qap_device host side: CSR construction from the reference's dense R1CS, the transpose
    used for the CRS terms, and the determinant factor of r1cs_to_qap_times_lcm -- no device needed."""
    import numpy as np
    from interactive_zkp_study_b200.zkp.groth16 import qap_device as qd
    from oracle import ref_path
    from tests.util import load
    for c in load("groth16_qap.json")["cases"]:
        A, B, C = c["r1cs_A"], c["r1cs_B"], c["r1cs_C"]
        r1cs = qd.SparseR1CS.from_dense(A, B, C)
        k, m = c["numGates"], c["numWires"]
        assert (r1cs.k, r1cs.m) == (k, m)
        for dense, (rp, ci, vals), which in zip((A, B, C), r1cs.mats, range(3)):
            back = [[0] * m for _ in range(k)]
            for g in range(k):
                for e in range(int(rp[g]), int(rp[g + 1])):
                    back[g][int(ci[e])] = vals[e]
            assert back == [[int(v) % curve_order for v in row] for row in dense]
            t_rp, t_ci, t_vals = r1cs.transposed(which)          # m x k
            assert len(t_rp) == m + 1 and int(t_rp[-1]) == len(ci)
            tb = [[0] * k for _ in range(m)]
            for w in range(m):
                cols = [int(x) for x in t_ci[int(t_rp[w]):int(t_rp[w + 1])]]
                assert cols == sorted(cols)
                for e, g in zip(range(int(t_rp[w]), int(t_rp[w + 1])), cols):
                    tb[w][g] = t_vals[e]
            assert tb == [list(col) for col in zip(*back)]
        # the witness satisfies the sparse system: (A.w) o (B.w) = C.w
        w = c["witness"]
        dot = lambda mat, g: sum(v * w[int(ci)] for ci, v in zip(mat[1][int(mat[0][g]):int(mat[0][g + 1])],
                                                               mat[2][int(mat[0][g]):int(mat[0][g + 1])])) % curve_order
        for g in range(k):
            assert dot(r1cs.mats[0], g) * dot(r1cs.mats[1], g) % curve_order == dot(r1cs.mats[2], g)
    for k in (1, 2, 4, 6, 9):
        assert qd.vandermonde_det(k) == ref_path.qap_vandermonde_det(k)
    rows = qd.SparseR1CS.from_rows(2, 3, [{0: 1}, {2: 5, 1: 7}], [{1: 1}, {0: 1}], [{2: 1}, {2: 1}])
    assert list(rows.mats[0][1]) == [0, 1, 2] and rows.mats[0][2] == [1, 7, 5]      # columns sorted within a row
    sel = qd._Selection([2, 3, 4])
    assert sel.run == (2, 3) and qd._Selection([0, 2]).run is None and len(qd._Selection([])) == 0
