"""Drop-in for /root/reference/zkp/groth16/proving.py: same signatures, same results.

The reference walks a numWires x numGates grid of affine scalar multiplications
(:27-31, :39-43, :56-60).  Algebraically each grid is ONE multi-scalar multiplication over the
numGates CRS points with scalars u_j = sum_i Rx_i * M_ij, so each proof element is a
single GPU MSM with the fixed terms appended:

  proof_a = [alpha]_1 + sum_j uA_j [x^j]_1 + r [delta]_1
  proof_b = [beta]_2  + sum_j uB_j [x^j]_2 + s [delta]_2
  proof_c = s A + r (beta_1 + sum_j uB_j [x^j]_1 + s delta_1) - (s r) delta_1
            + sum_{i not public} Rx_i sigma1_4[i] + sum_{i < numGates-1} Hx_i sigma1_5[i]
          = s A + r beta_1 + sum_j (r uB_j) [x^j]_1 + (the two sums)      (the delta terms cancel)

Scalars are reduced mod the curve order; the group has that order, so k*P == (k mod r)*P and the
affine result is the unique representative the reference computes.
"""
from ... import native, tables
from ...compat import FQ, FR, G1, G2, curve_order, g1_from_ints, g2_from_ints  # noqa: F401
from .poly_utils import getNumGates, getNumWires

g1 = G1
g2 = G2
pointInf1 = None
pointInf2 = None


def _u(Rx, M):
    """u_j = sum_i Rx_i * M_ij  (device mat-vec)."""
    rows, cols = len(M), len(M[0])
    flat = [int(x) % curve_order for row in M for x in row]
    rx = [int(x) % curve_order for x in Rx[:rows]]
    return native.fr_vec_from_bytes(
        native.fr_vec_matrix(native.fr_vec_bytes(rx), native.fr_vec_bytes(flat), rows, cols))


def _msm_g1(points, scalars):
    n = len(points)
    return g1_from_ints(native.g1_msm(native.g1_vec_bytes(points), native.fr_vec_bytes(scalars), n))


def proof_a(sigma1_1, sigma1_2, Ax, Rx, r):
    numGates = getNumGates(Ax)
    u = _u(Rx, Ax)
    points = [sigma1_1[0]] + list(sigma1_2[:numGates]) + [sigma1_1[2]]
    scalars = [1] + u + [int(r) % curve_order]
    return _msm_g1(points, scalars)


def proof_b(sigma2_1, sigma2_2, Bx, Rx, s):
    numGates = getNumGates(Bx)
    u = _u(Rx, Bx)
    points = [sigma2_1[0]] + list(sigma2_2[:numGates]) + [sigma2_1[2]]
    scalars = [1] + u + [int(s) % curve_order]
    n = len(points)
    return g2_from_ints(native.g2_msm(native.g2_vec_bytes(points), native.fr_vec_bytes(scalars), n))


def proof_c(sigma1_1, sigma1_2, sigma1_4, sigma1_5, Bx, Rx, Hx, s, r, prf_A, pub_r_indexs=None):
    if pub_r_indexs == None:  # noqa: E711  (kept as the reference writes it)
        pub_r_indexs = [0, 1]
    numGates = getNumGates(Bx)
    numWires = getNumWires(Bx)
    s_i, r_i = int(s) % curve_order, int(r) % curve_order
    u = _u(Rx, Bx)
    ru = native.fr_vec_from_bytes(native.fr_vec_op(3, native.fr_vec_bytes(u), native.fe_bytes(r_i), len(u)))
    points = [prf_A, sigma1_1[1]] + list(sigma1_2[:numGates])
    scalars = [s_i, r_i] + ru
    for i in range(numWires):
        if i in pub_r_indexs:
            continue  # placeholder (FQ(0), FQ(0)) entries are never touched (setup.py:39,50)
        points.append(sigma1_4[i])
        scalars.append(int(Rx[i]) % curve_order)
    for i in range(numGates - 1):
        points.append(sigma1_5[i])
        scalars.append(int(Hx[i]) % curve_order)
    return _msm_g1(points, scalars)


def build_rpub_enum(pub_r_indexs, r_vec):
    return [(i, r_vec[i]) for i in pub_r_indexs]
