"""not gpu: the device's modular inversion (csrc/modinv30.cuh, batched division steps) is plain C++ apart from its
function qualifiers, so the very source the kernels use is compiled here with g++ and checked against Python's
pow(x, -1, p) for both BN254 fields (what py_ecc's prime_field_inv returns, with inv(0) = 0)."""
import os
import random
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "interactive_zkp_study_b200", "csrc")
P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617

HARNESS = r"""
#define __host__
#define __device__
#include "bn254_params.cuh"
#include "modinv30.cuh"
#include <cstdio>
using namespace zkp;
int main() {
  char line[256];
  while (fgets(line, sizeof line, stdin)) {
    uint32_t x[8], out[8];
    for (int i = 0; i < 8; i++) { unsigned v; sscanf(line + 2 + 8 * (7 - i), "%8x", &v); x[i] = v; }
    if (line[0] == '0') modinv30::inverse<FpParams>(x, out); else modinv30::inverse<FrParams>(x, out);
    for (int i = 7; i >= 0; i--) printf("%08x", out[i]);
    printf("\n");
  }
}
"""


@pytest.mark.skipif(shutil.which("g++") is None, reason="no host compiler")
def test_division_step_inverse_matches_python(tmp_path):
    src = tmp_path / "h.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "h"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", CSRC, str(src), "-o", str(exe)], check=True)
    rng = random.Random(30)
    cases = []
    for field, m in ((0, P), (1, R)):
        vals = [0, 1, 2, 3, m - 1, m - 2, (m + 1) // 2, (1 << 255) % m, 1 << 128, 1 << 253, m // 3, (1 << 30) - 1, 1 << 30]
        vals += [rng.randrange(m) for _ in range(1500)]
        vals += [rng.randrange(1 << k) for k in range(1, 254) for _ in range(2)]
        cases += [(field, m, v % m) for v in vals]
    text = "".join("%d %064x\n" % (f, v) for f, _, v in cases)
    out = subprocess.run([str(exe)], input=text, capture_output=True, text=True, check=True).stdout.split()
    assert len(out) == len(cases)
    for (f, m, v), o in zip(cases, out):
        assert int(o, 16) == (pow(v, -1, m) if v else 0), (f, hex(v))
