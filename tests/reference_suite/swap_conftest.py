"""conftest.py of the STAGED copy of the reference's own test suites (tools/stage_reference_suite.py puts it
at the root of baseline/_ref/reference_suite/, next to the copied `zkp/` and `tests/` of the reference).

Before any test module is imported it replaces the reference's hot-path modules by the GPU mirrors of
interactive_zkp_study_b200, so the reference's UNMODIFIED tests
(tests/plonk/test_prover.py, test_e2e.py, test_crypto.py, test_foundation.py, tests/groth16/test_proving.py,
test_setup.py, test_verifying.py, test_integration.py, test_poly_utils.py, ...) exercise libzkp_b200
through the reference's own call signatures.  Everything that is not on the hot path (circuit.py,
code_to_r1cs.py, qap_creator*.py, determinant.py, the tests themselves) stays the reference's code.

py_ecc 7.0.1 cannot be installed in this image (no network): oracle/shim/py_ecc restates its bn128 surface
and stands in for it HERE, for what the reference's tests and verifiers need from it that is out of scope
for the GPU path (FQ / FQ2 / FQ12 containers, pairing, is_on_curve).  A deployment has the real py_ecc.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.environ.get("ZKP_B200_REPO") or os.path.abspath(os.path.join(HERE, "..", "..", ".."))
for p in (os.path.join(REPO, "oracle", "shim"), REPO, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

SWAPPED = ("groth16.proving", "groth16.poly_utils", "groth16.setup", "groth16.verifying", "plonk.verifier", "plonk.field",
           "plonk.polynomial", "plonk.kzg", "plonk.utils", "plonk.permutation", "plonk.srs", "plonk.preprocessor",
           "plonk.transcript", "plonk.prover", "plonk.prover.round1", "plonk.prover.round2", "plonk.prover.round3",
           "plonk.prover.round4", "plonk.prover.round5")

import zkp  # noqa: E402  (the staged copy of the reference's package: parents of the swapped modules)
import zkp.groth16  # noqa: E402,F401
import zkp.plonk  # noqa: E402,F401

for name in SWAPPED:
    mod = importlib.import_module("interactive_zkp_study_b200.zkp." + name)
    sys.modules["zkp." + name] = mod
    parent, _, leaf = ("zkp." + name).rpartition(".")
    setattr(sys.modules[parent], leaf, mod)


def pytest_report_header(config):
    from interactive_zkp_study_b200 import native
    try:
        device = native.device_info()["name"]
    except native.ZkpB200Error as e:      # collection-only runs on a box without a GPU; every compute test fails loudly
        device = "NONE (%s)" % str(e)[:60]
    return ["reference suites on the libzkp_b200 mirrors: %d modules swapped, device %s" % (len(SWAPPED), device)]


def pytest_sessionfinish(session, exitstatus):
    from interactive_zkp_study_b200 import native
    try:
        print("\n[mirror] kernels launched by libzkp_b200 during this session: %d" % native.launch_count())
    except native.ZkpB200Error:
        pass
