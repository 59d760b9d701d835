"""CPU: the C ABI builds, loads, and exports exactly what include/zkp_b200.h declares; without a
GPU every compute entry point fails loudly (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header="zkp_b200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zkp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from interactive_zkp_study_b200 import _lib
    lib = _lib.load_library()
    product, diag = _declared(), _declared("zkp_b200_diag.h")
    assert len(product) >= 40
    for name in product + diag:
        assert hasattr(lib, name), name
    assert sorted(_lib.PROTOTYPES) == sorted(product + diag)  # the ctypes table and the two headers agree
    # experiments and self-test hooks stay out of the product header
    assert not [n for n in product if n.startswith("zkp_dbg_") or n in ("zkp_latency_probe", "zkp_imad_peak")]


def test_header_cites_the_reference_interface():
    src = open(os.path.join(ROOT, "include", "zkp_b200.h")).read()
    for cite in ("zkp/plonk/kzg.py:59-67", "zkp/groth16/proving.py", "zkp/plonk/polynomial.py:292-341",
                 "zkp/groth16/poly_utils.py:116-125", "zkp/plonk/srs.py:78-82"):
        assert cite in src


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_cpu_fallback_without_a_gpu():
    from interactive_zkp_study_b200 import _lib, native
    lib = _lib.load_library()
    out = (ctypes.c_uint8 * 64)()
    inf = ctypes.c_int()
    rc = lib.zkp_g1_msm(None, None, 0, out, ctypes.byref(inf))
    assert rc == -2                                   # ZKP_ERR_NOT_INITIALISED
    assert b"no CPU fallback" in lib.zkp_last_error()
    with pytest.raises(native.ZkpB200Error):
        native.g1_msm(b"", b"", 0)
    from interactive_zkp_study_b200.zkp.plonk.polynomial import Polynomial
    with pytest.raises(native.ZkpB200Error):
        Polynomial([1, 2]) * Polynomial([3, 4])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "interactive_zkp_study_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)
                assert "oracle." not in text.replace("oracle/", ""), os.path.join(dirpath, f)


def test_oracle_shim_never_masquerades_as_py_ecc():
    """The oracle's py_ecc restatement is importable only as oracle.shim.py_ecc; loading it must not
    make `import py_ecc` succeed for the product (compat.py probes for the real package)."""
    import sys
    from oracle import plonk_verifier
    bn = plonk_verifier._pairing()
    assert bn.__name__ == "oracle.shim.py_ecc.bn128"
    import importlib
    if "py_ecc" not in sys.modules:          # nothing named py_ecc is importable in this image
        with pytest.raises(ImportError):
            importlib.import_module("py_ecc")
    from interactive_zkp_study_b200 import compat
    assert compat.HAVE_PY_ECC is False or "site-packages" in sys.modules["py_ecc"].__file__


def test_engine_options_are_validated_by_name_and_range():
    """zkp_msm_set_option is host-only state: every documented name is accepted with 0 (= automatic) and with a
    value in range, anything else is refused with ZKP_ERR_INVALID_ARGUMENT and an error text (no GPU needed)."""
    from interactive_zkp_study_b200 import _lib
    lib = _lib.load_library()
    good = {"window_bits": 16, "accumulate": 2, "tree_items": 64, "tree_rounds": 3, "parts": 4,
            "reduce_radix": 8, "wide_log2": 15, "quad_log2": 14}
    try:
        for name, value in good.items():
            assert lib.zkp_msm_set_option(name.encode(), value) == 0, name
        for name, value in (("reduce_radix", 3), ("reduce_radix", 64), ("accumulate", 3), ("parts", 9),
                            ("window_bits", 21), ("wide_log2", 32), ("no_such_option", 1)):
            assert lib.zkp_msm_set_option(name.encode(), value) == -3, (name, value)   # ZKP_ERR_INVALID_ARGUMENT
            assert lib.zkp_last_error()
        assert lib.zkp_msm_set_option(None, 0) == -3
    finally:
        for name in good:
            assert lib.zkp_msm_set_option(name.encode(), 0) == 0
