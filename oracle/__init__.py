"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's prover hot path (BN254 group law, Fr NTT,
Groth16/PLONK prover functions).  Nothing under this directory is product
code: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or the timed CPU baseline.  The product package
(``interactive_zkp_study_b200``) never imports it and raises if its CUDA
library is missing.

Parity pinning (see DESIGN.md "Oracle"):
  * ``oracle/shim/py_ecc`` restates py-ecc==7.0.1's non-optimised bn128
    surface (third-party, pinned in /root/reference/requirements.txt:14,
    NOT vendored under /root/reference and not installable here).
  * The shim is pinned by running the reference's OWN 469 pytest cases on it
    (``oracle/run_reference_tests.sh``), by public BN254 known answers
    (EIP-196 2*G1, generator order, pairing bilinearity) and by the two
    numeric comments in the reference that are valid.
  * ``oracle/ref_path.py`` restates the reference's hot-path functions over
    plain ints and is pinned against the reference's own modules imported
    from /root/reference (``tests/golden/make_golden.py`` -> fixtures).
"""
