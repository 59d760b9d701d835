"""Stages the reference's own pytest suites for a run over the GPU mirrors.

Copies /root/reference/{zkp,tests} -- unmodified -- into baseline/_ref/reference_suite/ (git-ignored, so no
reference source enters the history; NOT gpurun-ignored, so the copy travels to the GPU box, where
/root/reference does not exist) and drops tests/reference_suite/swap_conftest.py beside them as conftest.py.

    python tools/stage_reference_suite.py            # here, in the build container
    gpurun -- 'cd baseline/_ref/reference_suite && python -m pytest tests -q -p no:cacheprovider'
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ZKP_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref", "reference_suite")


def main():
    if not os.path.isdir(os.path.join(REF, "zkp")):
        sys.exit("reference tree not found at %s (the staging step runs in the build container)" % REF)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", ".pytest_cache")
    for sub in ("zkp", "tests"):
        shutil.copytree(os.path.join(REF, sub), os.path.join(DST, sub), ignore=ignore)
    shutil.copy(os.path.join(ROOT, "tests", "reference_suite", "swap_conftest.py"), os.path.join(DST, "conftest.py"))
    n = sum(len(files) for _, _, files in os.walk(DST))
    print("staged %d files under %s" % (n, DST))


if __name__ == "__main__":
    main()
