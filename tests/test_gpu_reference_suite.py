"""-m gpu: the reference's OWN 469 pytest cases, unmodified, run over the GPU mirrors (north_star: "the tests
run unchanged").  tools/stage_reference_suite.py stages /root/reference/{zkp,tests} into the git-ignored
baseline/_ref/reference_suite/ (it ships to the GPU box with the snapshot) together with
tests/reference_suite/swap_conftest.py, which swaps the hot-path modules for the mirrors before collection.
Fails if the staged suite is present and red; skipped when it has not been staged."""
import os
import re
import subprocess
import sys
import time

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref", "reference_suite")


def test_reference_suites_pass_over_the_mirrors(native):
    if not os.path.isdir(os.path.join(STAGED, "tests")):
        pytest.skip("reference suite not staged (run tools/stage_reference_suite.py in the build container)")
    env = dict(os.environ, ZKP_B200_REPO=ROOT, PYTHONDONTWRITEBYTECODE="1")
    t0 = time.time()
    p = subprocess.run([sys.executable, "-m", "pytest", "tests", "-q", "-p", "no:cacheprovider", "-x"], cwd=STAGED, env=env,
                       capture_output=True, text=True, timeout=1500)
    wall = time.time() - t0
    tail = p.stdout[-3000:]
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "reference_suite_on_mirrors.log"), "w") as fh:
            fh.write("# reference suites (unmodified) over the libzkp_b200 mirrors; wall %.1f s\n" % wall)
            fh.write(p.stdout[-20000:])
            fh.write(p.stderr[-4000:])
    assert p.returncode == 0, tail
    m = re.search(r"(\d+) passed", p.stdout)
    assert m and int(m.group(1)) >= 469, tail
    launched = re.search(r"kernels launched by libzkp_b200 during this session: (\d+)", p.stdout)
    assert launched and int(launched.group(1)) > 1000, "the suite did not run on the GPU path"
