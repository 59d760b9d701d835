#!/usr/bin/env bash
# TEST INFRASTRUCTURE (oracle pinning).  Runs the reference's OWN pytest suites
# (/root/reference/tests, 469 cases) with oracle/shim standing in for the absent
# third-party py-ecc==7.0.1.  Only meaningful in the build container
# (/root/reference does not exist on the GPU box).  Output is summarised in
# tests/golden/reference_suite_on_shim.txt.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
cd /root/reference
PYTHONDONTWRITEBYTECODE=1 PYTHONPATH="$HERE/shim" \
  python -m pytest tests/ -q -p no:cacheprovider "$@"
