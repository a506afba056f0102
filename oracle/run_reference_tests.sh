#!/bin/bash
# Runs the reference's own hot-path tests, unmodified, from /root/reference (read-only),
# through the stand-in modules in oracle/ref_shims.  Build-container only.
here="$(cd "$(dirname "$0")" && pwd)"
export PYTHONPATH="$here/ref_shims:/root/reference:$PYTHONPATH"
export NUMBA_CACHE_DIR=/tmp/numba_cache
cd /tmp && exec python -m pytest -p no:cacheprovider -q \
  /root/reference/tests/test_bsplines.py /root/reference/tests/test_representation.py \
  /root/reference/tests/test_calculator.py /root/reference/tests/test_distances.py \
  /root/reference/tests/test_geometry.py /root/reference/tests/test_optimize.py \
  /root/reference/tests/test_composition.py "$@"
