#!/bin/bash
# Runs the reference's UNMODIFIED test-suite against this repository's overlay of `src.core` (and
# `src.dnn_test_prio.aggregate_statistics`).  Needs the reference checkout (default /root/reference) — i.e. the
# authoring container; on a GPU box every test can pass, on a CPU-only box the tests that score fail loudly
# ("needs a CUDA device"), which is the no-CPU-fallback rule at work.
#   tools/run_reference_tests.sh [reference_root] [extra pytest args]
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
shift || true
cd "$REPO"
PYTHONPATH="$REPO:$REF" python - "$REF" "$@" <<'PY'
import sys, types
import numpy as np
ref = sys.argv[1]
if not hasattr(np, "int"):
    np.int = int                      # alias removed in NumPy 1.24, used by the reference's tests/helpers
# uncertainty_wizard is not installable offline: the stub of oracle/ref_harness.py (test infrastructure)
from oracle import ref_harness
sys.modules.setdefault("uncertainty_wizard", ref_harness._uwiz_stub())
sys.modules.setdefault("uncertainty_wizard.quantifiers", sys.modules["uncertainty_wizard"].quantifiers)
import pytest
sys.exit(pytest.main([f"{ref}/tests", "--import-mode=importlib", "-q", "-p", "no:cacheprovider", "--rootdir", ref,
                      "--ignore", f"{ref}/tests/test_model.py"] + sys.argv[2:]))   # test_model.py needs TensorFlow
PY
