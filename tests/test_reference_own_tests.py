"""The reference's OWN test files (tests/core/test_ndmps.py, tests/utils/test_core.py, test_filetools.py,
test_metrics.py) run against the reference's own source with the oracle standing in for quimb / scikit-image
(``tests/golden/run_reference_tests.py``).  Build container only: skipped where /root/reference does not exist (the
GPU box); never part of ``-m gpu``."""
import subprocess
import sys
from pathlib import Path

import pytest

RUNNER = Path(__file__).resolve().parent / "golden" / "run_reference_tests.py"
PACKAGE_RUNNER = Path(__file__).resolve().parent / "golden" / "run_reference_tests_on_package.py"


@pytest.mark.skipif(not Path("/root/reference/tests/core/test_ndmps.py").exists(), reason="reference tree not present")
def test_reference_test_suite_passes_on_the_oracle_stand_ins():
    run = subprocess.run([sys.executable, str(RUNNER)], capture_output=True, text=True, timeout=900)
    tail = run.stdout.strip().splitlines()[-1] if run.stdout.strip() else run.stderr[-400:]
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
    assert " passed" in tail and "failed" not in tail and "error" not in tail, tail
    assert int(tail.split(" passed")[0].split()[-1]) >= 65, tail


@pytest.mark.skipif(not Path("/root/reference/tests/evaluation/test_benchmark.py").exists(), reason="reference tree not present")
def test_reference_host_side_tests_pass_on_this_package():
    """tests/utils/test_core.py, test_filetools.py and tests/evaluation/test_benchmark.py of the reference, unmodified,
    against the drop-in's ``utils.core``, ``utils.filetools`` and ``evaluation.benchmark`` (no GPU: the benchmark tests
    replace ``NDMPS`` and the metrics through the module's globals, which the drop-in keeps)."""
    run = subprocess.run([sys.executable, str(PACKAGE_RUNNER)], capture_output=True, text=True, timeout=600)
    tail = [line for line in run.stdout.strip().splitlines() if " passed" in line]
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
    assert tail and "failed" not in tail[-1] and "error" not in tail[-1], run.stdout[-500:]
    assert int(tail[-1].split(" passed")[0].split()[-1]) >= 59, tail[-1]
