"""Run the REFERENCE'S OWN host-side test files against THIS package (the drop-in), not against the reference's source:
tests/utils/test_core.py, tests/utils/test_filetools.py and tests/evaluation/test_benchmark.py import
``imgcompressionmps.utils.core``, ``.utils.filetools`` and ``.evaluation.benchmark`` - the import paths the package
keeps - and need no GPU (the benchmark tests replace ``NDMPS`` and the metric functions through the module's globals).

Build container only (needs /root/reference; nothing is written there):
    python tests/golden/run_reference_tests_on_package.py [pytest args]

The package is imported first, so the ``sys.path`` lines of the reference's conftest / test files cannot shadow it; the
runner checks that every module the tests exercise really came from this repository.
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[2]
PKG = ROOT / "img-compression-mps_b200"
REF = Path("/root/reference")
FILES = ["tests/utils/test_core.py", "tests/utils/test_filetools.py", "tests/evaluation/test_benchmark.py"]


class _Origin:
    """After the run: the modules under test must be the package's, not the reference's."""

    def pytest_sessionfinish(self, session, exitstatus):
        for name in ("imgcompressionmps.utils.core", "imgcompressionmps.utils.filetools", "imgcompressionmps.evaluation.benchmark"):
            origin = getattr(sys.modules.get(name), "__file__", "") or ""
            if not origin.startswith(str(PKG)):
                print(f"\\n{name} was imported from {origin!r}, not from {PKG}")
                session.exitstatus = 3


if __name__ == "__main__":
    if not REF.exists():
        print("reference tree not present")
        sys.exit(5)
    sys.path.insert(0, str(PKG))
    import imgcompressionmps                                  # noqa: F401  (first: later imports resolve inside it)
    import imgcompressionmps.evaluation.benchmark             # noqa: F401
    import imgcompressionmps.utils.core                       # noqa: F401
    import imgcompressionmps.utils.filetools                  # noqa: F401
    args = [str(REF / f) for f in FILES] + ["-q", "-p", "no:cacheprovider", "--rootdir", str(REF), "-W", "ignore"] + sys.argv[1:]
    sys.exit(pytest.main(args, plugins=[_Origin()]))
