"""Run the REFERENCE'S OWN test files against the reference's own source, with quimb / scikit-image replaced by the
stand-ins of ``make_golden_reference_exec.py`` (the oracle's restatement of those libraries).

Build container only (needs /root/reference, which is read-only: the cache provider is off and nothing is written there):
    python tests/golden/run_reference_tests.py [pytest args]

What a green run says: with the oracle's TT-SVD / bond compression / overlap / SSIM in the place of the two absent
libraries, the reference's package passes the tests its authors wrote for this path (tests/core/test_ndmps.py,
tests/utils/test_core.py, test_filetools.py, test_metrics.py) - round trip at atol 1e-10, norm 1 at rel 1e-12,
write-through arrays and norm refresh at rel 1e-12, shrinking element counts, disk ratio in (0, 1), 20 printed lines.
"""
import sys
from pathlib import Path

import pytest

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden_reference_exec as stand_ins      # noqa: E402

REF = Path("/root/reference")
FILES = ["tests/core/test_ndmps.py", "tests/utils/test_core.py", "tests/utils/test_filetools.py", "tests/utils/test_metrics.py"]

if __name__ == "__main__":
    if not REF.exists():
        print("reference tree not present")
        sys.exit(5)
    stand_ins.install_stand_ins()
    sys.path.insert(0, str(REF / "src"))
    args = [str(REF / f) for f in FILES] + ["-q", "-p", "no:cacheprovider", "--rootdir", str(REF), "-W", "ignore"] + sys.argv[1:]
    sys.exit(pytest.main(args))
