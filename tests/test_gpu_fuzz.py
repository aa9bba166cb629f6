"""Random-shape parity fuzz (scripts/fuzz_parity.py) as a regression test: a fixed, seeded set of ragged / odd / pitched shapes of
the predictive, the GGN kernels, the SYRK and EPIG against the oracle, inside the supported envelope stated in DESIGN.md section 2."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("seed", [11, 12])
def test_seeded_fuzz_cases(seed):
    res = subprocess.run([sys.executable, str(ROOT / "scripts" / "fuzz_parity.py"), "n=160", str(seed)], cwd=ROOT,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
