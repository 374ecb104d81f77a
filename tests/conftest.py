import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))

    return load


class ParityLog:
    """Collects the ACHIEVED relative errors of the parity assertions and writes them to
    gpurun_out/parity_r02.txt at the end of the session (copied to profiles/ by hand after a GPU run)."""

    def __init__(self):
        self.rows = []

    def check(self, test, case, quantity, err, scale, tol_rel, note=""):
        """Records err/scale and asserts it below tol_rel.  `scale` is what the tolerance is relative to."""
        scale = max(float(scale), 1e-300)
        ratio = float(err) / scale
        self.rows.append((test, case, quantity, ratio, tol_rel, note))
        assert ratio < tol_rel, (test, case, quantity, ratio, tol_rel)

    def dump(self):
        if not self.rows:
            return
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_r02.txt"), "w") as f:
            f.write("# achieved parity errors, device (through the C ABI) against the reference fixtures / the pinned oracle\n")
            f.write("# columns: test | case | quantity | achieved error / scale | asserted bound | scale\n")
            for t, c, q, r, tol, note in self.rows:
                f.write(f"{t:46s} {c:30s} {q:14s} {r:10.3e}  < {tol:7.1e}  {note}\n")


_PARITY = ParityLog()


@pytest.fixture(scope="session")
def parity_log():
    return _PARITY


def pytest_sessionfinish(session, exitstatus):
    _PARITY.dump()
