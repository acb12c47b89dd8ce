"""pytest configuration: registers the `gpu` marker, puts the repo root (for
`oracle`, `__graft_entry__`) and the package directory (for `spindyn`) on
sys.path, and builds the native pieces once per session."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "spindynamics.jl_b200")
for p in (ROOT, PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

import spindyn as sd  # noqa: E402  (importing needs neither the .so nor a GPU)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")
