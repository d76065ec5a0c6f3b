import os
import subprocess
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Build the native library, launcher, benchmarks and C test programs once per session."""
    subprocess.run(["make", "-C", str(REPO), "-j8", "all"], check=True, capture_output=True)
    return REPO


def acxrun(repo, np_, prog, *args, env=None, timeout=180):
    e = dict(os.environ)
    e.setdefault("ACX_WATCHDOG_MS", "30000")
    if env:
        e.update(env)
    cmd = [str(repo / "bin" / "acxrun"), "-np", str(np_), "--timeout", str(timeout - 10), str(prog), *map(str, args)]
    return subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=timeout)
