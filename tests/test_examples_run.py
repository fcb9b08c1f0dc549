"""reference tests/test_examples_run.py:13-20: every script under examples/ runs to completion (exit code 0).
The scripts drive the B200 engine, so this is a GPU test; sizes are reduced through FDTS_EXAMPLE_N where the
reference's size would take minutes with the stand-in direct solver."""
import glob
import os
import subprocess
import sys
from os.path import abspath, basename, dirname, join

import pytest

pytestmark = pytest.mark.gpu
ROOT = abspath(join(dirname(__file__), ".."))
SMALL = {"burgers.py": "10", "bbm.py": "500", "cahn-hilliard.py": "24"}


@pytest.fixture(params=sorted(glob.glob(join(ROOT, "examples", "*.py"))), ids=lambda x: basename(x))
def py_file(request):
    return abspath(request.param)


def test_demo_runs(py_file):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    if basename(py_file) in SMALL:
        env["FDTS_EXAMPLE_N"] = SMALL[basename(py_file)]
    subprocess.check_call([sys.executable, py_file], env=env, cwd=ROOT)
