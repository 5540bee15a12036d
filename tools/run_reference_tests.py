"""Run the REFERENCE's own test files against the product (authoring container only: needs /root/reference).

A throw-away package named `pygam` is created in a temporary directory: it re-exports pygam_b200, swaps in the
oracle-backed test engine (no GPU needed: what is exercised is the public API, the term grammar and the host logic),
and borrows the reference's dataset loaders.  The reference's test files are read where they lie; nothing is copied into
the repository.

    python tools/run_reference_tests.py            # the nine files that test the public API, the penalties, the input checks and the repr helpers
    python tools/run_reference_tests.py -k summary # extra arguments go to pytest

Expected: 137 passed, 1 skipped (test_datasets.py and test_gen_imgs.py are not run: dataset loaders and
documentation images are outside the path).
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/pygam"
FILES = ["test_terms.py", "test_GAM_methods.py", "test_GAMs.py", "test_gridsearch.py", "test_partial_dependence.py", "test_penalties.py", "test_utils.py", "test_core.py",
         "test_GAM_params.py"]

SHIM = '''
import sys, types
for p in ({root!r}, {root!r} + "/tests", {root!r} + "/tests/golden", {root!r} + "/baseline/stub"):
    sys.path.insert(0, p)
import numpy as np
import pygam_b200 as _pg
from pygam_b200 import *  # noqa
from pygam_b200 import engine as _pe, gam as _gam, terms as _terms, penalties as _pen
from oracle_backend import OracleBackend
_pe.swap_backend_for_tests(OracleBackend())
for _n in dir(_pg):
    if not _n.startswith("__"):
        globals()[_n] = getattr(_pg, _n)
def _alias(name, mod):
    sys.modules["pygam." + name] = mod
    globals()[name] = mod
_alias("pygam", _gam); _alias("terms", _terms); _alias("penalties", _pen)
from pygam_b200 import utils as _utils, core as _core
_alias("utils", _utils); _alias("core", _core)
'''


def main():
    if not os.path.isdir(REF):
        sys.exit("the reference is not available here")
    tmp = tempfile.mkdtemp(prefix="refcheck_")
    try:
        os.makedirs(os.path.join(tmp, "pygam"))
        os.makedirs(os.path.join(tmp, "tests"))
        with open(os.path.join(tmp, "pygam", "__init__.py"), "w") as f:
            f.write(SHIM.format(root=ROOT))
        os.symlink(os.path.join(REF, "datasets"), os.path.join(tmp, "pygam", "datasets"))
        for name in FILES + ["conftest.py", "__init__.py"]:
            os.symlink(os.path.join(REF, "tests", name), os.path.join(tmp, "tests", name))
        cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider"] + [os.path.join("tests", f) for f in FILES]
        return subprocess.call(cmd + sys.argv[1:], cwd=tmp)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
