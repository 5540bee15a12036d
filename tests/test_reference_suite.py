"""The reference's OWN test files, unmodified, against this package (authoring container only: it needs the reference
checkout, which never travels to the GPU box -- the test skips there)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = "/root/reference/pygam/tests"


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference checkout is not available here")
def test_the_reference_test_files_pass_against_the_product():
    """tools/run_reference_tests.py: test_terms, test_GAM_methods, test_GAMs, test_gridsearch, test_partial_dependence,
    test_GAM_params, test_penalties, test_utils and test_core of pyGAM 0.12.0 through a throw-away `pygam` package that
    re-exports pygam_b200 with the oracle-backed test engine (public API, host logic; the kernels have their own
    parity tests)"""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_tests.py"), "-x"], capture_output=True,
                         text=True, timeout=900)
    tail = "\n".join(res.stdout.splitlines()[-15:])
    assert res.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
