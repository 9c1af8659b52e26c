"""The C-ABI library loads without a GPU and exports every symbol include/cu2b.h declares."""
import os
import re
import subprocess

import cu2rec_b200 as cu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "cu2b.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cu2b_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_reports_version():
    lib = cu._lib.load()
    assert lib.cu2b_version() == 100
    assert lib.cu2b_last_error() is not None


def test_every_declared_symbol_is_exported_and_bound():
    declared = _declared()
    assert len(declared) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", cu._lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\b(cu2b_[a-z0-9_]+)\b", out))
    assert [s for s in declared if s not in exported] == []
    assert sorted(cu._lib.SYMBOLS) == declared  # the ctypes table covers the whole header


def test_no_oracle_in_product():
    """The product must not link, load or import the oracle."""
    out = subprocess.run(["ldd", cu._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    for d in ("cu2rec_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "liboracle" not in src and "import oracle" not in src and "mf_oracle" not in src, f


def test_compute_entry_points_fail_loudly_without_gpu():
    import numpy as np
    import pytest
    from conftest import HAS_GPU
    if HAS_GPU:
        pytest.skip("GPU present")
    with pytest.raises(cu._lib.Cu2bError):
        cu.get_error_metrics_gpu(np.ones(4, np.float32))
