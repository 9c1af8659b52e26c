"""The header-only C++ shim (cu2rec_b200/csrc/cu2rec_shim.h) on the host: a small program written against
the reference's own names (readCSV, createSparseMatrix, read_array, initialize_normal_array, writeToFile,
config::Config) is compiled here and must print what tests/test_util.cu / tests/test_config.cu assert."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim_output(tmp_path_factory, fixtures_dir):
    tmp = tmp_path_factory.mktemp("shim")
    exe = tmp / "shim_host_check"
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "cu2rec_b200", "csrc"), "-o", str(exe), os.path.join(ROOT, "tests", "shim_host_check.cpp"),
                    "-L" + os.path.join(ROOT, "cu2rec_b200", "lib"), "-lcu2b", "-Wl,-rpath," + os.path.join(ROOT, "cu2rec_b200", "lib")],
                   check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe), fixtures_dir, str(tmp)], check=True, capture_output=True, text=True)
    return out.stdout.splitlines(), out.stderr, tmp


def _line(lines, prefix):
    return [ln for ln in lines if ln.startswith(prefix)][0][len(prefix):].strip()


def test_shim_reads_and_builds_what_the_reference_tests_expect(shim_output, golden_dir):
    lines, _, _ = shim_output
    ref = json.load(open(os.path.join(golden_dir, "reference_tests.json")))
    assert _line(lines, "read_csv") .startswith("rows=6 cols=5 n=18 global_bias=3.5555")  # test_util.cu:28-31
    for tag, key in (("sparse", "test_util.cu:123-125"), ("missing", "test_util.cu:170-172")):
        g = ref[key]
        assert [int(x) for x in _line(lines, tag + " indptr").split()] == g["indptr"]
        assert [int(x) for x in _line(lines, tag + " indices").split()] == g["indices"]
        assert [float(x) for x in _line(lines, tag + " data").split()] == [float(x) for x in g["data"]]
    first = [float(x) for x in _line(lines, "read_array").split("first")[1].split()]
    assert np.allclose(first, ref["test_util.cu:43"]["read_array_first10"], atol=1e-3)
    assert _line(lines, "read_array_missing") == "nullptr"  # util.cu:69-71


def test_shim_initialisation_writer_and_config(shim_output, golden_dir):
    lines, _, tmp = shim_output
    want = np.fromfile(os.path.join(golden_dir, "ref_init_normal_64_2.bin"), dtype=np.uint32)
    assert [int(x, 16) for x in _line(lines, "init_normal bits").split()] == want.tolist()
    assert _line(lines, "init_normal overloads_agree=") == "1"
    assert (tmp / "test_ratings_f2_p.csv").read_text() == "1.000000,1.000000\n" * 6  # test_util.cu:49-92
    assert (tmp / "test_ratings_f2_global_bias.csv").read_text() == "3.500000\n"
    assert _line(lines, "config defaults") == "total_iterations=5000 n_factors=50 check_error=500"  # config.h:25-45
    assert _line(lines, "config read") == "total_iterations=100 P_reg=0.200000"  # test_config.cu:14-15
    assert _line(lines, "config round_trip") == "total_iterations=250 P_reg=0.300000"  # test_config.cu:19-26
    assert _line(lines, "config unreadable_keeps") == "total_iterations=250"
    assert "Hyperparameters:" in lines and "total_iterations: 250" in lines and "is_train: true" in lines


def test_shim_error_behaviour(shim_output):
    lines, err, _ = shim_output
    assert _line(lines, "read_csv_missing") == "n=0" and "ERROR: The file isnt open." in err  # util.cu:41-44
    assert _line(lines, "unsorted").startswith("threw:") and "not grouped by ascending user" in _line(lines, "unsorted")
