"""Regenerates tests/golden/ from the reference itself. Run in the build container (needs
/root/reference and oracle/_ref/ref_harness built by `make -C oracle ref`):

    python tests/golden/make_golden.py

* fixtures/: the reference's own test inputs (data/test/*, tiny rating matrices + cfgs)
* ref_init_normal_*.bin : initialize_normal_array outputs   (util.cu:124-144 via ref_harness)
* ref_read_csv_*.json   : readCSV results                   (util.cu:17-45)
* ref_read_config.json  : read_config of test_config.cfg    (config.cu:7-13)
* ref_write_csv.csv     : writeCSV of a fixed float matrix  (util.cu:86-97)
* ref_read_array.json   : read_array of test_Q.csv          (util.cu:52-76)
* reference_tests.json  : the golden constants asserted by the reference's tests (file:line)
"""
import json
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
H = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def run(*a):
    return subprocess.run([H, *map(str, a)], check=True, capture_output=True, text=True).stdout


def main():
    fx = os.path.join(HERE, "fixtures")
    os.makedirs(fx, exist_ok=True)
    for f in sorted(os.listdir(os.path.join(REF, "data", "test"))):
        shutil.copy(os.path.join(REF, "data", "test", f), os.path.join(fx, f))
    for size, k in [(64, 2), (1000, 32), (257, 128), (50, 50)]:
        run("init_normal", size, k, os.path.join(HERE, "ref_init_normal_%d_%d.bin" % (size, k)))
    csv = {}
    for name in ["test_ratings.csv", "test_missing_user_ratings.csv", "test_ratings2.csv", "test_ratings3.csv",
                 "test_user_ratings.csv"]:
        tmp = "/tmp/_golden_csv.bin"
        run("read_csv", os.path.join(fx, name), tmp)
        raw = open(tmp, "rb").read()
        n, rows, cols = np.frombuffer(raw[:12], dtype=np.int32)
        gb = np.frombuffer(raw[12:16], dtype=np.float32)[0]
        trip = np.frombuffer(raw[16:], dtype=np.dtype([("u", "<i4"), ("i", "<i4"), ("r", "<f4")]))
        csv[name] = dict(n=int(n), rows=int(rows), cols=int(cols), global_bias_bits=int(np.float32(gb).view(np.uint32)),
                         users=trip["u"].tolist(), items=trip["i"].tolist(), ratings=trip["r"].tolist())
    json.dump(csv, open(os.path.join(HERE, "ref_read_csv.json"), "w"), indent=1)
    cfg = run("read_config", os.path.join(fx, "test_config.cfg")).split()
    json.dump(dict(fields=cfg), open(os.path.join(HERE, "ref_read_config.json"), "w"))
    # writeCSV on awkward floats (ties, negatives, tiny, large)
    vals = np.array([0.0, -0.0, 1.0, -1.5, 0.1234565, 0.1234575, 2.5e-7, 5e-7, 1.5e-6, 123456.789, -3.999999,
                     0.9999995, 1e-10, 3.4e12, 0.5, 0.0000005, 1.0000005, 7.0000005, 33.333333, -0.000001],
                    dtype=np.float32)
    rng = np.random.RandomState(7)
    more = (rng.standard_normal(100) * rng.choice([1e-3, 1.0, 50.0], 100)).astype(np.float32)
    mat = np.concatenate([vals, more]).astype(np.float32)
    mat.tofile(os.path.join(HERE, "write_csv_input.bin"))
    run("write_csv", os.path.join(HERE, "write_csv_input.bin"), 24, 5, os.path.join(HERE, "ref_write_csv.csv"))
    tmp = "/tmp/_golden_arr.bin"
    out = run("read_array", os.path.join(fx, "test_Q.csv"), tmp)
    raw = open(tmp, "rb").read()
    r, c = np.frombuffer(raw[:8], dtype=np.int32)
    json.dump(dict(n_rows=int(r), n_cols=int(c), values=np.frombuffer(raw[8:], dtype=np.float32).tolist()),
              open(os.path.join(HERE, "ref_read_array.json"), "w"))
    json.dump({
        "test_loss.cu:90": {"sum_sq_err": 74.0, "setup": "test_ratings.csv, k=2, P=Q=1, biases=1, global_bias=1"},
        "test_loss.cu:107-109,137-138": {"problem_sizes": [1, 33, 1024, 65536], "mae": 1.0, "rmse": 1.0},
        "test_util.cu:28-31": {"rows": 6, "cols": 5, "n": 18, "global_bias": 3.5555555555555, "tol": 1e-3},
        "test_util.cu:43": {"read_array_first10": list(range(10)), "tol": 1e-3},
        "test_util.cu:123-125": {"indptr": [0, 4, 7, 10, 13, 16, 18],
                                 "indices": [0, 1, 2, 4, 0, 1, 2, 0, 1, 2, 0, 1, 2, 1, 3, 4, 3, 4],
                                 "data": [1, 1, 1, 5, 3, 3, 3, 4, 4, 4, 5, 5, 5, 2, 4, 4, 5, 5]},
        "test_util.cu:170-172": {"indptr": [0, 4, 4, 7, 10, 13, 15],
                                 "indices": [0, 1, 2, 4, 0, 1, 2, 0, 1, 2, 1, 3, 4, 3, 4],
                                 "data": [1, 1, 1, 5, 4, 4, 4, 5, 5, 5, 2, 4, 4, 5, 5]},
        "test_config.cu:14-15": {"total_iterations": 100, "P_reg": 0.2, "tol": 1e-4},
        "test_training.cu:45": "losses[0] >= losses[9] after 10 iterations, k=2, lr=1e-3, reg=0.1, train==test",
        "test_sgd.cu:134-145": "no NaN in P, Q, biases after one update (k=1, lr=0.07, reg=0.1)",
    }, open(os.path.join(HERE, "reference_tests.json"), "w"), indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
