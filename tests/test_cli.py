"""bin/mf and bin/predict: the process surface of the drop-in (mf.cu:16-99, predict.cu:72-146)."""
import os
import re
import subprocess

import numpy as np
import pytest

import cu2rec_b200 as cu
import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MF = os.path.join(ROOT, "bin", "mf")
PREDICT = os.path.join(ROOT, "bin", "predict")


def _write_csv(path, r):
    with open(path, "w") as f:
        f.write("userId,itemId,rating\n")
        f.write("".join("%d,%d,%.1f\n" % (u + 1, i + 1, x) for u, i, x in r))


def test_cli_argument_handling_matches_reference():
    assert os.path.exists(MF) and os.path.exists(PREDICT), "run __graft_entry__.build()"
    assert subprocess.run([MF]).returncode == 255           # mf.cu:17-19 `return -1`
    p = subprocess.run([MF, "-x"], capture_output=True, text=True)
    assert p.returncode == 1 and "Unknown option." in p.stdout  # mf.cu:28-29
    assert subprocess.run([PREDICT]).returncode == 2        # predict.cu:73-75
    p = subprocess.run([PREDICT, "-z"], capture_output=True, text=True)
    assert p.returncode == 1 and "Unknown option." in p.stdout


@pytest.mark.gpu
def test_mf_cli_end_to_end_and_outputs_feed_reference_predict(tmp_path):
    tr, te = cu.synth_ratings(400, 120, 12000, rank=4, noise=0.3, integer_ratings=True, seed=3)
    _write_csv(tmp_path / "train.csv", tr)
    _write_csv(tmp_path / "test.csv", te)
    (tmp_path / "c.cfg").write_text("0 120 8 0.01 42 0.02 0.02 0.02 0.02")
    p = subprocess.run([MF, "-c", str(tmp_path / "c.cfg"), str(tmp_path / "train.csv"), str(tmp_path / "test.csv")],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    out = p.stdout
    # stdout contract (mf.cu:37, config.cu:51-63, training.cu:135,137,177)
    assert re.search(r"^Free memory: \d+\n\n", out)
    assert "Hyperparameters:\ntotal_iterations: 120\nn_factors: 8\nlearning_rate: 0.010000\n" in out
    tr_lines = re.findall(r"^TRAIN: Iteration (\d+) GPU MAE: ([0-9.]+) RMSE: ([0-9.]+)$", out, re.M)
    te_lines = re.findall(r"^TEST: Iteration (\d+) GPU MAE: ([0-9.]+) RMSE: ([0-9.]+)$", out, re.M)
    assert [int(x[0]) for x in tr_lines] == [int(x[0]) for x in te_lines] == [1, 120]
    assert re.search(r"^Time taken for 120 of iterations is [0-9.]+$", out, re.M)
    assert float(te_lines[-1][2]) < float(te_lines[0][2])
    # the five component files next to the training file (mf.cu:65-87, util.cu:99-103)
    U, I, k = 400, int(max(tr["item"].max(), te["item"].max())) + 1, 8
    shapes = {"p": (U, k), "q": (I, k), "user_bias": (U, 1), "item_bias": (I, 1), "global_bias": (1, 1)}
    for comp, (r, c) in shapes.items():
        path = tmp_path / ("train_f8_%s.csv" % comp)
        lines = path.read_text().splitlines()
        assert len(lines) == r and all(len(l.split(",")) == c for l in lines), comp
        assert all(re.fullmatch(r"-?\d+\.\d{6}", v) for v in lines[0].split(","))
    gb = float((tmp_path / "train_f8_global_bias.csv").read_text())
    assert abs(gb - tr["rating"].astype(np.float64).mean()) < 1e-5
    # the unmodified reference predict consumes OUR files; ours consumes them too
    _write_csv(tmp_path / "user.csv", np.array([(0, 3, 5.0), (0, 10, 1.0), (0, 40, 4.0)], dtype=cu.RATING_DTYPE))
    (tmp_path / "p.cfg").write_text("0 50 8 0.05 42 0.02 0.02 0.02 0.02")
    args = ["-c", str(tmp_path / "p.cfg"), "-i", str(tmp_path / "train_f8_item_bias.csv"), "-g",
            str(tmp_path / "train_f8_global_bias.csv"), "-q", str(tmp_path / "train_f8_q.csv"), str(tmp_path / "user.csv")]
    ours = subprocess.run([PREDICT, *args], capture_output=True, text=True)
    assert ours.returncode == 0, ours.stderr
    recs = re.findall(r"^Rank: (\d+)\tItem: (\d+)\tEstimated rating: (-?[0-9.]+)$", ours.stdout, re.M)
    assert len(recs) == I - 3 and [int(r[0]) for r in recs] == list(range(1, I - 2))
    assert {3, 10, 40}.isdisjoint({int(r[1]) for r in recs})
    est = [float(r[2]) for r in recs]
    assert est == sorted(est, reverse=True)
    # extension: every user of the trained model at once through the tcgen05 batched predict
    allu = subprocess.run([PREDICT, "-c", str(tmp_path / "c.cfg"), "-i", str(tmp_path / "train_f8_item_bias.csv"), "-g",
                           str(tmp_path / "train_f8_global_bias.csv"), "-q", str(tmp_path / "train_f8_q.csv"), "-p",
                           str(tmp_path / "train_f8_p.csv"), "-u", str(tmp_path / "train_f8_user_bias.csv"), "-k", "5", "-x",
                           str(tmp_path / "train.csv")], capture_output=True, text=True)
    assert allu.returncode == 0, allu.stderr
    rows = re.findall(r"^User: (\d+)\tRank: (\d+)\tItem: (\d+)\tEstimated rating: (-?[0-9.]+)$", allu.stdout, re.M)
    assert len(rows) == U * 5 and [int(r[1]) for r in rows[:5]] == [1, 2, 3, 4, 5]
    seen = set(zip(tr["user"].tolist(), tr["item"].tolist()))
    assert not any((int(u), int(i)) in seen for u, _, i, _ in rows)
    Pm, Qm = np.loadtxt(tmp_path / "train_f8_p.csv", delimiter=",", dtype=np.float32), np.loadtxt(tmp_path / "train_f8_q.csv", delimiter=",", dtype=np.float32)
    ubm, ibm = np.loadtxt(tmp_path / "train_f8_user_bias.csv", dtype=np.float32), np.loadtxt(tmp_path / "train_f8_item_bias.csv", dtype=np.float32)
    ex = cu.createSparseMatrix(tr, U, I)
    want_i, _ = O.predict_topk(Pm, Qm, ubm, ibm, np.float32(gb), 5, exclude=(ex.indptr, ex.indices))
    assert [int(r[2]) for r in rows] == want_i.ravel().tolist()
    if O.ref_binary("predict"):
        ref = subprocess.run([O.ref_binary("predict"), *args], capture_output=True, text=True)
        assert ref.returncode == 0 and "Recommendations:" in ref.stdout, ref.stderr[-300:]
        ref_recs = re.findall(r"^Rank: (\d+)\tItem: (\d+)\tEstimated rating: (-?[0-9.]+)$", ref.stdout, re.M)
        # same catalogue, same frozen-Q intent: the two top-20 lists overlap heavily
        top_ours, top_ref = {int(r[1]) for r in recs[:20]}, {int(r[1]) for r in ref_recs[:20]}
        assert len(top_ours & top_ref) >= 10, (sorted(top_ours), sorted(top_ref))


@pytest.mark.gpu
def test_mf_cli_without_config_uses_reference_defaults(tmp_path):
    tr, te = cu.synth_ratings(60, 40, 600, rank=2, noise=0.3, seed=4)
    _write_csv(tmp_path / "a.csv", tr)
    _write_csv(tmp_path / "b.csv", te)
    p = subprocess.run([MF, str(tmp_path / "a.csv"), str(tmp_path / "b.csv")], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "total_iterations: 5000\nn_factors: 50\n" in p.stdout  # config.h:25-27
    its = [int(x) for x in re.findall(r"^TEST: Iteration (\d+)", p.stdout, re.M)]
    assert its == [1] + list(range(500, 5001, 500))
    assert os.path.exists(tmp_path / "a_f50_p.csv")


@pytest.mark.gpu
def test_experiment_grid_harness_on_ml100k_shape(tmp_path):
    """experiments/run_grid.py (the reference's experiments/cu2rec.sh grid through bin/mf): one data set
    x two iteration counts x one factor count, synthetic data generated on the fly."""
    import json
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "experiments", "run_grid.py"), "--datasets", "ml-100k",
                        "--iterations", "50", "200", "--factors", "50", "--data-dir", str(tmp_path / "data"),
                        "--results-dir", str(tmp_path / "results"), "--tag", "t"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    rows = [json.loads(l) for l in open(tmp_path / "results" / "t.jsonl")]
    assert [(r["iterations"], r["factors"]) for r in rows] == [(50, 50), (200, 50)]
    for r in rows:
        assert r["dataset"] == "ml-100k" and r["users"] == 943 and r["items"] == 1682
        assert r["updates_per_s"] > 0 and 0.5 < r["test_rmse"] < 2.0 and r["wall_seconds"] >= r["train_seconds"]
    assert rows[1]["test_rmse"] < rows[0]["test_rmse"]  # more iterations, lower loss
    log = open(tmp_path / "results" / "t.txt").read()
    assert "Done with 50 factors with 200 iterations on ml-100k" in log and "TEST: Iteration" in log  # cu2rec.sh:17
    assert "| ml-100k (synthetic) | 200 |" in open(tmp_path / "results" / "t.md").read()
