"""Parity at the sizes BASELINE.json names ("at BASELINE.json's full sizes"), all on the B200 box:

  C2  MovieLens-20M shape, k = 64    : Hogwild on one GPU and DSGD logical ranks vs the sequential CPU oracle
  C3  Netflix shape, k = 128         : the same, checks at iterations 1 / 50 / 100
  C5  batched predict, Netflix shape : a user sample against the CPU brute force, bit for bit
  the regime in which round 1's 4-GPU run diverged, reproduced on one device (full occupancy on one of
  four item blocks): the default must stay finite and on the single-GPU trajectory, and the non-finite guard
  must fire when the stability bound is switched off
  bin/mf against the reference's own CPU binary (oracle/_ref/mf_cpu) on a MovieLens-small-shaped problem

The oracle trainer is sequential SGD with the same sampler stream and initialisation (oracle/mf_oracle.cpp restates
mf_sequential.cu:102-143); Hogwild differs from it by the interleaving of updates only, so the bar is the north
star's: test RMSE within 0.5 % at equal iterations."""
import os
import re
import subprocess
import threading

import numpy as np
import pytest

import bench
import cu2rec_b200 as cu
import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _env(pairs, fn):
    old = {k: os.environ.get(k) for k in pairs}
    os.environ.update({k: str(v) for k, v in pairs.items()})
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _logical_dsgd(tr, te, U, I, k, world, iters, ce, init):
    part = cu.dsgd_partition(tr, U, I, world)
    P, Q, ub, ib = init
    mu = np.float32(tr["rating"].astype(np.float64).sum() / len(tr))
    ranks = []
    for r in range(world):
        inp = cu.dsgd_rank_inputs(tr, te, U, I, part, r, P, Q, ub, ib)
        ranks.append(cu.Dsgd(r, world, inp, part, cu.Config(total_iterations=iters, n_factors=k, check_error=ce), mu))
    hs = [d.handle for d in ranks]
    for d in ranks:
        d.connect(hs)
    errs = []

    def work(d):
        try:
            d.run(iters)
        except Exception as e:  # pragma: no cover
            errs.append(e)
    th = [threading.Thread(target=work, args=(d,)) for d in ranks]
    [t.start() for t in th]
    [t.join(timeout=600) for t in th]
    assert not errs, errs
    lg = ranks[0].log()
    for d in ranks:
        d.close()
    return lg


def _against_oracle(workload, k, iters, ce, world):
    tr, te, U, I = bench.make_workload(workload)
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    mu = np.float32(tr["rating"].astype(np.float64).sum() / len(tr))
    init = tuple(cu.initialize_normal_array(n, k) for n in (U * k, I * k, U, I))
    box = {}
    th = threading.Thread(target=lambda: box.update(log=O.train(  # the CPU oracle runs while the GPU trains
        (mtr.indptr, mtr.indices, mtr.data), (mte.indptr, mte.indices, mte.data), *init, mu, O.hyper(k), 42, iters,
        check_error=ce)[-1]))
    th.start()
    with cu.Session(mtr, mte, cu.Config(total_iterations=iters, n_factors=k, check_error=ce), *init, mu) as s:
        s.run(iters)
        one = s.log()
    many = _logical_dsgd(tr, te, U, I, k, world, iters, ce, init)
    th.join()
    want = box["log"]
    assert [r["iteration"] for r in want] == [r["iteration"] for r in one] == [r["iteration"] for r in many]
    for name, got in (("one GPU", one), ("DSGD x%d" % world, many)):
        for g, w in zip(got, want):
            for key in ("test_rmse", "train_rmse"):
                assert np.isfinite(g[key]) and abs(g[key] - w[key]) / w[key] < 0.005, (name, key, g, w)
    print(workload, "oracle", [round(r["test_rmse"], 5) for r in want], "one GPU", [round(r["test_rmse"], 5) for r in one],
          "DSGD", [round(r["test_rmse"], 5) for r in many])


def test_c2_ml20m_shape_k64_matches_the_sequential_oracle():
    _against_oracle("ml20m", 64, 300, 100, world=2)


def test_c3_netflix_shape_k128_matches_the_sequential_oracle():
    _against_oracle("netflix", 128, 100, 50, world=4)


def test_c5_batched_predict_full_size_sample_is_bit_exact():
    k, topk = 128, 10
    tr, te, U, I = bench.make_workload("netflix")
    mtr = cu.createSparseMatrix(tr, U, I)
    mu = np.float32(tr["rating"].astype(np.float64).sum() / len(tr))
    init = lambda n: cu.initialize_normal_array(n, k)
    cfg = cu.Config(total_iterations=200, n_factors=k, check_error=200)
    with cu.Session(mtr, cu.createSparseMatrix(te, U, I), cfg, init(U * k), init(I * k), init(U), init(I), mu) as s:
        s.run(200)  # a trained-looking model so that the scores are not pure noise
        P, Q, ub, ib = s.download()
    items, scores, ms = cu.predict_topk(P, Q, ub, ib, mu, topk, exclude=mtr)
    sample = np.sort(np.random.RandomState(0).choice(U, 192, replace=False))
    sub_ptr = np.concatenate([[0], np.cumsum([mtr.indptr[u + 1] - mtr.indptr[u] for u in sample])]).astype(np.int32)
    sub_idx = np.concatenate([mtr.indices[mtr.indptr[u]:mtr.indptr[u + 1]] for u in sample]).astype(np.int32)
    wi, ws = O.predict_topk(P[sample], Q, ub[sample], ib, mu, topk, exclude=(sub_ptr, sub_idx))
    assert np.array_equal(items[sample], wi)
    assert np.array_equal(scores[sample].view(np.uint32), ws.view(np.uint32))
    assert ms["candidates_ms"] < 50.0  # the whole 480 189 x 17 770 sweep (measured: ~6 ms)


def test_four_gpu_regime_on_one_device_stays_finite_and_the_guard_fires_without_the_bound():
    """One of four item blocks of the Netflix shape (all users, 4 442 items: the most popular item takes ~1 % of the
    draws), every user group the device holds in flight: lr x share x groups is ~1, four times the measured bound
    (kStableLoad). Round 1 ran this regime with an in-flight cap tuned to a load of 0.5 and diverged on the
    driver's 4-GPU box with rc 0. Now: (a) the default (bias steps of the popular items thinned to the bound) stays
    finite and within 1 % of the capped run that applies every step; (b) with the bound switched off the
    device-side guard turns the NaN into CU2B_ERR_DIVERGED."""
    k, iters = 128, 256
    tr, te, U, I = bench.make_workload("nfblock4")
    mu = np.float32(tr["rating"].astype(np.float64).sum() / len(tr))
    init = tuple(cu.initialize_normal_array(n, k) for n in (U * k, I * k, U, I))
    part = cu.dsgd_partition(tr, U, I, 1)
    inp = cu.dsgd_rank_inputs(tr, te, U, I, part, 0, *init)

    def run(env, lr=0.01):
        def go():
            d = cu.Dsgd(0, 1, inp, part, cu.Config(total_iterations=2 * iters, n_factors=k, check_error=iters,
                                                   learning_rate=lr), mu)
            d.connect([d.handle])
            try:
                d.run(2 * iters)
                return d.log(), d.stats()
            finally:
                d.close()
        return _env(env, go)
    lg, st = run({})
    assert all(np.isfinite(r["test_rmse"]) and np.isfinite(r["train_rmse"]) for r in lg)
    # the exact comparator: every item-side step applied, user groups capped at a quarter of the bound's load
    capped, _ = run({"CU2B_DSGD_THIN_BIAS": "0", "CU2B_INFLIGHT_LR": "0.125"})
    # (the two runs differ in how many user groups interleave, which alone moves the RMSE of this 512-iteration run
    # by 0.2-0.4 % between grid sizes -- profiles/r2_dsgd_stability_map.md section 1 -- so the bar here is 1 %; the
    # north star's 0.5 % against ONE GPU over 4 000 iterations is what the multi-GPU bench lines show: <= 0.16 %)
    for a, b, tol in zip(lg, capped, (0.005, 0.01, 0.01)):
        assert abs(a["test_rmse"] - b["test_rmse"]) / b["test_rmse"] < tol, (a, b)
    assert st["updates"] == 2 * iters * U
    # no thinning, no cap. At lr = 0.01 this is a load of ~0.5-1, the chaotic edge: the same run diverges in most
    # calls and trains in some (profiles/r2_dsgd_stability_map.md), so the certain case is used here: lr = 0.08 is a
    # load of >= 4, far beyond the delayed-gradient limit of pi / 2.
    with pytest.raises(cu._lib.Cu2bError) as err:
        run({"CU2B_DSGD_THIN_BIAS": "0", "CU2B_INFLIGHT_LR": "0"}, lr=0.08)
    assert err.value.status == 6 and "non-finite" in str(err.value)


def _write_csv(path, r):
    with open(path, "w") as f:
        f.write("userId,itemId,rating\n")
        f.write("".join("%d,%d,%.1f\n" % (u + 1, i + 1, x) for u, i, x in r))


@pytest.mark.skipif(O.ref_binary("mf_cpu") is None, reason="oracle/_ref not built")
def test_bin_mf_matches_the_reference_cpu_binary_on_a_movielens_small_shape(tmp_path):
    """`bin/mf -c cfg train.csv test.csv` against the UNMODIFIED reference mf_cpu on the same files: the shape of the
    data set the reference bundles (610 users x 9 724 items x 100 836 ratings, 80 / 20 split), k = 32, 1000
    iterations (the bundled file itself is not redistributable and /root/reference does
    not exist on the GPU box). mf_cpu seeds from std::random_device: its own run-to-run spread is ~0.1 %."""
    tr, te = cu.synth_ratings(610, 9724, 100836, rank=8, noise=0.4, integer_ratings=False, test_fraction=0.2, seed=11)
    # mf_sequential.cu:111 draws from the INCLUSIVE range [low, high]: with probability 1 / (degree + 1) per iteration
    # the last user reads one element past the rating arrays and the reference dies on what it finds there. Give the
    # last user 4 000 more ratings so that a 1000-iteration run usually survives (O.run_mf_cpu retries the rest).
    rng = np.random.RandomState(11)
    last = tr[tr["user"] == 609]
    extra_items = np.setdiff1d(np.arange(9724), np.concatenate([last["item"], te[te["user"] == 609]["item"]]))[:4000]
    extra = np.zeros(len(extra_items), dtype=cu.RATING_DTYPE)
    extra["user"], extra["item"], extra["rating"] = 609, extra_items, rng.randint(1, 11, len(extra_items)) * 0.5
    tr = np.concatenate([tr, extra])
    tr = tr[np.lexsort((tr["item"], tr["user"]))]
    _write_csv(tmp_path / "train.csv", tr)
    _write_csv(tmp_path / "test.csv", te)
    (tmp_path / "c.cfg").write_text("0 1000 32 0.01 42 0.02 0.02 0.02 0.02")
    ours = subprocess.run([os.path.join(ROOT, "bin", "mf"), "-c", str(tmp_path / "c.cfg"), str(tmp_path / "train.csv"),
                           str(tmp_path / "test.csv")], capture_output=True, text=True)
    assert ours.returncode == 0, ours.stderr
    ref_out = O.run_mf_cpu(tmp_path / "c.cfg", tmp_path / "train.csv", tmp_path / "test.csv")
    last = lambda out: float([l for l in out.splitlines() if l.startswith("TEST:")][-1].split()[-1])
    its = lambda out: [int(m) for m in re.findall(r"^TEST: Iteration (\d+)", out, re.M)]
    assert its(ours.stdout) == [1, 500, 1000]
    a, b = last(ours.stdout), last(ref_out)
    print("bin/mf", a, "reference mf_cpu", b)
    assert abs(a - b) / b < 0.005, (a, b)
