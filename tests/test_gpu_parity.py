"""GPU parity tests (run on the B200 box with -m gpu). Every call goes through the C ABI of
libcu2b.so; the CPU oracle (tests/oracle.py) and the prebuilt unmodified reference binaries
(oracle/_ref, compiled for sm_100a) are the checkers.

Bars: bit-exact for integer / index work (sampler, CSR) and for the deterministic replay of the
update arithmetic against the oracle's KERNEL flavour; fp32 tolerances, stated per test, for
comparisons against the reference's op order; 1e-5 relative for the loss; 0.5 % for final RMSE."""
import json
import os
import subprocess

import numpy as np
import pytest

import cu2rec_b200 as cu
import oracle as O

pytestmark = pytest.mark.gpu

KS = [1, 2, 3, 7, 16, 32, 50, 64, 100, 128, 200, 256, 300, 512]


def _model(rng, U, I, k, scale=None):
    s = scale if scale is not None else 1.0 / np.sqrt(k)
    P = (rng.standard_normal((U, k)) * s).astype(np.float32)
    Q = (rng.standard_normal((I, k)) * s).astype(np.float32)
    ub = (rng.standard_normal(U) * 0.1).astype(np.float32)
    ib = (rng.standard_normal(I) * 0.1).astype(np.float32)
    return P, Q, ub, ib


def _random_matrix(rng, U, I, max_deg, empty_frac=0.1):
    deg = rng.randint(1, max_deg + 1, U)
    deg[rng.rand(U) < empty_frac] = 0
    rows = []
    for u in range(U):
        for i in rng.choice(I, min(deg[u], I), replace=False):
            rows.append((u, i, float(rng.randint(1, 11)) / 2))
    r = np.array(rows, dtype=cu.RATING_DTYPE)
    return r, cu.createSparseMatrix(r, U, I)


# ---------------------------------------------------------------------------------------------
# loss (loss.cu)
# ---------------------------------------------------------------------------------------------
def test_loss_golden_74(fixtures_dir):
    # tests/test_loss.cu:23-90
    r, rows, cols, _ = cu.readCSV(os.path.join(fixtures_dir, "test_ratings.csv"))
    m = cu.createSparseMatrix(r, rows, cols)
    k = 2
    P, Q = np.ones((rows, k), np.float32), np.ones((cols, k), np.float32)
    ub, ib = np.ones(rows, np.float32), np.ones(cols, np.float32)
    err = cu.calculate_loss_gpu(P, Q, k, m, ub, ib, 1.0)
    loss = np.float32(0)
    for e in err:
        loss = np.float32(loss + e * e)
    assert loss == 74.0
    mae, rmse = cu.loss(P, Q, k, m, ub, ib, 1.0)
    assert rmse == np.float32(np.sqrt(74.0 / 18)) and mae == np.float32(np.abs(err).astype(np.float64).sum() / 18)


@pytest.mark.parametrize("n", [1, 33, 1 << 10, 1 << 16, (1 << 20) + 7])
def test_total_loss_all_ones(n):
    # tests/test_loss.cu:106-147 (the grid/block sweep is a launch detail of the reference kernel)
    mae, rmse = cu.get_error_metrics_gpu(np.ones(n, np.float32))
    assert mae == 1 and rmse == 1


@pytest.mark.parametrize("k", KS)
def test_residuals_and_metrics_vs_oracle(k):
    rng = np.random.RandomState(100 + k)
    U, I = 300, 200
    r, m = _random_matrix(rng, U, I, 40)
    P, Q, ub, ib = _model(rng, U, I, k)
    mu = 3.3
    err = cu.calculate_loss_gpu(P, Q, k, m, ub, ib, mu)
    want_k = O.residuals(m.indptr, m.indices, m.data, P, Q, ub, ib, mu, k, O.FLAVOUR_LOSS_KERNEL)
    assert err.view(np.uint32).tolist() == want_k.view(np.uint32).tolist()  # same op order => same bits
    want_r = O.residuals(m.indptr, m.indices, m.data, P, Q, ub, ib, mu, k, O.FLAVOUR_REF)
    np.testing.assert_allclose(err, want_r, rtol=0, atol=2e-5)  # reference op order: fp32 rounding only
    mae, rmse = cu.loss(P, Q, k, m, ub, ib, mu)
    omae, ormse, _, _ = O.loss(m.indptr, m.indices, m.data, P, Q, ub, ib, mu, k, O.FLAVOUR_REF)
    assert abs(mae - omae) / omae < 1e-5 and abs(rmse - ormse) / ormse < 1e-5
    mae2, rmse2 = cu.get_error_metrics_gpu(err)
    assert mae2 == mae and rmse2 == rmse
    assert (mae, rmse) == cu.loss(P, Q, k, m, ub, ib, mu)  # bitwise reproducible


def test_loss_large_ragged_matrix_vs_oracle():
    rng = np.random.RandomState(9)
    tr, _ = cu.synth_ratings(20000, 3000, 1500000, integer_ratings=True)
    m = cu.createSparseMatrix(tr, 20000, 3000)
    k = 128
    P, Q, ub, ib = _model(rng, 20000, 3000, k)
    mae, rmse = cu.loss(P, Q, k, m, ub, ib, 3.6)
    omae, ormse, _, _ = O.loss(m.indptr, m.indices, m.data, P, Q, ub, ib, 3.6, k, O.FLAVOUR_REF)
    assert abs(mae - omae) / omae < 1e-5 and abs(rmse - ormse) / ormse < 1e-5


@pytest.mark.skipif(O.ref_binary("ref_harness") is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("k", [2, 32, 50, 128])
def test_loss_vs_reference_gpu_kernels(tmp_path, k):
    """Our fused loss vs the UNMODIFIED reference loss_kernel + total_loss_kernel (loss.cu) running
    on this GPU: 1e-5 relative on mae / rmse, fp32 rounding on the residuals."""
    rng = np.random.RandomState(200 + k)
    U, I = 500, 300
    r, m = _random_matrix(rng, U, I, 60)
    P, Q, ub, ib = _model(rng, U, I, k)
    mu = np.float32(3.4)
    with open(tmp_path / "in.bin", "wb") as f:
        np.array([U, I, m.nonzeros, k], np.int32).tofile(f)
        np.array([mu], np.float32).tofile(f)
        for a in (m.indptr, m.indices, m.data, P, Q, ub, ib):
            a.tofile(f)
    subprocess.run([O.ref_binary("ref_harness"), "loss_gpu", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")],
                   check=True, capture_output=True)
    out = np.fromfile(tmp_path / "out.bin", dtype=np.float32)
    ref_err, ref_mae, ref_rmse = out[:-2], out[-2], out[-1]
    err = cu.calculate_loss_gpu(P, Q, k, m, ub, ib, mu)
    np.testing.assert_allclose(err, ref_err, rtol=0, atol=2e-5)
    mae, rmse = cu.loss(P, Q, k, m, ub, ib, mu)
    assert abs(mae - ref_mae) / ref_mae < 1e-5 and abs(rmse - ref_rmse) / ref_rmse < 1e-5


# ---------------------------------------------------------------------------------------------
# sampler (sgd.cu:27-37)
# ---------------------------------------------------------------------------------------------
def test_sampler_bit_exact_vs_oracle():
    rng = np.random.RandomState(1)
    r, m = _random_matrix(rng, 700, 500, 30, empty_frac=0.2)
    for seed, it0, n in [(42, 0, 5), (7, 1000, 3), (-3, 2 ** 20, 2)]:
        got = cu.sample_per_user(m, seed, it0, n)
        want = O.sample_per_user(m.indptr, m.indices, m.data, seed, it0, n)
        assert got.tobytes() == want.tobytes()
    assert len(cu.sample_per_user(m, 1, 0, 0)) == 0


# ---------------------------------------------------------------------------------------------
# update arithmetic (sgd.cu:40-72 / mf_sequential.cu:114-141)
# ---------------------------------------------------------------------------------------------
def _stream(rng, U, I, n):
    s = np.zeros(n, dtype=cu.RATING_DTYPE)
    s["user"], s["item"] = rng.randint(0, U, n), rng.randint(0, I, n)
    s["rating"] = rng.randint(1, 6, n)
    return s


@pytest.mark.parametrize("k", KS)
def test_serial_replay_bit_exact_vs_oracle(k):
    rng = np.random.RandomState(300 + k)
    U, I, n = 50, 40, 1500  # heavy reuse of rows: every update depends on earlier ones
    P, Q, ub, ib = _model(rng, U, I, k)
    s = _stream(rng, U, I, n)
    cfg = cu.Config(n_factors=k, learning_rate=0.02, P_reg=0.03, Q_reg=0.04, user_bias_reg=0.05, item_bias_reg=0.06)
    got = cu.sgd_apply(s, P, Q, ub, ib, 3.5, cfg, order=1)
    want = O.sgd_apply_stream(s, P.ravel(), Q.ravel(), ub, ib, 3.5, O.hyper_from_cfg(cfg), O.FLAVOUR_KERNEL)
    for g, w in zip(got, want):
        assert g.ravel().view(np.uint32).tolist() == w.view(np.uint32).tolist()
    ref = O.sgd_apply_stream(s, P.ravel(), Q.ravel(), ub, ib, 3.5, O.hyper_from_cfg(cfg), O.FLAVOUR_REF)
    for g, w in zip(got, ref):  # mf_sequential op order: only the dot-product order differs
        np.testing.assert_allclose(g.ravel(), w, rtol=0, atol=5e-5)


def test_single_update_known_answer():
    # the tests/test_sgd.cu setup, pinned numerically
    f = np.float32
    mu = f(64.0 / 18.0)
    cfg = cu.Config(n_factors=1, learning_rate=0.07, P_reg=0.1, Q_reg=0.1, user_bias_reg=0.1, item_bias_reg=0.1)
    for r in (1.0, 3.0, 5.0):
        s = np.array([(0, 0, r)], dtype=cu.RATING_DTYPE)
        P, Q, ub, ib = cu.sgd_apply(s, [1], [1], [1], [1], mu, cfg, order=0)
        err = f(r) - f(f(f(mu + f(1)) + f(1)) + f(1))
        want = f(1) + f(0.07) * f(err - f(f(0.1) * f(1)))
        assert [P[0], Q[0], ub[0], ib[0]] == [want] * 4


@pytest.mark.parametrize("k", [1, 7, 32, 50, 128, 256, 300])
def test_hogwild_conflict_free_stream_bit_exact(k):
    # distinct users and items => no races => the parallel kernel must equal the sequential oracle
    rng = np.random.RandomState(400 + k)
    n = 5000
    P, Q, ub, ib = _model(rng, n, n, k)
    s = np.zeros(n, dtype=cu.RATING_DTYPE)
    s["user"], s["item"] = rng.permutation(n), rng.permutation(n)
    s["rating"] = rng.randint(1, 6, n)
    cfg = cu.Config(n_factors=k, learning_rate=0.05)
    got = cu.sgd_apply(s, P, Q, ub, ib, 3.5, cfg, order=0)
    want = O.sgd_apply_stream(s, P.ravel(), Q.ravel(), ub, ib, 3.5, O.hyper_from_cfg(cfg), O.FLAVOUR_KERNEL)
    for g, w in zip(got, want):
        assert g.ravel().view(np.uint32).tolist() == w.view(np.uint32).tolist()


def test_hogwild_is_train_false_freezes_items():
    rng = np.random.RandomState(5)
    P, Q, ub, ib = _model(rng, 100, 80, 16)
    s = _stream(rng, 100, 80, 3000)
    cfg = cu.Config(n_factors=16, is_train=0)
    P2, Q2, ub2, ib2 = cu.sgd_apply(s, P, Q, ub, ib, 3.5, cfg, order=0)
    assert np.array_equal(Q2.reshape(Q.shape), Q) and np.array_equal(ib2, ib)
    assert not np.array_equal(P2.reshape(P.shape), P)


def test_hogwild_with_item_conflicts_stays_close_to_sequential():
    """The production pattern: every user appears once per segment (per-user sampling), items
    collide heavily (600 items, 40000 concurrent updates). Item-side steps are applied with L2
    atomic adds, so no step is lost; only the staleness of the rows read differs from the
    sequential replay => statistical closeness (racy by design, like the reference kernel)."""
    rng = np.random.RandomState(6)
    U, I, k, n = 40000, 600, 32, 40000
    P, Q, ub, ib = _model(rng, U, I, k)
    s = np.zeros(n, dtype=cu.RATING_DTYPE)
    s["user"], s["item"], s["rating"] = rng.permutation(U), rng.randint(0, I, n), rng.randint(1, 6, n)
    cfg = cu.Config(n_factors=k, learning_rate=0.01)
    got = cu.sgd_apply(s, P, Q, ub, ib, 3.5, cfg, order=0)
    want = O.sgd_apply_stream(s, P.ravel(), Q.ravel(), ub, ib, 3.5, O.hyper_from_cfg(cfg), O.FLAVOUR_KERNEL)
    start = (P.ravel(), Q.ravel(), ub, ib)
    for g, w, s0 in zip(got, want, start):
        assert np.all(np.isfinite(g))
        moved = np.sqrt(np.mean((w - s0) ** 2))
        assert np.sqrt(np.mean((g.ravel() - w) ** 2)) < 0.35 * moved  # much closer to the replay than to the start
    empty = cu.sgd_apply(s[:0], P, Q, ub, ib, 3.5, cfg, order=0)
    assert np.array_equal(empty[0].reshape(P.shape), P)


@pytest.mark.skipif(O.ref_binary("ref_harness") is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("k", [1, 8, 50])
def test_update_vs_reference_sgd_kernel(tmp_path, k):
    """One launch of the UNMODIFIED reference sgd_update (sgd.cu:22-75) on an input where it is
    deterministic: every active user has exactly one rating, all items distinct, and the first 32
    users are empty so the grid's surplus threads (sgd.cu:27-28) hit no rating. Its
    P / Q_target / biases must match our update arithmetic to fp32 rounding (nvcc contracts the
    reference's expressions into FMAs; ours are unfused like mf_sequential)."""
    rng = np.random.RandomState(500 + k)
    U, I = 32 + 200, 200
    r = np.zeros(200, dtype=cu.RATING_DTYPE)
    r["user"], r["item"] = 32 + np.arange(200), rng.permutation(200)
    r["rating"] = rng.randint(1, 6, 200)
    m = cu.createSparseMatrix(r, U, I)
    P, Q, ub, ib = _model(rng, U, I, k)
    mu, lr, regs = np.float32(3.5), np.float32(0.07), np.float32([0.1, 0.05, 0.02, 0.03])
    with open(tmp_path / "in.bin", "wb") as f:
        np.array([U, I, m.nonzeros, k], np.int32).tofile(f)
        np.array([mu], np.float32).tofile(f)
        for a in (m.indptr, m.indices, m.data, P, Q, ub, ib):
            a.tofile(f)
        np.array([lr, *regs], np.float32).tofile(f)
        np.array([1, 0], np.int32).tofile(f)
    subprocess.run([O.ref_binary("ref_harness"), "sgd_gpu", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")],
                   check=True, capture_output=True)
    out = np.fromfile(tmp_path / "out.bin", dtype=np.float32)
    rP, rQ, rub, rib = np.split(out, np.cumsum([U * k, I * k, U]))
    cfg = cu.Config(n_factors=k, learning_rate=float(lr), P_reg=float(regs[0]), Q_reg=float(regs[1]),
                    user_bias_reg=float(regs[2]), item_bias_reg=float(regs[3]))
    got = cu.sgd_apply(r, P, Q, ub, ib, mu, cfg, order=0)
    for g, w in zip(got, (rP, rQ, rub, rib)):
        np.testing.assert_allclose(g.ravel(), w, rtol=0, atol=2e-6)


# ---------------------------------------------------------------------------------------------
# deterministic conflict-free mode vs the mf_sequential update on the same ordering
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,B", [(1, 4), (8, 16), (32, 7), (50, 64), (128, 32), (256, 16), (300, 8)])
def test_blocked_mode_bit_exact_vs_sequential_replay(k, B):
    """cu2b_sgd_blocked == sequential replay of the same ratings in the schedule's canonical order
    with the oracle's KERNEL flavour, bit for bit (two passes, heavy row reuse); and within a
    stated fp32 tolerance (5e-5 abs) of the mf_sequential.cu:114-141 op order (REF flavour)."""
    rng = np.random.RandomState(600 + k)
    U, I, n = 300, 200, 20000
    coo = np.zeros(n, dtype=cu.RATING_DTYPE)
    coo["user"], coo["item"], coo["rating"] = np.sort(rng.randint(0, U, n)), rng.randint(0, I, n), rng.randint(1, 6, n)
    P, Q, ub, ib = _model(rng, U, I, k)
    cfg = cu.Config(n_factors=k, learning_rate=0.02, P_reg=0.03, Q_reg=0.04, user_bias_reg=0.05, item_bias_reg=0.06)
    got = cu.sgd_blocked(coo, P, Q, ub, ib, 3.5, cfg, B, n_passes=2)
    order = O.block_schedule_order(coo, U, I, B)
    replay = np.concatenate([coo[order], coo[order]])
    want = O.sgd_apply_stream(replay, P.ravel(), Q.ravel(), ub, ib, 3.5, O.hyper_from_cfg(cfg), O.FLAVOUR_KERNEL)
    for g, w in zip(got, want):
        assert g.ravel().view(np.uint32).tolist() == w.view(np.uint32).tolist()
    ref = O.sgd_apply_stream(replay, P.ravel(), Q.ravel(), ub, ib, 3.5, O.hyper_from_cfg(cfg), O.FLAVOUR_REF)
    for g, w in zip(got, ref):
        np.testing.assert_allclose(g.ravel(), w, rtol=0, atol=5e-5)
    again = cu.sgd_blocked(coo, P, Q, ub, ib, 3.5, cfg, B, n_passes=2)
    for g, a in zip(got, again):
        assert np.array_equal(g, a)  # run-to-run reproducible


def test_blocked_mode_training_is_reproducible_and_converges():
    tr, te, mtr, mte, mu = _small_problem(U=800, I=300, n=40000)
    outs = []
    for _ in range(2):
        cfg = cu.Config(total_iterations=200, n_factors=16, check_error=50, mode=cu.MODE_DETERMINISTIC, n_blocks=32)
        outs.append(cu.train(mtr, mte, cfg, mu))
    a, b = outs
    assert [r["test_rmse"] for r in a["log"]] == [r["test_rmse"] for r in b["log"]]
    assert np.array_equal(a["P"], b["P"]) and np.array_equal(a["Q"], b["Q"])
    assert a["log"][-1]["test_rmse"] < a["log"][0]["test_rmse"]
    assert a["stats"]["updates"] == (200 * 800 // mtr.nonzeros) * mtr.nonzeros  # whole passes only


# ---------------------------------------------------------------------------------------------
# training loop (training.cu)
# ---------------------------------------------------------------------------------------------
def test_training_loop_reference_test(fixtures_dir):
    # tests/test_training.cu:21-45
    r, rows, cols, gb = cu.readCSV(os.path.join(fixtures_dir, "test_ratings.csv"))
    m = cu.createSparseMatrix(r, rows, cols)
    cfg = cu.Config(total_iterations=10, seed=42, n_factors=2, learning_rate=1e-3, P_reg=0.1, Q_reg=0.1,
                    user_bias_reg=0.1, item_bias_reg=0.1)
    out = cu.train(m, m, cfg, gb)
    assert out["losses"][0] >= out["losses"][9]
    assert [row["iteration"] for row in out["log"]] == [1, 10]
    assert np.all(np.isnan(out["losses"][1:9]))
    assert cfg.cur_iterations == 10
    for name in ("P", "Q", "user_bias", "item_bias"):
        assert np.all(np.isfinite(out[name]))
    # A6: every array starts from mt19937(42) => epoch-0 state shares the oracle's init
    assert out["P"].shape == (rows, 2) and out["Q"].shape == (cols, 2)


def _small_problem(U=1500, I=400, n=60000, seed=11):
    tr, te = cu.synth_ratings(U, I, n, rank=4, noise=0.3, integer_ratings=True, seed=seed)
    return tr, te, cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I), np.float32(tr["rating"].astype(np.float64).mean())


@pytest.mark.parametrize("k,round_iters", [(8, 1), (32, 1), (50, 1), (8, 16), (32, 16), (50, 16), (128, 16), (32, 64)])
def test_training_rmse_parity_vs_oracle_trainer(k, round_iters):
    """Hogwild GPU training vs the sequential CPU restatement (pinned to the compiled mf_cpu in
    tests/test_oracle_pins.py) at equal iterations, same sampler stream, same init: final and
    TEST RMSE within 0.5 %, every intermediate check within 1.5 %."""
    tr, te, mtr, mte, mu = _small_problem()
    U, I = mtr.rows, mtr.cols
    iters, ce = 300, 100
    # round_iters = 1: the reference's iteration-synchronous order (mf_sgd_hogwild + ordering gate);
    # > 1: the iteration-tiled schedule (mf_sgd_user_tiles), same draws, different interleaving
    cfg = cu.Config(total_iterations=iters, n_factors=k, check_error=ce, round_iters=round_iters)
    out = cu.train(mtr, mte, cfg, mu)
    init = lambda n: O.init_normal(n, k)
    *_, olog = O.train((mtr.indptr, mtr.indices, mtr.data), (mte.indptr, mte.indices, mte.data), init(U * k), init(I * k),
                       init(U), init(I), mu, O.hyper(k), 42, iters, check_error=ce)
    assert [r["iteration"] for r in out["log"]] == [r["iteration"] for r in olog] == [1, 100, 200, 300]
    for g, w in zip(out["log"], olog):  # intermediate checks: 1.5 %, final: the 0.5 % bar
        for key in ("train_rmse", "test_rmse", "train_mae", "test_mae"):
            assert abs(g[key] - w[key]) / w[key] < 0.015, (key, g, w)
    for key in ("train_rmse", "test_rmse"):
        assert abs(out["log"][-1][key] - olog[-1][key]) / olog[-1][key] < 0.005, (key, out["log"][-1], olog[-1])
    assert out["log"][-1]["test_rmse"] < out["log"][0]["test_rmse"]
    assert out["stats"]["updates"] == iters * U


def _disjoint_items_problem(U, d, seed=5):
    """Every item is rated by exactly one user: users never touch a common row, so ANY Hogwild
    schedule must equal the sequential restatement bit for bit (sampler + update arithmetic)."""
    rng = np.random.RandomState(seed)
    deg = rng.randint(1, d + 1, U)
    deg[0] = 1  # a one-rating user: every draw repeats the item (look-ahead re-read path)
    n = int(deg.sum())
    r = np.zeros(n, dtype=cu.RATING_DTYPE)
    r["user"] = np.repeat(np.arange(U), deg)
    r["item"] = rng.permutation(n)
    r["rating"] = rng.randint(1, 6, n)
    order = np.lexsort((r["item"], r["user"]))
    return r[order], n


@pytest.mark.parametrize("k,round_iters,variant", [
    (128, 32, {}), (128, 32, {"CU2B_TUNE_PF": "1"}), (128, 7, {"CU2B_TUNE_MINB": "8"}), (64, 32, {}),
    (64, 48, {"CU2B_TUNE_PF": "1"}), (32, 32, {}), (50, 16, {"CU2B_TUNE_PF": "1"}), (8, 8, {}), (3, 4, {"CU2B_TUNE_PF": "1"}),
    (200, 32, {}), (256, 20, {"CU2B_TUNE_PF": "1"}), (128, 32, {"CU2B_TILE_PIPE": "tma"}), (32, 16, {"CU2B_TILE_PIPE": "tma"}),
    (128, 1, {})])
def test_user_major_schedules_bit_exact_on_disjoint_items(k, round_iters, variant, monkeypatch):
    """mf_sgd_user_rounds (sampler fused into the update kernel; with and without item-row
    look-ahead), mf_sgd_user_tiles (separate sampler + TMA tiles) and mf_sgd_hogwild (round 1)
    against the sequential CPU restatement, KERNEL flavour: identical bits in P, Q and both biases
    after 45 iterations (rounds of uneven length included)."""
    for name, val in variant.items():
        monkeypatch.setenv(name, val)
    U = 700
    tr, I = _disjoint_items_problem(U, 6)
    te = tr[::3].copy()
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    mu = np.float32(3.0)
    iters = 45
    init = lambda n: cu.initialize_normal_array(n, k)
    P, Q, ub, ib = init(U * k), init(I * k), init(U), init(I)
    cfg = cu.Config(total_iterations=iters, n_factors=k, check_error=20, round_iters=round_iters)
    with cu.Session(mtr, mte, cfg, P, Q, ub, ib, mu) as s:
        s.run(iters)
        got = s.download()
        glog = s.log()
    oP, oQ, oub, oib, olog = O.train((mtr.indptr, mtr.indices, mtr.data), (mte.indptr, mte.indices, mte.data), P, Q, ub, ib, mu,
                                     O.hyper(k), 42, iters, check_error=20, flavour=O.FLAVOUR_KERNEL)
    for g, w, name in zip(got, (oP, oQ, oub, oib), ("P", "Q", "user_bias", "item_bias")):
        assert np.array_equal(np.asarray(g).ravel().view(np.uint32), w.view(np.uint32)), name
    assert [r["iteration"] for r in glog] == [r["iteration"] for r in olog]


def test_training_schedule_on_device_follows_reference_rule():
    """training.cu:129,146-155 evaluated on the device: replaying the rule on the validation RMSE
    sequence the run itself logged must reproduce the logged learning rates exactly. The
    validation set holds the training pairs with inverted ratings, so fitting the training set
    makes validation worse and the schedule is guaranteed to fire."""
    tr, te, mtr, mte, mu = _small_problem(U=600, I=200, n=20000)
    inv = tr.copy()
    inv["rating"] = 6.0 - inv["rating"]
    minv = cu.createSparseMatrix(inv, mtr.rows, mtr.cols)
    cfg = cu.Config(total_iterations=200, n_factors=8, check_error=20, learning_rate=0.05, patience=2.0)
    out = cu.train(mtr, minv, cfg, mu)
    f = np.float32
    lr, patience, val = f(0.05), 2, np.finfo(np.float32).max
    for row in out["log"]:
        last, val = val, f(row["test_rmse"])
        if last < val:
            patience -= 1
        if patience <= 0:
            patience, lr = 2, f(lr * f(0.2))
        assert row["learning_rate"] == lr, row
    assert out["log"][-1]["learning_rate"] < f(0.05) * f(0.2) * 1.01  # fired at least once
    assert cfg.learning_rate == out["log"][-1]["learning_rate"] and cfg.cur_iterations == 200


@pytest.mark.parametrize("round_iters", [1, 16])
def test_session_resume_equals_single_run(round_iters):
    tr, te, mtr, mte, mu = _small_problem(U=500, I=150, n=15000)
    U, I, k = mtr.rows, mtr.cols, 16
    init = lambda n: cu.initialize_normal_array(n, k)
    P, Q, ub, ib = init(U * k), init(I * k), init(U), init(I)
    cfg = cu.Config(total_iterations=40, n_factors=k, check_error=10, round_iters=round_iters)
    with cu.Session(mtr, mte, cfg, P, Q, ub, ib, mu) as s:
        s.run(40)
        log_a, ev_a = s.log(), s.eval()
    with cu.Session(mtr, mte, cfg, P, Q, ub, ib, mu) as s:
        s.run(13); s.run(27)
        log_b, ev_b = s.log(), s.eval()
        assert s.config().cur_iterations == 40
    assert [r["iteration"] for r in log_a] == [r["iteration"] for r in log_b] == [1, 10, 20, 30, 40]
    for a, b in zip(log_a, log_b):
        assert abs(a["test_rmse"] - b["test_rmse"]) / a["test_rmse"] < 2e-3  # Hogwild: not bit-identical
    assert abs(ev_a["test_rmse"] - log_a[-1]["test_rmse"]) < 1e-6


def test_session_reload_equals_a_fresh_session():
    """cu2b_session_reload: after some training, reloading the same problem + initial model and
    running again gives the bits of a fresh session (disjoint-items problem: any Hogwild schedule
    is deterministic there); a different shape is refused."""
    U, k, iters = 500, 128, 40
    tr, I = _disjoint_items_problem(U, 5, seed=9)
    te = tr[::4].copy()
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    mu = np.float32(3.0)
    init = lambda n: cu.initialize_normal_array(n, k)
    P, Q, ub, ib = init(U * k), init(I * k), init(U), init(I)
    cfg = cu.Config(total_iterations=iters, n_factors=k, check_error=10)
    with cu.Session(mtr, mte, cfg, P, Q, ub, ib, mu) as s:
        s.run(iters)
        fresh, fresh_log = s.download(), s.log()
    with cu.Session(mtr, mte, cfg, P, Q, ub, ib, mu) as s:
        s.run(17)  # leave the session mid-way, with a moved model and a non-empty log
        s.reload(mtr, mte, P, Q, ub, ib, mu)
        assert s.config().cur_iterations == 0 and s.stats()["updates"] == 0
        s.run(iters)
        again, again_log = s.download(), s.log()
        assert again_log == fresh_log
        for a, b in zip(again, fresh):
            assert np.array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))
        other = cu.createSparseMatrix(tr[:-3], U, I)
        with pytest.raises(cu._lib.Cu2bError):
            s.reload(other, mte, P, Q, ub, ib, mu)


def test_train_edge_cases():
    # one user, one rating; users without ratings; empty test matrix is rejected only if oversized
    r = np.array([(2, 1, 4.0)], dtype=cu.RATING_DTYPE)
    m = cu.createSparseMatrix(r, 4, 3)
    cfg = cu.Config(total_iterations=5, n_factors=3, check_error=2)
    out = cu.train(m, m, cfg, 4.0)
    assert [row["iteration"] for row in out["log"]] == [1, 2, 4, 5] and out["stats"]["updates"] == 5
    big = cu.createSparseMatrix(np.array([(0, 5, 1.0)], dtype=cu.RATING_DTYPE), 1, 6)
    with pytest.raises(cu._lib.Cu2bError):
        cu.train(m, big, cu.Config(total_iterations=1, n_factors=3), 4.0)
    with pytest.raises(cu._lib.Cu2bError):
        cu.train(m, m, cu.Config(total_iterations=1, n_factors=600), 4.0)


@pytest.mark.skipif(O.ref_binary("mf") is None, reason="oracle/_ref not built")
def test_training_vs_reference_gpu_binary(tmp_path):
    """End to end against the UNMODIFIED reference `mf` (sm_100a build) on the same CSV + cfg.
    The reference GPU path drops every item update but the first per iteration (early-bird gate,
    sgd.cu:49-63) and ping-pongs Q (training.cu:164-165); with many more users than
    items it therefore converges far more slowly per iteration than its own CPU path
    (mf_sequential), which is the comparator the parity bar is defined on. Here the bar is
    one-sided: at equal iterations we must not be worse than the reference GPU binary."""
    tr, te, mtr, mte, mu = _small_problem(U=3000, I=500, n=150000, seed=5)
    def write(path, r):
        with open(path, "w") as f:
            f.write("userId,itemId,rating\n")
            f.write("".join("%d,%d,%.1f\n" % (u + 1, i + 1, x) for u, i, x in r))
    write(tmp_path / "train.csv", tr)
    write(tmp_path / "test.csv", te)
    (tmp_path / "c.cfg").write_text("0 600 16 0.01 42 0.02 0.02 0.02 0.02")
    outp = subprocess.run([O.ref_binary("mf"), "-c", str(tmp_path / "c.cfg"), str(tmp_path / "train.csv"),
                           str(tmp_path / "test.csv")], capture_output=True, text=True, check=True).stdout
    ref_final = float([l for l in outp.splitlines() if l.startswith("TEST:")][-1].split()[-1])
    cfg = cu.Config()
    cfg.read_config(tmp_path / "c.cfg")
    r2, rows, cols, gb = cu.readCSV(tmp_path / "train.csv")
    out = cu.train(mtr, mte, cfg, gb)
    assert np.isfinite(ref_final) and out["log"][-1]["test_rmse"] <= ref_final * 1.005, (out["log"][-1], ref_final)


@pytest.mark.skipif(O.ref_binary("test_loss") is None, reason="oracle/_ref not built")
def test_reference_own_tests_pass_on_this_gpu(tmp_path, fixtures_dir):
    """Sanity for the GPU oracle: the reference's assert-based test binaries run green here."""
    d = tmp_path / "data" / "test"
    d.mkdir(parents=True)
    for f in os.listdir(fixtures_dir):
        (d / f).write_bytes(open(os.path.join(fixtures_dir, f), "rb").read())
    cwd = tmp_path / "matrix_factorization" / "tests"
    cwd.mkdir(parents=True)
    for name in ("test_loss", "test_sgd", "test_training"):
        p = subprocess.run([O.ref_binary(name)], cwd=cwd, capture_output=True, text=True)
        assert p.returncode == 0 and "PASSED" in p.stdout, (name, p.stdout[-300:], p.stderr[-300:])


# ---------------------------------------------------------------------------------------------
# full-size, size-independent properties (MovieLens-20M shape, k=64; Netflix shape in bench)
# ---------------------------------------------------------------------------------------------
def test_full_size_properties_ml20m_shape():
    U, I, k = 138493, 26744, 64
    tr, te = cu.synth_ratings(U, I, 20000263, integer_ratings=False)
    assert abs(len(tr) + len(te) - 20000263) < 0.01 * 20000263
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    mu = np.float32(tr["rating"].astype(np.float64).mean())
    init = lambda n: cu.initialize_normal_array(n, k)
    P, Q, ub, ib = init(U * k), init(I * k), init(U), init(I)
    # (1) sampler: integer work, every draw is one of the user's own ratings, one per user
    s = cu.sample_per_user(mtr, 42, 17, 2)
    assert s["user"].reshape(2, U).tolist() == [list(range(U))] * 2
    key_all = set((tr["user"].astype(np.int64) * I + tr["item"]).tolist())
    assert set((s["user"].astype(np.int64) * I + s["item"]).tolist()) <= key_all
    # (2) loss: linearity over a split of the matrix (checksum of checksums) and reproducibility
    mae, rmse = cu.loss(P, Q, k, mtr, ub, ib, mu)
    half = int(mtr.indptr[U // 2])
    a = cu.createSparseMatrix(tr[:half], U, I)
    b = cu.createSparseMatrix(tr[half:], U, I)
    (mae_a, rmse_a), (mae_b, rmse_b) = cu.loss(P, Q, k, a, ub, ib, mu), cu.loss(P, Q, k, b, ub, ib, mu)
    na, nb = half, len(tr) - half
    assert abs((mae_a * na + mae_b * nb) / (na + nb) - mae) / mae < 1e-6
    assert abs(np.sqrt((rmse_a ** 2 * na + rmse_b ** 2 * nb) / (na + nb)) - rmse) / rmse < 1e-6
    assert (mae, rmse) == cu.loss(P, Q, k, mtr, ub, ib, mu)
    # (3) training: loss goes down, model stays finite, update count is iterations x users; the
    #     tiled schedule lands within 0.5 % of the iteration-synchronous one at equal iterations
    finals = {}
    for r_it in (1, 16):
        cfg = cu.Config(total_iterations=300, n_factors=k, check_error=100, round_iters=r_it)
        out = cu.train(mtr, mte, cfg, mu)
        rm = [r["test_rmse"] for r in out["log"]]
        assert rm[-1] < rm[0] and np.isfinite(rm).all()
        assert out["stats"]["updates"] == 300 * U
        assert np.all(np.isfinite(out["P"])) and np.all(np.isfinite(out["Q"]))
        finals[r_it] = rm[-1]
    assert abs(finals[16] - finals[1]) / finals[1] < 0.005


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs as parity cases
# ---------------------------------------------------------------------------------------------
@pytest.mark.skipif(O.ref_binary("mf_cpu") is None, reason="oracle/_ref not built")
def test_config1_ml100k_shape_k32_vs_compiled_mf_cpu(tmp_path):
    """configs[0]: MovieLens-100K-shape synthetic, k=32, the reference's mf_sequential CPU path vs
    our GPU path at equal iterations: final TEST RMSE within 0.5 % (mean of two mf_cpu runs; it
    seeds from random_device)."""
    U, I = 943, 1682
    tr, te = cu.synth_ratings(U, I, 100000, integer_ratings=False)
    def write(path, r):
        with open(path, "w") as f:
            f.write("userId,itemId,rating\n")
            f.write("".join("%d,%d,%.1f\n" % (u + 1, i + 1, x) for u, i, x in r))
    write(tmp_path / "train.csv", tr)
    write(tmp_path / "test.csv", te)
    iters, k = 500, 32
    (tmp_path / "c.cfg").write_text("0 %d %d 0.01 42 0.02 0.02 0.02 0.02" % (iters, k))
    finals = []
    for _ in range(2):
        out = O.run_mf_cpu(tmp_path / "c.cfg", tmp_path / "train.csv", tmp_path / "test.csv")
        finals.append(float([l for l in out.splitlines() if l.startswith("TEST:")][-1].split()[-1]))
    ref = float(np.mean(finals))
    r2, rows, cols, gb = cu.readCSV(tmp_path / "train.csv")
    t2, trows, tcols, _ = cu.readCSV(tmp_path / "test.csv")
    rows, cols = max(rows, trows), max(cols, tcols)
    cfg = cu.Config()
    cfg.read_config(tmp_path / "c.cfg")
    out = cu.train(cu.createSparseMatrix(r2, rows, cols), cu.createSparseMatrix(t2, rows, cols), cfg, gb)
    got = out["log"][-1]["test_rmse"]
    assert abs(got - ref) / ref < 0.005, (got, finals)


def test_config4_deterministic_mode_k256_vs_sequential_ordering():
    """configs[3] at a size the CPU replay finishes in seconds: k=256, 64 x 64 blocks, 1.5 M
    ratings of the Netflix-shape generator. Bit-exact against the sequential replay in the
    schedule's order (KERNEL flavour), 5e-5 against the mf_sequential op order."""
    U, I, k, B = 40000, 6000, 256, 64
    tr, _ = cu.synth_ratings(U, I, 1500000, integer_ratings=True, seed=77)
    rng = np.random.RandomState(8)
    P, Q, ub, ib = _model(rng, U, I, k)
    cfg = cu.Config(n_factors=k)
    got = cu.sgd_blocked(tr, P, Q, ub, ib, 3.6, cfg, B, n_passes=1)
    order = O.block_schedule_order(tr, U, I, B)
    want = O.sgd_apply_stream(tr[order], P.ravel(), Q.ravel(), ub, ib, 3.6, O.hyper_from_cfg(cfg), O.FLAVOUR_KERNEL)
    for g, w in zip(got, want):
        assert np.array_equal(g.ravel().view(np.uint32), w.view(np.uint32))
    ref = O.sgd_apply_stream(tr[order], P.ravel(), Q.ravel(), ub, ib, 3.6, O.hyper_from_cfg(cfg), O.FLAVOUR_REF)
    for g, w in zip(got, ref):
        np.testing.assert_allclose(g.ravel(), w, rtol=0, atol=5e-5)


def test_config4_deterministic_mode_full_netflix_shape_is_reproducible():
    """configs[3] at full size (480 189 x 17 770 x ~100 M, k=256): size-independent properties --
    two independent runs of one deterministic pass give identical bits, the pass lowers the
    training loss, and the schedule is a permutation of the ratings."""
    U, I, k = 480189, 17770, 256
    tr, te = cu.synth_ratings(U, I, 100480507, integer_ratings=True)
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    mu = np.float32(tr["rating"].astype(np.float64).sum() / len(tr))
    iters = int(np.ceil(len(tr) / U))  # exactly one pass worth of updates
    runs = []
    for _ in range(2):
        cfg = cu.Config(total_iterations=iters, n_factors=k, check_error=iters, mode=cu.MODE_DETERMINISTIC, n_blocks=512)
        out = cu.train(mtr, mte, cfg, mu)
        runs.append(out)
        assert out["stats"]["updates"] == len(tr)
    a, b = runs
    assert np.array_equal(a["P"].view(np.uint32), b["P"].view(np.uint32))
    assert np.array_equal(a["Q"].view(np.uint32), b["Q"].view(np.uint32))
    assert [r["train_rmse"] for r in a["log"]] == [r["train_rmse"] for r in b["log"]]
    assert np.isfinite(a["log"][-1]["train_rmse"]) and a["log"][-1]["train_rmse"] < a["log"][0]["train_rmse"]


@pytest.mark.gpu
def test_session_from_device_resident_csr_and_invalid_item_ids():
    """cu2b_csr.on_device = 1 (matrix.h:11-19: the reference's CudaCSRMatrix IS device resident): the draw weights that
    feed the stability bound are computed on the device either way, and the session trains like the host-CSR one.
    Item ids outside [0, cols) are rejected before any kernel indexes Q with them."""
    import ctypes as C
    import torch
    tr, te = cu.synth_ratings(3000, 400, 120000, rank=4, noise=0.3, seed=2)
    U, I, k = 3000, 400, 16
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    mu = np.float32(tr["rating"].astype(np.float64).mean())
    init = lambda n: cu.initialize_normal_array(n, k)
    P, Q, ub, ib = init(U * k), init(I * k), init(U), init(I)
    cfg = cu.Config(total_iterations=120, n_factors=k, check_error=40)
    with cu.Session(mtr, mte, cfg, P, Q, ub, ib, mu) as s:
        s.run(120)
        host_log = s.log()

    class DevCsr(cu.CSRMatrix):
        def c(self):
            m = super().c()
            self._keep = [torch.from_numpy(a).cuda() for a in (self.indptr, self.indices, self.data)]
            m.indptr, m.indices, m.data = (C.c_void_p(t.data_ptr()) for t in self._keep)
            m.on_device = 1
            return m
    dtr = DevCsr(U, I, mtr.indptr, mtr.indices, mtr.data)
    dte = DevCsr(U, I, mte.indptr, mte.indices, mte.data)
    with cu.Session(dtr, dte, cfg, P, Q, ub, ib, mu) as s:
        s.run(120)
        dev_log = s.log()
    assert [r["iteration"] for r in dev_log] == [r["iteration"] for r in host_log]
    for a, b in zip(dev_log, host_log):
        assert abs(a["test_rmse"] - b["test_rmse"]) / b["test_rmse"] < 0.005
    bad = cu.CSRMatrix(U, I, mtr.indptr, mtr.indices.copy(), mtr.data)
    bad.indices[1234] = I  # one past the catalogue
    with pytest.raises(cu._lib.Cu2bError) as err:
        cu.Session(bad, mte, cfg, P, Q, ub, ib, mu)
    assert err.value.status == 1 and "item ids" in str(err.value)
    neg = cu.CSRMatrix(U, I, mtr.indptr, mtr.indices.copy(), mtr.data)
    neg.indices[7] = -1  # itemId 0 in a 1-based ratings file
    with pytest.raises(cu._lib.Cu2bError):
        cu.Session(neg, mte, cfg, P, Q, ub, ib, mu)


@pytest.mark.gpu
def test_release_cache_returns_the_pooled_device_memory():
    """Sessions allocate from a memory pool the library owns and keeps warm between calls (the reference
    cudaMallocs per call, matrix.cu:12-40); cu2b_release_cache hands it back to the driver."""
    tr, te = cu.synth_ratings(20000, 2000, 1000000, rank=4, noise=0.3, seed=6)
    U, I, k = 20000, 2000, 128
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    init = lambda n: cu.initialize_normal_array(n, k)
    cfg = cu.Config(total_iterations=2, n_factors=k, check_error=2)
    with cu.Session(mtr, mte, cfg, init(U * k), init(I * k), init(U), init(I), 3.5) as s:
        s.run(2)
    held = cu.device_info()["free_bytes"]
    cu._lib.check(cu._lib.load().cu2b_release_cache())
    assert cu.device_info()["free_bytes"] >= held + (8 << 20)  # P alone is 10 MB


@pytest.mark.gpu
def test_per_rating_sampler_stream_and_training():
    """CU2B_SAMPLER_PER_RATING (no reference counterpart; SURVEY 7 hard part 1): the update stream is a shuffled pass
    over the rating list, bit-identical to the oracle's evaluation of the same keyed permutation; training with it
    applies the same number of updates per iteration and is not worse than the reference's per-user distribution
    (it converges lower, which is why parity is claimed with per_user only)."""
    tr, te = cu.synth_ratings(3000, 500, 100000, rank=4, noise=0.3, seed=9)
    U, I, k = 3000, 500, 32
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    n = mtr.nonzeros
    first, count = n // 3, 2 * n + 1234          # starts inside pass 0, ends inside pass 2
    got = cu.sample_per_rating(mtr, 42, first, count)
    want = O.sample_per_rating(mtr.indptr, mtr.indices, mtr.data, 42, first, count)
    for f in ("user", "item", "rating"):
        assert np.array_equal(got[f], want[f]), f
    one_pass = cu.sample_per_rating(mtr, 42, n, n)  # exactly pass 1: every rating once
    key = lambda a: np.sort(a["user"].astype(np.int64) * I + a["item"])
    assert np.array_equal(key(one_pass), key(np.array(tr, dtype=cu.RATING_DTYPE)))
    mu = np.float32(tr["rating"].astype(np.float64).mean())
    per_user = cu.train(mtr, mte, cu.Config(total_iterations=300, n_factors=k, check_error=100), mu)
    per_rating = cu.train(mtr, mte, cu.Config(total_iterations=300, n_factors=k, check_error=100, sampler=cu.api.SAMPLER_PER_RATING), mu)
    assert per_rating["stats"]["updates"] == per_user["stats"]["updates"] == 300 * U
    a, b = per_rating["log"][-1]["test_rmse"], per_user["log"][-1]["test_rmse"]
    assert np.isfinite(a) and a <= b * 1.005, (a, b)
    part = cu.dsgd_partition(tr, U, I, 1)
    init = lambda m: cu.initialize_normal_array(m, k)
    inp = cu.dsgd_rank_inputs(tr, te, U, I, part, 0, init(U * k), init(I * k), init(U), init(I))
    with pytest.raises(cu._lib.Cu2bError):
        cu.Dsgd(0, 1, inp, part, cu.Config(total_iterations=10, n_factors=k, sampler=cu.api.SAMPLER_PER_RATING), mu)


@pytest.mark.gpu
def test_run_download_equals_run_then_download():
    """cu2b_session_run_download: the model leaves on a second stream while the run's last loss check reads it. On a
    problem whose items are never shared between users every schedule is deterministic, so the outputs must equal a
    run followed by a download bit for bit (item placement and padded biases included)."""
    rng = np.random.RandomState(12)
    U, k, iters = 3000, 128, 64
    deg = rng.randint(1, 6, U)
    n = int(deg.sum())
    tr = np.zeros(n, dtype=cu.RATING_DTYPE)
    tr["user"], tr["item"], tr["rating"] = np.repeat(np.arange(U), deg), rng.permutation(n), rng.randint(1, 6, n)
    tr = tr[np.lexsort((tr["item"], tr["user"]))]
    te = tr[::3].copy()
    mtr, mte = cu.createSparseMatrix(tr, U, n), cu.createSparseMatrix(te, U, n)
    init = lambda m: cu.initialize_normal_array(m, k)
    P, Q, ub, ib = init(U * k), init(n * k), init(U), init(n)
    cfg = cu.Config(total_iterations=iters, n_factors=k, check_error=32)
    with cu.Session(mtr, mte, cfg, P, Q, ub, ib, 3.0) as s:
        s.run(iters)
        want, want_log = s.download(), s.log()
    with cu.Session(mtr, mte, cfg, P, Q, ub, ib, 3.0) as s:
        got, got_log = s.run_download(iters), s.log()
    assert got_log == want_log and [r["iteration"] for r in got_log] == [1, 32, 64]
    for a, b in zip(got, want):
        assert np.array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))
    # a call that does not end on a check, and partial outputs
    with cu.Session(mtr, mte, cfg, P, Q, ub, ib, 3.0) as s:
        s.run(10)
        Pp, _, ubp, _ = s.download()
    with cu.Session(mtr, mte, cfg, P, Q, ub, ib, 3.0) as s:
        got10 = s.run_download(10)
    assert np.array_equal(got10[0].view(np.uint32), Pp.view(np.uint32)) and np.array_equal(got10[2].view(np.uint32), ubp.view(np.uint32))


@pytest.mark.gpu
def test_non_finite_model_is_reported_as_diverged_not_as_success():
    """The device-side guard (loss_kernels.cuh flag_non_finite) without any chaos involved: a model that is not finite
    when a loss check evaluates it makes cu2b_session_run / cu2b_train return CU2B_ERR_DIVERGED (the reference prints
    `nan` and exits 0, training.cu:118-160; round 1 did the same at 4 GPUs)."""
    tr, te = cu.synth_ratings(2000, 300, 60000, rank=4, noise=0.3, seed=4)
    U, I, k = 2000, 300, 16
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    mu = np.float32(tr["rating"].astype(np.float64).mean())
    init = lambda n: cu.initialize_normal_array(n, k)
    P, Q, ub, ib = init(U * k), init(I * k), init(U), init(I)
    cfg = cu.Config(total_iterations=40, n_factors=k, check_error=20)
    with cu.Session(mtr, mte, cfg, P, Q, ub, ib, mu) as s:  # the finite model trains
        s.run(40)
        assert all(np.isfinite(r["test_rmse"]) for r in s.log())
    for poison in (np.float32("nan"), np.float32("inf")):
        Qbad = Q.copy()
        Qbad[int(mtr.indices[0]) * k] = poison  # a rated item: the first train check sees it
        with cu.Session(mtr, mte, cfg, P, Qbad, ub, ib, mu) as s:
            with pytest.raises(cu._lib.Cu2bError) as err:
                s.run(40)
            assert err.value.status == 6 and "non-finite" in str(err.value)
    # a learning rate that makes plain SGD itself blow up: same status through the train() entry point
    wild = cu.Config(total_iterations=200, n_factors=k, check_error=50, learning_rate=50.0)
    with pytest.raises(cu._lib.Cu2bError) as err:
        cu.train(mtr, mte, wild, mu)
    assert err.value.status == 6
