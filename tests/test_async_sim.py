"""tools/async_sim (the CPU model behind DESIGN 6.1) is pinned to the oracle: with one update in flight and
rounds of one iteration it IS sequential SGD in the oracle trainer's order, bit for bit; with updates in
flight it is not, and far beyond the stability bound it diverges."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cu2rec_b200 as cu
import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    so = tmp_path_factory.mktemp("sim") / "libasync_sim.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Werror",
                    os.path.join(ROOT, "tools", "async_sim", "async_sim.cpp"), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    lib.async_sim_train.restype = C.c_int
    return lib


def _run(lib, mtr, mte, model, mu, k, iters, check, G, user_block, item_block, round_iters, inflight, stale, scale=None, mode=0):
    P, Q, ub, ib = (np.array(x, copy=True) for x in model)
    log = np.zeros(3 * (iters // check + 4), np.float64)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    n = lib.async_sim_train(mtr.rows, mtr.cols, p(mtr.indptr), p(mtr.indices), p(mtr.data), p(mte.indptr), p(mte.indices),
                            p(mte.data), p(P), p(Q), p(ub), p(ib), C.c_float(mu), k, C.c_float(0.01), C.c_float(0.02), 42,
                            iters, check, G, p(user_block), p(item_block), round_iters, inflight, C.c_float(stale), p(scale),
                            mode, p(log), len(log) // 3)
    return (P, Q, ub, ib), log[: 3 * n].reshape(-1, 3)


def _problem(U=400, I=120, n=12000, k=8):
    tr, te = cu.synth_ratings(U, I, n, rank=4, noise=0.3, seed=3)
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    mu = np.float32(tr["rating"].astype(np.float64).mean())
    init = lambda m: O.init_normal(m, k)
    return tr, mtr, mte, mu, (init(U * k), init(I * k), init(U), init(I))


def test_one_update_in_flight_is_the_oracle_trainer_bit_for_bit(sim):
    k, iters = 8, 30
    tr, mtr, mte, mu, model = _problem(k=k)
    got, log = _run(sim, mtr, mte, model, mu, k, iters, 10, 1, None, None, 1, 1, 1.0)
    P, Q, ub, ib, olog = O.train((mtr.indptr, mtr.indices, mtr.data), (mte.indptr, mte.indices, mte.data), *model, mu,
                                 O.hyper(k), 42, iters, check_error=10, use_decay=False)
    for a, b in zip(got, (P, Q, ub, ib)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert [int(x) for x in log[:, 0]] == [10, 20, 30]
    assert abs(log[-1, 1] - olog[-1]["test_rmse"]) < 1e-5 and log[:, 2].max() == 0  # never an unseen step


def test_updates_in_flight_change_the_result_and_too_many_diverge(sim):
    k, iters = 8, 60
    tr, mtr, mte, mu, model = _problem(U=2000, I=40, n=30000, k=k)  # 40 items: the hottest takes several % of the draws
    seq, _ = _run(sim, mtr, mte, model, mu, k, iters, 20, 1, None, None, 20, 1, 1.0)
    few, log_few = _run(sim, mtr, mte, model, mu, k, iters, 20, 1, None, None, 20, 64, 1.0)
    assert not np.array_equal(seq[1], few[1]) and log_few[:, 2].max() >= 1
    assert np.isfinite(log_few[-1, 1]) and log_few[-1, 1] < 1.2
    many, log_many = _run(sim, mtr, mte, model, mu, k, 400, 100, 1, None, None, 20, 2000, 8.0)
    assert not np.isfinite(log_many[-1, 1]) or log_many[-1, 1] > 5  # lr x unseen steps far beyond 2
    # the bias-only remedy (mode 2: scale the bias step of the overloaded items) restores stability
    deg = np.diff(mtr.indptr)
    w = np.bincount(mtr.indices, weights=np.repeat(1.0 / np.maximum(deg, 1), deg), minlength=mtr.cols)
    scale = np.minimum(1.0, 0.5 / (0.01 * (w / w.sum()) * 2000 * 8.0)).astype(np.float32)
    fixed, log_fixed = _run(sim, mtr, mte, model, mu, k, 400, 100, 1, None, None, 20, 2000, 8.0, scale=scale, mode=2)
    assert np.isfinite(log_fixed[-1, 1]) and log_fixed[-1, 1] < 1.2
