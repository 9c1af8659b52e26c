"""Stability of asynchronous SGD under a bounded number of in-flight updates (CPU model, see
async_sim.cpp): reproduces the limit DESIGN.md 6.1 measured on 8 x B200 and evaluates the
staleness-aware item step scale proposed for round 2. Writes one JSON line per configuration.

    g++ -O3 -march=native -std=c++17 -fPIC -shared tools/async_sim/async_sim.cpp -o tools/async_sim/libasync_sim.so
    python tools/async_sim/run.py --out profiles/r1_async_stability_sim.jsonl
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def problem(U, I, n, k):
    import cu2rec_b200 as cu
    tr, te = cu.synth_ratings(U, I, n, rank=16, noise=0.5, integer_ratings=True)
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    mu = np.float32(tr["rating"].astype(np.float64).mean())
    init = lambda m: cu.initialize_normal_array(m, k)
    return tr, mtr, mte, mu, (init(U * k), init(I * k), init(U), init(I))


def draw_weights(mtr):
    """w[i] = expected draws of item i per iteration (one uniform draw per user per iteration)."""
    deg = np.diff(mtr.indptr)
    per_rating = np.repeat(1.0 / np.maximum(deg, 1), deg)
    return np.bincount(mtr.indices, weights=per_rating, minlength=mtr.cols)


def run_one(job):
    name, G, inflight, stale, scale_budget, args = job
    mode = {"scale": 0, "thinned": 1, "bias scaled": 2, "bias thinned": 3, "rows thinned": 4}.get(
        next((t for t in ("bias scaled", "bias thinned", "rows thinned", "thinned", "scale") if t in name), "scale"), 0)
    import cu2rec_b200 as cu
    lib = C.CDLL(os.path.join(ROOT, "tools", "async_sim", "libasync_sim.so"))
    tr, mtr, mte, mu, (P, Q, ub, ib) = problem(args["U"], args["I"], args["n"], args["k"])
    U, I, k, lr = args["U"], args["I"], args["k"], args["lr"]
    part = cu.dsgd_partition(tr, U, I, G)
    inv = np.empty(I, np.int64)
    inv[part.item_new] = np.arange(I)
    item_block = (np.searchsorted(part.item_block_ptr, part.item_new, side="right") - 1).astype(np.int32)
    user_block = np.ascontiguousarray(part.user_block, dtype=np.int32)
    w = draw_weights(mtr)
    block_tot = np.bincount(item_block, weights=w, minlength=G)
    share = w / block_tot[item_block]  # share of the draws of its own block
    hot = float(share.max())
    if inflight <= 0:  # given as a budget: lr * hot_share * inflight = -inflight
        inflight = max(1, int(round(-inflight / (lr * hot))))
    scale = None
    if scale_budget > 0:  # staleness-aware item step: keep lr * (unseen steps on the item) <= budget
        scale = np.minimum(1.0, scale_budget / (lr * share * inflight * stale + 1e-30)).astype(np.float32)
    cap = args["iters"] // args["check"] + 4
    log = np.zeros(3 * cap, np.float64)
    t0 = time.time()
    lib.async_sim_train.restype = C.c_int
    n = lib.async_sim_train(U, I, _p(mtr.indptr), _p(mtr.indices), _p(mtr.data), _p(mte.indptr), _p(mte.indices), _p(mte.data),
                            _p(P), _p(Q), _p(ub), _p(ib), C.c_float(mu), k, C.c_float(lr), C.c_float(args["reg"]), 42,
                            args["iters"], args["check"], G, _p(user_block), _p(item_block), args["round"], inflight, C.c_float(stale),
                            _p(scale) if scale is not None else None, mode, _p(log), cap)
    rows = log[: 3 * min(n, cap)].reshape(-1, 3)
    rm = [None if not np.isfinite(r) else round(float(r), 5) for r in rows[:, 1]]
    return {"config": name, "ranks": G, "inflight_per_rank": inflight, "hot_share_in_block": round(hot, 5),
            "stale_factor": stale,
            "lr_x_share_x_inflight": round(lr * hot * inflight, 3),
            "scaled_items": int((scale < 1).sum()) if scale is not None else 0,
            "min_scale": round(float(scale.min()), 3) if scale is not None else 1.0,
            "iterations": [int(x) for x in rows[:, 0]], "test_rmse": rm,
            "max_unseen_steps_at_a_read": [int(x) for x in rows[:, 2]],
            "diverged": bool(len(rm) and (rm[-1] is None or rm[-1] > 5)), "seconds": round(time.time() - t0, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r1_async_stability_sim.jsonl"))
    ap.add_argument("--users", type=int, default=60000)
    ap.add_argument("--items", type=int, default=17770)
    ap.add_argument("--ratings", type=int, default=12_500_000)
    ap.add_argument("-k", type=int, default=32)
    ap.add_argument("--iters", type=int, default=2048)
    ap.add_argument("--check", type=int, default=512)
    ap.add_argument("--round", type=int, default=64)
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    a = ap.parse_args()
    args = {"U": a.users, "I": a.items, "n": a.ratings, "k": a.k, "lr": 0.01, "reg": 0.02, "iters": a.iters,
            "check": a.check, "round": a.round}
    full = 148 * 5 * 8  # lane groups of the default kernel at k = 128: 148 SMs x 5 CTAs x 8 warps
    jobs = [("sequential (1 in flight)", 1, 1, 1.0, 0.0, args)]
    for stale in (1.0, 2.0):
        jobs += [
            ("1 GPU, full occupancy", 1, full, stale, 0.0, args),
            ("8 ranks, budget 0.5 (product cap)", 8, -0.5, stale, 0.0, args),
            ("8 ranks, budget 1.0 (diverged on B200)", 8, -1.0, stale, 0.0, args),
            ("8 ranks, budget 2.0", 8, -2.0, stale, 0.0, args),
            ("8 ranks, full occupancy, no cap", 8, full, stale, 0.0, args),
            ("8 ranks, full occupancy, item step scale (budget 0.5)", 8, full, stale, 0.5, args),
            ("8 ranks, full occupancy, item steps thinned (budget 0.5)", 8, full, stale, 0.5, args),
            ("8 ranks, full occupancy, only the item bias scaled (budget 0.5)", 8, full, stale, 0.5, args),
            ("8 ranks, full occupancy, only the item bias thinned (budget 0.5)", 8, full, stale, 0.5, args),
            ("8 ranks, full occupancy, only the item rows thinned (budget 0.5)", 8, full, stale, 0.5, args),
        ]
    with mp.Pool(min(a.procs, len(jobs))) as pool, open(a.out, "w") as f:
        for res in pool.imap(run_one, jobs):
            f.write(json.dumps(res) + "\n")
            f.flush()
            print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
