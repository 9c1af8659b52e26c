"""Convergence vs DSGD round length: logical ranks on ONE GPU against the single-GPU session.
usage: python tools/dsgd_round_sweep.py <world> <iters> <check> <round> [<round> ...]"""
import json
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cu2rec_b200 as cu  # noqa: E402

world, iters, ce = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rounds = [int(x) for x in sys.argv[4:]]
wl = os.environ.get("SWEEP_WORKLOAD", "ml20m")
k = int(os.environ.get("SWEEP_K", "64"))
tr, te, U, I = bench.make_workload(wl)
mu = np.float32(tr["rating"].astype(np.float64).sum() / len(tr))
init = lambda n: cu.initialize_normal_array(n, k)
P, Q, ub, ib = init(U * k), init(I * k), init(U), init(I)
mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
cfg = cu.Config(total_iterations=iters, n_factors=k, check_error=ce)
with cu.Session(mtr, mte, cfg, P, Q, ub, ib, mu) as s:
    s.run(iters)
    ref = s.log()
    st = s.stats()
print(json.dumps({"mode": "single", "test_rmse": [round(r["test_rmse"], 5) for r in ref],
                  "Gups": st["updates"] / st["total_ms"] / 1e6}), flush=True)
part = cu.dsgd_partition(tr, U, I, world)
inputs = [cu.dsgd_rank_inputs(tr, te, U, I, part, r, P, Q, ub, ib) for r in range(world)]
for T in rounds:
    os.environ["CU2B_DSGD_ROUND"] = str(T)
    ranks = [cu.Dsgd(r, world, inputs[r], part, cu.Config(total_iterations=iters, n_factors=k, check_error=ce), mu)
             for r in range(world)]
    hs = [d.handle for d in ranks]
    for d in ranks:
        d.connect(hs)
    th = [threading.Thread(target=d.run, args=(iters,)) for d in ranks]
    [t.start() for t in th]
    [t.join() for t in th]
    lg = ranks[0].log()
    tot = sum(d.stats()["updates"] for d in ranks)
    ms = max(d.stats()["total_ms"] for d in ranks)
    print(json.dumps({"mode": "dsgd%d" % world, "round": T, "test_rmse": [round(r["test_rmse"], 5) for r in lg],
                      "final_vs_single_pct": 100 * (lg[-1]["test_rmse"] / ref[-1]["test_rmse"] - 1),
                      "Gups_shared_gpu": tot / ms / 1e6}), flush=True)
    for d in ranks:
        d.close()
