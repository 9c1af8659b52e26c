"""A/B sweep of Hogwild kernel settings on a bench workload (GPU box).
usage: [SWEEP_K=128 SWEEP_WORKLOAD=netflix] python tools/sweep_sgd.py "label;ENV=VAL;ENV=VAL" ...
e.g. "w0;CU2B_WMODE=0" "w3;CU2B_WMODE=3" "chunk128;CU2B_TUNE_CHUNK=128"
(round-1 history: profiles/r1_sweep*.jsonl were taken with a build that also exposed unroll /
occupancy / weak-memop template variants through CU2B_SGD_TUNE="wmode,unr,occ,memop")"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cu2rec_b200 as cu  # noqa: E402

variants = sys.argv[1:] or ["1,2,1"]
iters = int(os.environ.get("SWEEP_ITERS", "200"))
tr, te, U, I = bench.make_workload(os.environ.get("SWEEP_WORKLOAD", "netflix"))
k = int(os.environ.get("SWEEP_K", "128"))
mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
mu = np.float32(tr["rating"].astype(np.float64).sum() / len(tr))
init = lambda n: cu.initialize_normal_array(n, k)
P0, Q0, ub0, ib0 = init(U * k), init(I * k), init(U), init(I)
for v in variants:
    # "w,u,o,m[;ENV=VAL;ENV=VAL]"
    parts = v.split(";")
    for kv in parts[1:]:
        key, val = kv.split("=")
        os.environ[key] = val
    cfg = cu.Config(total_iterations=10 ** 6, n_factors=k, check_error=10 ** 6,
                    is_train=int(os.environ.get("SWEEP_IS_TRAIN", "1")))
    with cu.Session(mtr, mte, cfg, P0, Q0, ub0, ib0, mu) as s:
        s.run(iters)
        s.stats(reset=True)
        s.run(iters)
        st = s.stats()
        ev = s.eval()
    print(json.dumps({"variant": v, "sgd_Gups": st["updates"] / st["sgd_ms"] / 1e6, "sgd_ms": st["sgd_ms"],
                      "total_Gups": st["updates"] / st["total_ms"] / 1e6, "sampler_ms": st["sampler_ms"],
                      "test_rmse": ev["test_rmse"]}), flush=True)
    for kv in parts[1:]:
        os.environ.pop(kv.split("=")[0], None)
