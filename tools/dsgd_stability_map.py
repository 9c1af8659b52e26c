"""Stability / throughput map of the DSGD sub-epoch kernel on the single-block stand-in workloads
(nfblock8 / nfblock4 / nfblock2 = what one rank sees inside a sub-epoch at 8 / 4 / 2 GPUs): for every
(grid, thinning mode, budget) does the run stay finite, how fast is the kernel, and what test RMSE does it
reach. One JSON line per run.

usage: python tools/dsgd_stability_map.py <workload> <grid,grid,...> <mode:budget,mode:budget,...>
       modes: none | bias | rows (= row and bias steps)      e.g.  nfblock8 296,592,1184 none:0,bias:0.25,rows:0.25
env:   SWEEP_ITERS (iterations per check, default 256), SWEEP_CHECKS (default 2), SWEEP_ROUND (default 64)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cu2rec_b200 as cu  # noqa: E402

wl = sys.argv[1]
grids = [int(x) for x in sys.argv[2].split(",")]
modes = [m.split(":") for m in sys.argv[3].split(",")]
k = int(os.environ.get("SWEEP_K", "128"))
iters = int(os.environ.get("SWEEP_ITERS", "256"))
checks = int(os.environ.get("SWEEP_CHECKS", "2"))
os.environ["CU2B_DSGD_ROUND"] = os.environ.get("SWEEP_ROUND", "64")
tr, te, U, I = bench.make_workload(wl)
mu = np.float32(tr["rating"].astype(np.float64).sum() / len(tr))
init = lambda n: cu.initialize_normal_array(n, k)
P, Q, ub, ib = init(U * k), init(I * k), init(U), init(I)
part = cu.dsgd_partition(tr, U, I, 1)
inp = cu.dsgd_rank_inputs(tr, te, U, I, part, 0, P, Q, ub, ib)
for g in grids:
    for mode, budget in modes:
        for name in ("CU2B_DSGD_THIN", "CU2B_DSGD_THIN_BIAS"):
            os.environ.pop(name, None)
        os.environ["CU2B_DSGD_GRID"] = str(g)
        if mode == "bias":
            os.environ["CU2B_DSGD_THIN_BIAS"] = budget
        elif mode == "rows":
            os.environ["CU2B_DSGD_THIN"] = budget
        d = cu.Dsgd(0, 1, inp, part, cu.Config(total_iterations=checks * iters, n_factors=k, check_error=iters), mu)
        d.connect([d.handle])
        diverged = None
        try:
            d.run(iters)
            d.stats(reset=True)
            for _ in range(checks - 1):
                d.run(iters)
        except cu._lib.Cu2bError as exc:
            diverged = "status %d" % exc.status
        st = d.stats()
        row = {"workload": wl, "grid": g, "mode": mode, "budget": float(budget), "diverged": diverged,
               "ib_stride": os.environ.get("CU2B_IB_STRIDE"), "placement": os.environ.get("CU2B_PLACEMENT"),
               "sgd_Gups": round(st["updates"] / st["sgd_ms"] / 1e6, 3) if st["sgd_ms"] and not diverged else None,
               "test_rmse": [(r["iteration"], round(r["test_rmse"], 4)) for r in d.log()]}
        print(json.dumps(row), flush=True)
        d.close()
