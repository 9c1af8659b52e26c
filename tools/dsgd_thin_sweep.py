"""Item-step thinning (CU2B_DSGD_THIN) vs the in-flight cap on the single-block stand-in workload
(`nfblock8`: what one rank sees inside one sub-epoch at 8 GPUs): throughput of mf_sgd_user_runs at forced
grid sizes, with and without thinning, and the test RMSE after the same number of iterations.
usage: python tools/dsgd_thin_sweep.py <grid> [<grid> ...]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, cu2rec_b200 as cu
k = int(os.environ.get("SWEEP_K", "128"))
iters = int(os.environ.get("SWEEP_ITERS", "256"))
tr, te, U, I = bench.make_workload(os.environ.get("SWEEP_WORKLOAD", "nfblock8"))
mu = np.float32(tr["rating"].astype(np.float64).sum() / len(tr))
init = lambda n: cu.initialize_normal_array(n, k)
P, Q, ub, ib = init(U * k), init(I * k), init(U), init(I)
part = cu.dsgd_partition(tr, U, I, 1)
inp = cu.dsgd_rank_inputs(tr, te, U, I, part, 0, P, Q, ub, ib)
os.environ["CU2B_DSGD_ROUND"] = os.environ.get("SWEEP_ROUND", "64")
for g in sys.argv[1:]:
    budget = os.environ.get("SWEEP_THIN", "0.5")
    for thin in os.environ.get("SWEEP_VARIANTS", ",bias,rows+bias").split(","):  # default kernel (grid forced, no cap), bias steps only, row and bias steps
        os.environ["CU2B_DSGD_GRID"] = g
        os.environ.pop("CU2B_DSGD_THIN", None)
        os.environ.pop("CU2B_DSGD_THIN_BIAS", None)
        if thin == "bias":
            os.environ["CU2B_DSGD_THIN_BIAS"] = budget
        elif thin:
            os.environ["CU2B_DSGD_THIN"] = budget
        d = cu.Dsgd(0, 1, inp, part, cu.Config(total_iterations=2 * iters, n_factors=k, check_error=iters), mu)
        d.connect([d.handle])
        diverged = None
        try:
            d.run(iters); d.stats(reset=True); d.run(iters)
        except cu._lib.Cu2bError as exc:  # CU2B_ERR_DIVERGED: the device-side non-finite guard
            diverged = str(exc)[:160]
        st = d.stats()
        if diverged or not st["sgd_ms"]:
            print(json.dumps({"grid": int(g), "thinned": thin or None, "budget": budget if thin else None, "diverged": diverged,
                              "workload": os.environ.get("SWEEP_WORKLOAD", "nfblock8"), "ib_stride": os.environ.get("CU2B_IB_STRIDE"),
                              "test_rmse": [(r["iteration"], round(r["test_rmse"], 4)) for r in d.log()]}), flush=True)
            d.close()
            continue
        ups = st["updates"] / (st["sgd_ms"] / 1e3)
        print(json.dumps({"grid": int(g), "thinned": thin or None, "budget": budget if thin else None, "workload": os.environ.get("SWEEP_WORKLOAD", "nfblock8"),
                          "ib_stride": os.environ.get("CU2B_IB_STRIDE"), "sgd_Gups": round(ups / 1e9, 3), "sgd_ms": round(st["sgd_ms"], 2),
                          "sampler_ms": round(st["sampler_ms"], 2),
                          "test_rmse": [(r["iteration"], round(r["test_rmse"], 4)) for r in d.log()]}), flush=True)
        d.close()
