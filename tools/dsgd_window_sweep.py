"""Read -> atomic-add window of the DSGD sub-epoch kernel: one rank (world = 1, all items in one
block) on one GPU, grid size forced through CU2B_DSGD_GRID. window = in-flight lane groups /
throughput. usage: python tools/dsgd_window_sweep.py <grid> [<grid> ...]   (SWEEP_ROUND = round length)"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, cu2rec_b200 as cu
k = int(os.environ.get("SWEEP_K", "128"))
iters = int(os.environ.get("SWEEP_ITERS", "128"))
tr, te, U, I = bench.make_workload(os.environ.get("SWEEP_WORKLOAD", "netflix"))
mu = np.float32(tr["rating"].astype(np.float64).sum() / len(tr))
init = lambda n: cu.initialize_normal_array(n, k)
P, Q, ub, ib = init(U * k), init(I * k), init(U), init(I)
part = cu.dsgd_partition(tr, U, I, 1)
inp = cu.dsgd_rank_inputs(tr, te, U, I, part, 0, P, Q, ub, ib)
os.environ["CU2B_DSGD_ROUND"] = os.environ.get("SWEEP_ROUND", "8")
for g in sys.argv[1:]:
    os.environ["CU2B_DSGD_GRID"] = g
    d = cu.Dsgd(0, 1, inp, part, cu.Config(total_iterations=10 ** 6, n_factors=k, check_error=10 ** 6), mu)
    d.connect([d.handle])
    d.run(iters); d.stats(reset=True); d.run(iters)
    st = d.stats()
    groups = int(g) * 8 * (32 // min(32, max(1, 1 << int(np.ceil(np.log2(max(1, (k + 3) // 4)))))))
    ups = st["updates"] / (st["sgd_ms"] / 1e3)
    print(json.dumps({"grid": int(g), "groups_in_flight": groups, "sgd_Gups": ups / 1e9, "window_us": groups / ups * 1e6,
                      "sampler_ms": st["sampler_ms"], "sgd_ms": st["sgd_ms"]}), flush=True)
    d.close()
