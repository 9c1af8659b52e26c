import os, sys, threading
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import bench, cu2rec_b200 as cu
world, iters, ce, T, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
os.environ["CU2B_DSGD_ROUND"] = T
tr, te, U, I = bench.make_workload("ml20m")
part = cu.dsgd_partition(tr, U, I, world)
init = lambda n: cu.initialize_normal_array(n, k)
P, Q, ub, ib = init(U*k), init(I*k), init(U), init(I)
mu = np.float32(tr["rating"].astype(np.float64).mean())
ranks = [cu.Dsgd(r, world, cu.dsgd_rank_inputs(tr, te, U, I, part, r, P, Q, ub, ib), part, cu.Config(total_iterations=iters, n_factors=k, check_error=ce), mu) for r in range(world)]
hs = [d.handle for d in ranks]
[d.connect(hs) for d in ranks]
th = [threading.Thread(target=d.run, args=(iters,)) for d in ranks]
[t.start() for t in th]; [t.join() for t in th]
print("log", [(r["iteration"], round(r["train_rmse"], 4), round(r["test_rmse"], 4)) for r in ranks[0].log()])
for r, d in enumerate(ranks):
    Ps, Qn, ubs, ibn = d.download()
    bad_u = np.flatnonzero(~np.isfinite(Ps).all(1))
    print("rank", r, "P nonfinite rows", len(bad_u), "of", len(Ps), "max|P|", np.nanmax(np.abs(Ps)), "ub nonfinite", int((~np.isfinite(ubs)).sum()))
    for b in range(world):
        reg = Qn[part.item_block_ptr[b]:part.item_block_ptr[b+1]]
        print("   Q block", b, "nonfinite rows", int((~np.isfinite(reg).all(1)).sum()), "max|Q|", float(np.nanmax(np.abs(reg))), "ib max", float(np.nanmax(np.abs(ibn[part.item_block_ptr[b]:part.item_block_ptr[b+1]]))))
    if len(bad_u):
        users = np.flatnonzero(part.user_block == r)
        deg = np.bincount(tr["user"], minlength=U)
        print("   first bad users (orig id, degree):", [(int(users[x]), int(deg[users[x]])) for x in bad_u[:8]])
