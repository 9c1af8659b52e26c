import os, sys, time, threading
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import cu2rec_b200 as cu
world, iters, ce, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
os.environ["CU2B_DSGD_ROUND"] = T
tr, te = cu.synth_ratings(3000, 400, 120000, rank=4, noise=0.3, seed=21)
U, I, k = 3000, 400, 16
part = cu.dsgd_partition(tr, U, I, world)
init = lambda n: cu.initialize_normal_array(n, k)
P, Q, ub, ib = init(U*k), init(I*k), init(U), init(I)
mu = np.float32(tr["rating"].astype(np.float64).mean())
ranks = [cu.Dsgd(r, world, cu.dsgd_rank_inputs(tr, te, U, I, part, r, P, Q, ub, ib), part, cu.Config(total_iterations=iters, n_factors=k, check_error=ce), mu) for r in range(world)]
hs = [d.handle for d in ranks]
[d.connect(hs) for d in ranks]
t0 = time.time()
errs = []
def work(d):
    try: d.run(iters)
    except Exception as e: errs.append(str(e))
th = [threading.Thread(target=work, args=(d,)) for d in ranks]
[t.start() for t in th]; [t.join() for t in th]
print("world", world, "round", T, "time %.2fs" % (time.time() - t0), "errs", errs, "log", [(r["iteration"], round(r["test_rmse"], 4)) for r in ranks[0].log()][-2:], flush=True)
