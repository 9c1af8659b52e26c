"""Wall-clock phases of the end-to-end path bench.py times (session create from pinned host
buffers, run, download into pinned buffers, destroy). CU2B_TRACE=1 on the last repetition adds
the library's own phase marks (they synchronise the device, so that repetition is slower)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, cu2rec_b200 as cu
tr, te, U, I = bench.make_workload("netflix")
k = 128
mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
mu = np.float32(3.6)
init = lambda n: cu.initialize_normal_array(n, k)
hp = [bench.pin(getattr(mtr, n)) for n in ("indptr", "indices", "data")]
hq = [bench.pin(getattr(mte, n)) for n in ("indptr", "indices", "data")]
ptr, pte = cu.CSRMatrix(U, I, *hp), cu.CSRMatrix(U, I, *hq)
P, Q, ub, ib = (bench.pin(x) for x in (init(U * k), init(I * k), init(U), init(I)))
out = tuple(bench.pin(x) for x in (P, Q, ub, ib))
cfg = cu.Config(total_iterations=500, n_factors=k, check_error=500)
for rep in range(8):
    if rep in (1, 7): os.environ["CU2B_TRACE"] = "1"
    else: os.environ.pop("CU2B_TRACE", None)
    t0 = time.perf_counter()
    s = cu.Session(ptr, pte, cfg, P, Q, ub, ib, mu); t1 = time.perf_counter()
    if os.environ.get("TRACE_FUSED_DOWNLOAD", "1") == "1":
        s.run_download(500, out=out); t2 = t3 = time.perf_counter()   # D2H overlaps the last loss check
    else:
        s.run(500); t2 = time.perf_counter()
        s.download(out=out); t3 = time.perf_counter()
    lg = s.log(); t4 = time.perf_counter()
    s.close(); t5 = time.perf_counter()
    print("rep %d create %.1f run %.1f download %.1f log %.1f destroy %.1f total %.1f ms" % (
        rep, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t5-t4), 1e3*(t5-t0)), flush=True)
