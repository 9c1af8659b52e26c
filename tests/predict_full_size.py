"""BASELINE config 5: batched predict of the whole Netflix-shape catalogue (480 189 x 17 770, k=128),
top-10 unrated items per user, on the tensor cores; a random sample of users is checked against
the CPU brute force (the oracle: this script lives under tests/ because only tests may use it).
Prints one JSON line (profiles/r1_predict.json).   python tests/predict_full_size.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench  # noqa: E402
import cu2rec_b200 as cu  # noqa: E402
import oracle as O  # noqa: E402

k, topk = int(os.environ.get("PREDICT_K", "128")), int(os.environ.get("PREDICT_TOPK", "10"))
tr, te, U, I = bench.make_workload("netflix")
mtr = cu.createSparseMatrix(tr, U, I)
mu = np.float32(tr["rating"].astype(np.float64).sum() / len(tr))
# a trained-looking model: a few hundred iterations so that scores are not pure noise
init = lambda n: cu.initialize_normal_array(n, k)
cfg = cu.Config(total_iterations=300, n_factors=k, check_error=300)
with cu.Session(mtr, cu.createSparseMatrix(te, U, I), cfg, init(U * k), init(I * k), init(U), init(I), mu) as s:
    s.run(300)
    P, Q, ub, ib = s.download()
t0 = time.perf_counter()
items, scores, ms = cu.predict_topk(P, Q, ub, ib, mu, topk, exclude=mtr)
wall = time.perf_counter() - t0
items, scores, ms = cu.predict_topk(P, Q, ub, ib, mu, topk, exclude=mtr)  # warm
rng = np.random.RandomState(0)
sample = np.sort(rng.choice(U, 256, replace=False))
sub_ptr = np.concatenate([[0], np.cumsum([mtr.indptr[u + 1] - mtr.indptr[u] for u in sample])]).astype(np.int32)
sub_idx = np.concatenate([mtr.indices[mtr.indptr[u]:mtr.indptr[u + 1]] for u in sample]).astype(np.int32)
wi, ws = O.predict_topk(P[sample], Q, ub[sample], ib, mu, topk, exclude=(sub_ptr, sub_idx))
exact_items = bool(np.array_equal(items[sample], wi))
exact_scores = bool(np.array_equal(scores[sample].view(np.uint32), ws.view(np.uint32)))
kp = ((k + 63) // 64 * 64) if k > 128 else ((k + 31) // 32 * 32)
flops = 2.0 * U * I * kp  # the tensor cores see the zero-padded rows
print(json.dumps({"config": "batched predict %d users x %d items, k=%d, top-%d, rated items excluded" % (U, I, k, topk),
                  "candidates_ms": ms["candidates_ms"], "rescore_ms": ms["rescore_ms"],
                  "k_padded": kp, "tf32_tflops": flops / (ms["candidates_ms"] * 1e-3) / 1e12,
                  "users_per_s": U / ((ms["candidates_ms"] + ms["rescore_ms"]) * 1e-3),
                  "e2e_wall_s_first_call_incl_h2d_bitmap_d2h": wall,
                  "sample_users_checked": int(len(sample)), "items_exact": exact_items, "scores_bit_exact": exact_scores}))
