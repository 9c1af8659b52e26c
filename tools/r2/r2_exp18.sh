#!/bin/bash
# Two B200: cu2b_train / bin/mf with config token 17 = 2 on two real devices (one host thread + context per rank inside
# one process, fused sub-epoch kernels), against the single-GPU run of the same command line.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out /tmp/mf2
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, cu2rec_b200 as cu
tr, te = cu.synth_ratings(138493, 26744, 20000263, integer_ratings=False, seed=20240607)
cu.write_ratings_csv("/tmp/mf2/train.csv", tr)
cu.write_ratings_csv("/tmp/mf2/test.csv", te)
open("/tmp/mf2/one.cfg", "w").write("0 600 64 0.01 42 0.02 0.02 0.02 0.02 32 2 0.2 200 0 0 0 1")
open("/tmp/mf2/two.cfg", "w").write("0 600 64 0.01 42 0.02 0.02 0.02 0.02 32 2 0.2 200 0 0 0 2")
PY
( time bin/mf -c /tmp/mf2/one.cfg /tmp/mf2/train.csv /tmp/mf2/test.csv ) > gpurun_out/r2_mf_n1.txt 2>&1
echo "n1 rc=$?"; grep -E "^TEST|Time taken|cu2b:" gpurun_out/r2_mf_n1.txt
md5sum /tmp/mf2/train_f64_q.csv | cut -c1-12
( time bin/mf -c /tmp/mf2/two.cfg /tmp/mf2/train.csv /tmp/mf2/test.csv ) > gpurun_out/r2_mf_n2.txt 2>&1
echo "n2 rc=$?"; grep -E "^TEST|Time taken|cu2b:|what" gpurun_out/r2_mf_n2.txt
ls -la /tmp/mf2/*_f64_*.csv | awk '{print $5, $9}'
# the written model reproduces the logged test RMSE (oracle-free check through the library's own loss entry point)
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, cu2rec_b200 as cu
te, rows, cols, _ = cu.readCSV("/tmp/mf2/test.csv")
P, r, c = cu.read_array("/tmp/mf2/train_f64_p.csv"); Q, _, _ = cu.read_array("/tmp/mf2/train_f64_q.csv")
ub, _, _ = cu.read_array("/tmp/mf2/train_f64_user_bias.csv"); ib, _, _ = cu.read_array("/tmp/mf2/train_f64_item_bias.csv")
gb, _, _ = cu.read_array("/tmp/mf2/train_f64_global_bias.csv")
U, I = ub.size, ib.size
m = cu.createSparseMatrix(te, U, I)
print("loss of the files written by the 2-GPU run (mae, rmse):", cu.loss(P, Q, 64, m, ub, ib, gb[0]))
PY
