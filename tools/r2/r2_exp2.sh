#!/bin/bash
# Round-2 GPU call 2 (one B200): row placement micro-benchmark + stability map of the DSGD sub-epoch kernel.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
( time ./tools/micro/l2_rows placement > $O/r2_l2_rows_placement.jsonl ) 2> $O/r2_l2_rows_placement.err
echo "micro rc=$?"
M="none:0,bias:0.125,bias:0.25,bias:0.5,rows:0.125,rows:0.25,rows:0.5"
python tools/dsgd_stability_map.py nfblock8 148,296,444,592,888,1184 $M > $O/r2_stability_nfblock8.jsonl 2> $O/r2_stability_nfblock8.err
echo "map8 rc=$?"
python tools/dsgd_stability_map.py nfblock4 296,444,592,888,1184 $M > $O/r2_stability_nfblock4.jsonl 2> $O/r2_stability_nfblock4.err
echo "map4 rc=$?"
python tools/dsgd_stability_map.py nfblock2 592,888,1184 none:0,bias:0.25,bias:0.5,rows:0.25 > $O/r2_stability_nfblock2.jsonl 2> $O/r2_stability_nfblock2.err
echo "map2 rc=$?"
# longer horizon at full occupancy: does a run that is stable at 512 iterations stay stable at 2000?
SWEEP_ITERS=500 SWEEP_CHECKS=4 python tools/dsgd_stability_map.py nfblock8 1184 bias:0.25,bias:0.5,rows:0.25,rows:0.5 > $O/r2_stability_nfblock8_long.jsonl 2> $O/r2_stability_nfblock8_long.err
echo "map8 long rc=$?"
