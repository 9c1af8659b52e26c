#!/bin/bash
# Round-2 GPU call 40 (one B200): batched predict with the ring-buffered epilogue -- bit-exact tests, then the full-size line at k = 128.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_predict.py tests/test_at_size.py -m gpu -x -q -k "predict" 2>&1 | tail -4
PREDICT_K=128 timeout 100 python tests/predict_full_size.py 2>/dev/null | tee gpurun_out/r2_predict_ring.jsonl
