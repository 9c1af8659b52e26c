#!/bin/bash
# Round-2 GPU call 45 (one B200): batched predict with four TMEM accumulators + ring-buffered epilogue (normal build):
# bit-exact tests, then the full-size lines.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_predict.py tests/test_at_size.py -m gpu -x -q -k "predict" 2>&1 | tail -2
for k in 128 300 50; do PREDICT_K=$k timeout 25 python tests/predict_full_size.py 2>/dev/null | tee -a gpurun_out/r2_predict_acc4.jsonl; done
