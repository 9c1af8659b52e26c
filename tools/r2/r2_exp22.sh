#!/bin/bash
# Final-tree bench lines at N = 8 and N = 4 (as the driver launches them) + reference arm under torchrun.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { # n tag args...
  local n=$1 tag=$2; shift 2
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $n "$@" > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.log
  echo "$tag rc=$? $(wc -c < gpurun_out/bench_${tag}.json) bytes"
}
run 8 r2_n8_final --steps 4 --warmup 4
run 4 r2_n4_final --steps 4 --warmup 4
run 8 r2_n8_final_default_steps
tail -3 gpurun_out/bench_r2_n8_final.log
