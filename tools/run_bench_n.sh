#!/bin/bash
# usage: tools/run_bench_n.sh <ngpus> <tag> [extra bench args...]   (env passes through)
N=$1; TAG=$2; shift 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 4 --warmup 3 "$@" > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.log
echo "rc=$? $(wc -c < gpurun_out/bench_${TAG}.json) bytes"
