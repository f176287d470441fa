#!/bin/bash
# bench line under torchrun at the given rank counts on a multi-GPU box.  Usage: tools/scale_bench.sh <tag> <n> [<n> ...]
tag=$1; shift
out=gpurun_out; mkdir -p $out
port=29700
for n in "$@"; do
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 10 --warmup 3 --no-cpu \
      2> $out/${tag}_bench_n$n.err | grep '^{' > $out/${tag}_bench_n$n.json
done
ls -la $out | grep ${tag}_
