#!/bin/bash
# One multi-GPU session on an N-GPU box: link ceiling at 1, 2, 4, ... ranks, then the bench line at the larger rank counts.
# Usage: tools/scale_session.sh <tag> <max_gpus>; files land in gpurun_out/.
tag=${1:-rX}; maxn=${2:-8}
out=gpurun_out; mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
lscpu | grep -E "^CPU\(s\)|Model name|Socket|NUMA node\(s\)|Thread" >> $out/${tag}_topo.txt
free -g | head -2 >> $out/${tag}_topo.txt
port=29600
for n in 1 2 4 8; do
  [ $n -gt $maxn ] && break
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port tools/link_ceiling_probe.py \
      2> $out/${tag}_link_n$n.err | grep '^{' > $out/${tag}_link_n$n.jsonl
done
for n in 8 4; do
  [ $n -gt $maxn ] && continue
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 10 --warmup 3 --no-cpu \
      2> $out/${tag}_bench_n$n.err | grep '^{' > $out/${tag}_bench_n$n.json
done
ls -la $out | grep ${tag}_
