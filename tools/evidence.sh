#!/bin/bash
# Round evidence on one B200: bench lines (own arm, reference arm), the ncu launch list of the bench command and one
# `ncu --set full` capture of the steady-state kernels.  Usage: tools/evidence.sh <tag>; files land in gpurun_out/.
# Afterwards, here:  python tools/ncu_summary.py gpurun_out/<tag>_full16.ncu-rep --json profiles/ncu_kmap16.json k_map16 23680 6144
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
# launch list (never a bench value): first 400 launches of a short run
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-regimes --no-multicell --e2e-blocks 512 > $out/${tag}_ncu_launches.log 2>&1
# full capture: one steady-state iteration of the 16-bit decoder (map, x1, map, x2, compact ...) at 23 680 blocks
ncu --set full --clock-control none --import-source on --launch-skip 70 --launch-count 8 -f -o $out/${tag}_full16 \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-regimes --no-multicell --e2e-blocks 512 --blocks 23680 > $out/${tag}_ncu_full16.log 2>&1
ls -la $out | grep ${tag}_
timeout 120 python tools/exact_path_probe.py > $out/${tag}_exact_path_probe.txt 2>&1
timeout 120 python tools/amplitude_probe.py > $out/${tag}_amplitude_probe.txt 2>&1
timeout 200 python tools/batch_k_sweep.py > $out/${tag}_batch_k_sweep.txt 2>&1
