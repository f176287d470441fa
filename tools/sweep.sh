#!/bin/bash
# runs bench.py against each experimental library variant in tools/variants/
for f in openair4g_b200/lib/liboai_turbo_b200.so tools/variants/lib_*.so; do
  OAI_TURBO_LIB=$PWD/$f python bench.py --steps 4 --warmup 3 --no-cpu --e2e-blocks 512 --blocks 42624 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$f', round(d['value']), d['roofline']['kernel_ms'])"
done
