#!/bin/bash
for v in 1 0 1 0; do
  DLGRAD_FUSED_ATTN_BWD=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-ops --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bwd fused=$v', round(d['ms_per_step'],2), 'ms/step', round(d['value'],2), 'steps/s')"
done
