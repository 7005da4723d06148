#!/bin/bash
# same-box A/B of an environment switch inside the GPT step:  VAR=DLGRAD_TC_STREAMK bash scripts/gpu_ab.sh
VAR=${VAR:-DLGRAD_TC_STREAMK}
for v in 1 0 1 0; do
  env $VAR=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-ops --no-cpu 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$VAR=$v', round(d['ms_per_step'],2), 'ms/step', round(d['value'],2), 'steps/s')"
done
