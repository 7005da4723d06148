#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tc_probe.py v4 > gpurun_out/tc_probe_v4.log 2>&1; echo "exit $?" >> gpurun_out/tc_probe_v4.log; grep -c rel_err gpurun_out/tc_probe_v4.log; grep "FAILED\|exit\|accuracy" gpurun_out/tc_probe_v4.log
: > gpurun_out/tc_trace3.txt
for g in 1 2 3; do
while read -r a; do
  echo "## batch=$g" >> gpurun_out/tc_trace3.txt
  DLGRAD_TC_MMA_BATCH=$g TC_TRACE_ITERS=20 DLGRAD_TC_TRACE=3 timeout 120 python scripts/tc_trace.py $a >> gpurun_out/tc_trace3.txt 2>&1 || echo "FAIL $a" >> gpurun_out/tc_trace3.txt
done <<'EOL'
4096 4096 4096 128 0 4
4096 4096 4096 192 0 4
8192 768 768 192 0 4
8192 768 768 128 0 4
1024 64 1024 64 0 4
EOL
done
grep -v "^#" gpurun_out/tc_trace3.txt | paste - - | sed 's/CTA0: //'
