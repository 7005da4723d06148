#!/bin/bash
mkdir -p gpurun_out
for part in v4 sk; do
timeout 400 python scripts/tc_probe.py $part > gpurun_out/tc_probe_$part.log 2>&1; echo "exit $?" >> gpurun_out/tc_probe_$part.log; grep -c rel_err gpurun_out/tc_probe_$part.log; grep "FAILED\|exit\|accuracy\|U\[0,1)" gpurun_out/tc_probe_$part.log
done
grep "^perf" gpurun_out/tc_probe_sk.log | grep "sk=0"
: > gpurun_out/tc_trace3.txt
for g in 1 2; do
while read -r a; do
  echo -n "issuers=$g  $a : " >> gpurun_out/tc_trace3.txt
  DLGRAD_TC_ISSUERS=$g TC_TRACE_ITERS=20 DLGRAD_TC_TRACE=3 timeout 120 python scripts/tc_trace.py $a 2>&1 | tail -1 >> gpurun_out/tc_trace3.txt || echo "FAIL $a" >> gpurun_out/tc_trace3.txt
done <<'EOL'
4096 4096 4096 128 0 4
4096 4096 4096 192 0 4
8192 768 768 192 0 4
8192 768 768 128 0 4
1024 64 1024 64 0 4
EOL
done
cat gpurun_out/tc_trace3.txt
